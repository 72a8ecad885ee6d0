mkdir -p gpurun_out/r2b7
timeout 120 python tools/chain_time.py 4 > gpurun_out/r2b7/time_default.txt 2>&1; sed -n 5,8p gpurun_out/r2b7/time_default.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --no-cpu --no-extra --steps 10 2>/dev/null | cut -c1-200
