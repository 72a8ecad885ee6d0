mkdir -p gpurun_out/r2b12
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 120 python tools/chain_time.py 4 > gpurun_out/r2b12/time_fold.txt 2>&1; sed -n 5,12p gpurun_out/r2b12/time_fold.txt
python bench.py --no-cpu --no-extra --steps 10 2>/dev/null > gpurun_out/r2b12/bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2b12/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['per_entry_ms']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_launch')})"
python bench.py --no-cpu --no-extra --steps 10 --rays 1024 2>/dev/null | cut -c1-160
python bench.py --workload full --no-cpu --steps 5 2>/dev/null | cut -c1-200
