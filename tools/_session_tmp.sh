mkdir -p gpurun_out/r2b2
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for d in 1 0; do for r in 8192 1024; do
MODA_B200_DEFER_WGRAD=$d python bench.py --no-cpu --no-extra --rays $r --steps 10 2>gpurun_out/r2b2/err_${d}_$r.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('defer=$d rays=$r', d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
python bench.py --workload full --no-cpu --steps 5 2>gpurun_out/r2b2/err_full.txt | tee gpurun_out/r2b2/full.json | cut -c1-400
