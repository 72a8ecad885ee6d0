#!/bin/bash
# Closing measurement session of round 2 on one GPU box: GPU tests, smoke, bench lines (train / reference arm / default-flag
# step), ncu launch list of the default-flag step and full captures of the Sinkhorn kernels (the chain / weight-gradient /
# skin-warp captures of tools/gpu_session2.sh are unchanged since r02_s26).
#   gpurun --timeout 1500 -- 'bash tools/gpu_session3.sh <tag>'
tag=${1:-r02_sX}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -n 2 $out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 500 python bench.py > $out/bench_train.json 2> $out/bench_train.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 python bench.py --workload full --steps 10 > $out/bench_full.json 2> $out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches_full.csv \
  python bench.py --workload full --steps 1 --warmup 1 > $out/launches_full_bench.log 2>&1
python tools/launch_summary.py $out/launches_full.csv > $out/launches_full.summary.txt; head -12 $out/launches_full.summary.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k "regex:sinkhorn_[mrcg]" --launch-skip 4 -c 4 -f -o $out/sinkhorn python tools/sinkhorn_profile.py > $out/ncu_sinkhorn.log 2>&1
timeout 600 $NCU -k regex:sinkhorn_pass --launch-skip 50 -c 2 -f -o $out/sinkhorn_pass python tools/sinkhorn_profile.py > $out/ncu_sinkhorn_pass.log 2>&1
python tools/ncu_summary.py $out/sinkhorn.ncu-rep > $out/ncu_full_sinkhorn.txt 2>/dev/null
python tools/ncu_summary.py $out/sinkhorn_pass.ncu-rep > $out/ncu_full_sinkhorn_pass.txt 2>/dev/null
rm -f $out/sinkhorn.ncu-rep $out/sinkhorn_pass.ncu-rep
ls -la $out
