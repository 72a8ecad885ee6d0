#!/bin/bash
# One GPU-box session: parity tests, the bench lines, the ncu launch list and full captures of the top kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag> [what...]'   (what: tests bench ref dqs grid launches full)
tag=${1:-sX}; shift
what=${*:-tests bench ref dqs grid launches full}
out=gpurun_out/$tag
mkdir -p $out
has() { [[ " $what " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest.log
  tail -5 $out/pytest.log
fi
if has bench; then
  timeout 400 python bench.py > $out/bench_train.json 2> $out/bench_train.err; echo "bench rc=$?"
  cat $out/bench_train.json
fi
if has ref; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
  cat $out/bench_reference.json
fi
if has dqs; then
  timeout 400 python bench.py --workload dqs > $out/bench_dqs.json 2> $out/bench_dqs.err; cat $out/bench_dqs.json
fi
if has grid; then
  timeout 400 python bench.py --workload grid > $out/bench_grid.json 2> $out/bench_grid.err; cat $out/bench_grid.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/launches_bench.log 2>&1
  python tools/launch_summary.py $out/launches.csv > $out/launches.summary.txt; head -24 $out/launches.summary.txt
fi
if has full; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel --launch-skip 2 -c 2 \
    -f -o $out/chain_trunk python tools/chain_profile.py trunk > $out/ncu_trunk.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel --launch-skip 2 -c 2 \
    -f -o $out/chain_skin python tools/chain_profile.py skin > $out/ncu_skin.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel --launch-skip 12 -c 3 \
    -f -o $out/wgrad python tools/chain_profile.py trunk > $out/ncu_wgrad.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:skin_warp --launch-skip 12 -c 4 \
    -f -o $out/skinwarp python bench.py --steps 1 --warmup 3 --no-cpu > $out/ncu_skinwarp.log 2>&1
  ls -la $out
fi
