#!/bin/bash
# Round-2 measurement session on one GPU box: bench lines, ncu launch list of the step, full captures of the main kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session2.sh <tag>'
tag=${1:-r02_sX}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 500 python bench.py > $out/bench_train.json 2> $out/bench_train.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --graph off > $out/launches_bench.log 2>&1
python tools/launch_summary.py $out/launches.csv > $out/launches.summary.txt; head -30 $out/launches.summary.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:chain_kernel --launch-skip 2 -c 2 -f -o $out/chain_trunk python tools/chain_profile.py trunk > $out/ncu_trunk.log 2>&1
timeout 600 $NCU -k regex:chain_kernel --launch-skip 2 -c 2 -f -o $out/chain_skin python tools/chain_profile.py skin > $out/ncu_skin.log 2>&1
timeout 600 $NCU -k regex:tc_wgrad --launch-skip 3 -c 3 -f -o $out/wgrad python tools/chain_profile.py trunk > $out/ncu_wgrad.log 2>&1
timeout 600 $NCU -k regex:tc_wgrad_multi --launch-skip 1 -c 1 -f -o $out/wgrad_skin python tools/chain_profile.py skin > $out/ncu_wgrad_skin.log 2>&1
timeout 600 $NCU -k regex:chain_kernel --launch-skip 4 -c 1 -f -o $out/chain_sigma python bench.py --workload grid --steps 1 --warmup 3 --no-cpu > $out/ncu_sigma.log 2>&1
timeout 600 $NCU -k regex:skin_warp_fwd --launch-skip 10 -c 2 -f -o $out/skinwarp_delta python bench.py --workload dqs --steps 1 --warmup 3 --no-cpu > $out/ncu_dqs.log 2>&1
timeout 600 $NCU -k regex:skin_warp --launch-skip 12 -c 4 -f -o $out/skinwarp python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --graph off > $out/ncu_skinwarp.log 2>&1
# summaries are made here (the .ncu-rep files together exceed what gpurun brings back); only the trunk capture travels
for r in chain_trunk chain_skin wgrad wgrad_skin chain_sigma skinwarp_delta skinwarp; do
  python tools/ncu_summary.py $out/$r.ncu-rep > $out/ncu_full_$r.txt 2>/dev/null
done
python tools/make_traffic.py $tag $out/chain_trunk.ncu-rep $out/chain_skin.ncu-rep $out/wgrad.ncu-rep $out/wgrad_skin.ncu-rep $out/skinwarp.ncu-rep \
  "$out/chain_sigma.ncu-rep:chain_trunk_sigma=chain_kernel" "$out/skinwarp_delta.ncu-rep:skin_warp_fwd_delta=skin_warp_fwd" > /dev/null 2>&1
cp profiles/traffic.json $out/traffic.json
rm -f $out/chain_skin.ncu-rep $out/wgrad.ncu-rep $out/wgrad_skin.ncu-rep $out/chain_sigma.ncu-rep $out/skinwarp_delta.ncu-rep $out/skinwarp.ncu-rep
ls -la $out
