#!/bin/bash
# kernel experiments: times the chains with each library variant (gpurun_variants/lib_*.so) and dumps event traces
# with the variants named lib_trace*.so
out=gpurun_out/${1:-v1}; mkdir -p $out
[ -z "$NO_BASE" ] && timeout 200 python tools/chain_time.py > $out/time_base.txt 2>&1
for v in gpurun_variants/lib_*.so; do
  n=$(basename $v .so); n=${n#lib_}
  if [[ $n == trace* ]]; then
    for w in "trunk fwd" "trunk bwd" "skin fwd" "skin bwd"; do
      MODA_B200_LIB=$v timeout 200 python tools/chain_trace.py $w > $out/${n}_${w// /_}.txt 2>&1
    done
  else
    MODA_B200_LIB=$v timeout 200 python tools/chain_time.py > $out/time_$n.txt 2>&1
  fi
done
grep -h -A9 "^lib" $out/time_*.txt | grep -v Warning | grep -v "print(\|Consider"
