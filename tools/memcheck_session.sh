#!/bin/bash
# compute-sanitizer memcheck of smoke() and of the GPU tests that exercise the chain / weight-gradient kernels:
#   gpurun --timeout 1500 -- 'bash tools/memcheck_session.sh <tag>'
tag=${1:-r02_memcheck}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $out/memcheck_smoke.txt 2>&1; echo "smoke rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q \
  -k "launch_modes or edge_cases or wgrad_multi or folded or deferred or motion_models_on_gpu or flat_parameter" > $out/memcheck_tests.txt 2>&1; echo "tests rc=$?"
tail -n 4 $out/memcheck_smoke.txt; tail -n 4 $out/memcheck_tests.txt
