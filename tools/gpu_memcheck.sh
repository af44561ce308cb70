#!/bin/bash
# compute-sanitizer memcheck over the smoke path (mini model, both arithmetics, graph replay included) and one
# production-shape forward + 3 reverse steps of a small batch.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|Invalid|smoke\]" gpurun_out/memcheck_smoke.log | tail -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/run_steps.py --steps 2 --warm 1 --batch 16 > gpurun_out/memcheck_steps.log 2>&1; echo "memcheck prod steps rc=$?"; grep -E "ERROR SUMMARY|Invalid|done" gpurun_out/memcheck_steps.log | tail -6
