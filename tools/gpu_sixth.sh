#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for mode in pair; do
  FOLDINGDIFF_B200_TC_MODE=$mode timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q > gpurun_out/test_gemm_$mode.log 2>&1; echo "gemm tests [$mode] rc=$?"
  tail -2 gpurun_out/test_gemm_$mode.log
done
grep -E "Error|error|assert" gpurun_out/test_gemm_pair.log | head -10
timeout 900 python -m pytest tests -m gpu -q -s -k "tc3x" > gpurun_out/test_tc3x.log 2>&1; echo "tc3x tests rc=$?"
grep -E "^\[|\]|passed|failed" gpurun_out/test_tc3x.log | tail -10
for mode in pair; do
FOLDINGDIFF_B200_TC_MODE=$mode timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench [$mode] rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$mode.json"))
    print("$mode: value", d["value"], "ms/pass", d["ms_per_step"])
    for k, v in d["kernels"].items(): print(f"  {k:16s} {v['ms_per_reverse_step']:.3f} ms  share {v['share']:.3f}")
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$mode.err").read()[-1500:])
PY
done
