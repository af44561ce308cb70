#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q -s > gpurun_out/test_k.log 2>&1; echo "kernel tests rc=$?"
grep -E "passed|failed|rror" gpurun_out/test_k.log | tail -5
timeout 900 python -m pytest tests -m gpu -q -s -k "tc3x" > gpurun_out/test_tc3x.log 2>&1; echo "tc3x tests rc=$?"
grep -E "^\[|\]|passed|failed" gpurun_out/test_tc3x.log | tail -12
timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_tc3x.json 2> gpurun_out/bench_tc3x.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_tc3x.json"))
    print("value", d["value"], "ms/pass", d["ms_per_step"])
    for k, v in d["kernels"].items(): print(f"  {k:16s} {v['ms_per_reverse_step']:.3f} ms  share {v['share']:.3f}")
    print(d["roofline"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_tc3x.err").read()[-2000:])
PY
