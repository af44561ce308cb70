#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py -m gpu -q > gpurun_out/test_k.log 2>&1; echo "kernel tests rc=$?"; tail -2 gpurun_out/test_k.log
timeout 900 python -m pytest tests -m gpu -q -s -k "tc3x" > gpurun_out/test_tc3x.log 2>&1; echo "tc3x tests rc=$?"
grep -E "^\[|\]|passed|failed|rror" gpurun_out/test_tc3x.log | tail -10
timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_q.json"))
    print("value", round(d["value"],2), "ms/pass", round(d["ms_per_step"],1), {k: round(v['ms_per_reverse_step'],3) for k,v in d['kernels'].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
