#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -s > gpurun_out/test_gemm.log 2>&1; echo "gemm tests rc=$?"
grep -E "rel err|passed|failed|rror" gpurun_out/test_gemm.log | tail -30
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -s > gpurun_out/test_attn.log 2>&1; echo "attention tests rc=$?"
grep -E "passed|failed|rror" gpurun_out/test_attn.log | tail -5
timeout 900 python -m pytest tests -m gpu -q -s -k "tc3x" > gpurun_out/test_tc3x.log 2>&1; echo "tc3x tests rc=$?"
grep -E "^\[|\]|passed|failed" gpurun_out/test_tc3x.log | tail -12
for cl in 2 1; do
FOLDINGDIFF_B200_TC_CLUSTER=$([ $cl = 1 ] && echo 1 || echo 0) timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_tc3x_cl$cl.json 2> gpurun_out/bench_tc3x_cl$cl.err; echo "bench cl=$cl rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_tc3x_cl$cl.json"))
    print("cluster $cl: value", d["value"], "ms/pass", d["ms_per_step"])
    for k, v in d["kernels"].items(): print(f"  {k:16s} {v['ms_per_reverse_step']:.3f} ms  share {v['share']:.3f}")
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_tc3x_cl$cl.err").read()[-2000:])
PY
done
