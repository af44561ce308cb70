#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_sampling.py -m gpu -q > gpurun_out/test_s.log 2>&1; echo "sampling tests rc=$?"; tail -3 gpurun_out/test_s.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_e2e.json"))
    print("value", round(d["value"],2), "ms/pass", round(d["ms_per_step"],1), "e2e", d["e2e"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_e2e.err").read()[-1500:])
PY
