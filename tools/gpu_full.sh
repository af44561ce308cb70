#!/bin/bash
# Full round check: build, every GPU test, smoke, default bench (value + e2e + parity + cpu baseline + config 3 / 5 sub-lines),
# one-step ncu metrics (-> tools/summarize_ncu.py step -> profiles/rNN_step_ncu.json).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "^\.*\[|passed|failed|rror|FAIL" gpurun_out/test_gpu_all.log | sed 's/^\.*//' | grep -v "^\[build\]" | tail -45
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print("value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 2),
          "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "clocks", d["clocks"], "roofline frac", d["roofline"] and round(d["roofline"]["frac"], 3))
    print("parity", json.dumps(d.get("parity"))[:900])
    print("workloads", json.dumps(d.get("workloads"))[:600])
    print({k: round(v['ms_per_reverse_step'], 3) for k, v in d['kernels'].items()})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_default.err").read()[-2500:])
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
bash tools/gpu_ncu_step.sh
