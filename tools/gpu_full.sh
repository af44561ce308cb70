#!/bin/bash
# Full round check: build, every GPU test, smoke, default bench (value + e2e + cpu baseline), other workloads, launch list.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "bench config5 rc=$?"
timeout 900 python bench.py --workload config3 --steps 1 --warmup 1 --timesteps 1000 --no-cpu-baseline --e2e-history final > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo "bench config3 rc=$?"
python - <<'PY'
import json
for n in ("default", "config5", "config3"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, "value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 2),
              "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "clocks", d["clocks"], "roofline frac", d["roofline"] and round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(n, "parse failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 200 --csv --log-file gpurun_out/launches_r01.csv python tools/run_steps.py --steps 2 --warm 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
