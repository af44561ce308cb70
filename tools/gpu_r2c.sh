#!/bin/bash
# Round-2 call C: PDL / Philox / status-flag build: full GPU suite, fine beta sweep, bench with and without PDL.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "^\[|\]: |passed|failed|rror|FAIL" gpurun_out/test_gpu_all.log | grep -v "^\[build\]" | tail -60
timeout 600 python tools/rz_sweep.py 0.9e-7 0.95e-7 0.975e-7 1.0e-7 1.05e-7 > gpurun_out/rz_sweep2.txt 2>&1; cat gpurun_out/rz_sweep2.txt
for pdl in 1 0; do
  FOLDINGDIFF_B200_PDL=$pdl timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_pdl$pdl.json"))
    print("pdl=$pdl value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "clocks", d["clocks"])
    print({k: round(v['ms_per_reverse_step'], 3) for k, v in d['kernels'].items()}, "sum", round(sum(v['ms_per_reverse_step'] for v in d['kernels'].values()), 3))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_pdl$pdl.err").read()[-1500:])
PY
done
