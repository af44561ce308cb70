#!/bin/bash
# tcgen05 attention bring-up: stage-by-stage check, forward parity with the kernel selected, A/B step timing.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
FOLDINGDIFF_B200_ATT=tc timeout 300 python tools/atc_debug.py > gpurun_out/atc_debug.log 2>&1; rc=$?; echo "atc_debug rc=$rc"; tail -12 gpurun_out/atc_debug.log
if [ $rc -eq 0 ]; then
  timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -k "alternative" > gpurun_out/test_alt.log 2>&1; echo "alt paths rc=$?"; tail -3 gpurun_out/test_alt.log
  for att in tc pool; do
    FOLDINGDIFF_B200_ATT=$att timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$att.json 2> gpurun_out/bench_$att.err; echo "bench $att rc=$?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$att.json"))
    print("$att value", round(d["value"],2), "ms/pass", round(d["ms_per_step"],1), {k: round(v['ms_per_reverse_step'],3) for k,v in d['kernels'].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$att.err").read()[-1500:])
PY
  done
fi
