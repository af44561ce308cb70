#!/bin/bash
# tcgen05 attention bring-up: stage-by-stage check, step timing per kernel variant, optional ncu capture.
#   ATTS="tc tc1" (variants to run)  NCU=1 (capture the first variant)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for att in ${ATTS:-tc}; do
  FOLDINGDIFF_B200_ATT=$att timeout 300 python tools/atc_debug.py > gpurun_out/atc_debug_$att.log 2>&1; rc=$?; echo "atc_debug $att rc=$rc"; cut -c1-200 gpurun_out/atc_debug_$att.log | tail -3
  if [ $rc -eq 0 ]; then
    FOLDINGDIFF_B200_ATT=$att timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$att.json 2> gpurun_out/bench_$att.err; echo "bench $att rc=$?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$att.json"))
    print("$att value", round(d["value"],2), "ms/pass", round(d["ms_per_step"],1), {k: round(v['ms_per_reverse_step'],3) for k,v in d['kernels'].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$att.err").read()[-1500:])
PY
  fi
done
if [ "${NCU:-0}" = "1" ]; then
  att=$(echo ${ATTS:-tc} | cut -d' ' -f1)
  FOLDINGDIFF_B200_ATT=$att timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/prof_atc -f python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_atc.log 2>&1; echo "ncu $att rc=$?"
fi
