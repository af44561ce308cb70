#!/bin/bash
# Round-2 call A: RZ-bias calibration, chain parity with / without the de-bias, baseline bench, full GPU suite.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "probe: $(cat tools/.probe 2>/dev/null)"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
FOLDINGDIFF_B200_RZ=0,0 timeout 600 python tools/rz_calib.py > gpurun_out/rz_calib.txt 2>&1; echo "rz_calib rc=$?"; tail -30 gpurun_out/rz_calib.txt
for rz in "0,0" ""; do
  echo "=== chain parity with FOLDINGDIFF_B200_RZ='$rz'"
  FOLDINGDIFF_B200_RZ="$rz" timeout 600 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_forward.py tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q -s -k "tc3x or gemm or attention" 2>&1 | grep -E "^\[|passed|failed|rror" | tail -30
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/test_gpu_all.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r2a.json"))
    print("value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 2), "clocks", d["clocks"])
    print({k: round(v['ms_per_reverse_step'], 3) for k, v in d['kernels'].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_r2a.err").read()[-1500:])
PY
