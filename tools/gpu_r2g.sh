#!/bin/bash
# Round-2 call G: 8-warp epilogue + residual from planes (default): bench, core tests, full ncu captures of attention and GEMMs.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_sampling.py -m gpu -q -x > gpurun_out/test_gpu_core.log 2>&1; echo "core gpu tests rc=$?"; tail -2 gpurun_out/test_gpu_core.log
for mode in planes fp32; do
  FOLDINGDIFF_B200_RESID=$mode timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extra-workloads --steps 2 --warmup 3 > gpurun_out/bench_g_$mode.json 2> gpurun_out/bench_g_$mode.err; echo "bench resid=$mode rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_g_$mode.json"))
    print("resid=$mode value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "clocks", d["clocks"]["sm_mhz"], "parity", d["parity"]["vs_fp32_cuda_cores"])
    print({k: round(v['ms_per_reverse_step'], 3) for k, v in d['kernels'].items()}, "sum", round(sum(v['ms_per_reverse_step'] for v in d['kernels'].values()), 3))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_g_$mode.err").read()[-2500:])
PY
done
bash tools/gpu_ncu.sh
