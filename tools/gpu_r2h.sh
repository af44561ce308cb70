#!/bin/bash
# Round-2 call H: software-pipelined GEMM epilogue, A/B against the previous epilogue on the same box.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_sampling.py tests/test_gpu_edges.py -m gpu -q -x > gpurun_out/test_gpu_core.log 2>&1; echo "core gpu tests rc=$?"; tail -2 gpurun_out/test_gpu_core.log
for lib in new prev new prev; do
  L="$PWD/foldingdiff_b200/csrc/libfoldingdiff_b200.so"; [ $lib = prev ] && L="$PWD/foldingdiff_b200/csrc/libfoldingdiff_b200_prev.so"
  FOLDINGDIFF_B200_LIB=$L timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extra-workloads --no-parity --steps 1 --warmup 3 > gpurun_out/bench_h_$lib.json 2> gpurun_out/bench_h_$lib.err; echo "bench lib=$lib rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_h_$lib.json"))
    print("lib=$lib value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "clocks", d["clocks"]["sm_mhz"])
    print({k: round(v['ms_per_reverse_step'], 3) for k, v in d['kernels'].items()}, "sum", round(sum(v['ms_per_reverse_step'] for v in d['kernels'].values()), 3))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_h_$lib.err").read()[-2500:])
PY
done
