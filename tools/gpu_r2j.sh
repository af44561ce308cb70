#!/bin/bash
# Round-2 call J: CUDA-graph replay of the reverse step: tests, then bench with and without it on the same box.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_edges.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/test_gpu_core.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/test_gpu_core.log
for g in 1 0 1 0; do
  FOLDINGDIFF_B200_GRAPH=$g timeout 600 python bench.py --no-cpu-baseline --no-extra-workloads --no-parity --steps 2 --warmup 3 > gpurun_out/bench_j_$g.json 2> gpurun_out/bench_j_$g.err; echo "bench graph=$g rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_j_$g.json"))
    print("graph=$g value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2), "clocks", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], "kernel sum", round(sum(v['ms_per_reverse_step'] for v in d['kernels'].values()), 3))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_j_$g.err").read()[-2000:])
PY
done
