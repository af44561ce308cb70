#!/bin/bash
# Round-2 call I: GEMM epilogue variants A/B on one box (library per variant, selected with FOLDINGDIFF_B200_LIB).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for rep in 1 2; do
for lib in p0 p1 p2 d0; do
  L="$PWD/foldingdiff_b200/csrc/libfoldingdiff_b200_$lib.so"
  FOLDINGDIFF_B200_LIB=$L timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extra-workloads --no-parity --steps 1 --warmup 3 > gpurun_out/bench_i_$lib.json 2> gpurun_out/bench_i_$lib.err; echo "bench lib=$lib rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_i_$lib.json"))
    print("lib=$lib value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "clocks", d["clocks"]["sm_mhz"], {k: round(v['ms_per_reverse_step'], 3) for k, v in d['kernels'].items() if 'gemm' in k or k == 'attention'})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_i_$lib.err").read()[-1500:])
PY
done
done
