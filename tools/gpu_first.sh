#!/bin/bash
# First GPU contact: SIMT parity, then the tcgen05 GEMM in isolation, then everything, then a short bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -k "not tc" -x -q -s > gpurun_out/test_simt.log 2>&1; echo "simt tests rc=$?"
tail -5 gpurun_out/test_simt.log
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -k "tc" -q -s > gpurun_out/test_tcgemm.log 2>&1; echo "tc gemm tests rc=$?"
tail -25 gpurun_out/test_tcgemm.log
timeout 900 python -m pytest tests -m gpu -k "tc3x" -q -s > gpurun_out/test_tc3x.log 2>&1; echo "tc3x tests rc=$?"
tail -15 gpurun_out/test_tc3x.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --gemm fp32 --steps 1 --warmup 1 --no-e2e --cpu-chains 64 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?"
cat gpurun_out/bench_fp32.json | head -c 3000
timeout 600 python bench.py --gemm tc3x --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_tc3x.json 2> gpurun_out/bench_tc3x.err; echo "bench tc3x rc=$?"
cat gpurun_out/bench_tc3x.json | head -c 3000
