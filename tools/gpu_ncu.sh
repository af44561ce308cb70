#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn4 -f python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 3 -c 5 -o gpurun_out/prof_gemm4 -f python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
