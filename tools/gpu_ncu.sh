#!/bin/bash
# Full-set ncu captures (source view) of the two kernel families, one launch each, from the second reverse step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 13 -c 1 -o gpurun_out/prof_attn_r2 -f python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 53 -c 4 -o gpurun_out/prof_gemm_r2 -f python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
