#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
FOLDINGDIFF_B200_ATT=tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/prof_atc -f python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_atc.log 2>&1; echo "ncu atc rc=$?"
tail -3 gpurun_out/ncu_atc.log
