#!/bin/bash
# One reverse step of BASELINE config 2 under ncu: time, DRAM bytes, tensor-pipe / issue activity of every launch.
# Summarise HERE (authoring box) right after it returns:  python tools/summarize_ncu.py step gpurun_out/step_metrics.csv profiles/rNN_step_ncu.json
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0, 'tools'); import summarize_ncu as s; print(s.source_sha())" > gpurun_out/step_source_sha.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
N=${FD_STEP_LAUNCHES:-63}
timeout 900 ncu --metrics $M --clock-control none -k regex:"tc_gemm|attention|embed_kernel|tail_kernel|layernorm" -s $N -c $N --csv \
  --log-file gpurun_out/step_metrics.csv python tools/run_steps.py --steps 1 --warm 1 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
tail -2 gpurun_out/ncu_step.log
