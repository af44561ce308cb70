#!/bin/bash
# Two-GPU check (gpurun --gpus 2): the torchrun CLI test that is skipped on one GPU, and bench.py --gpus 2 (self-check of the
# sharded public path against a single-rank rerun on the real NCCL path).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -q -s > gpurun_out/test_gpu_cli_2gpu.log 2>&1; echo "cli tests rc=$?"; tail -4 gpurun_out/test_gpu_cli_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_2gpu.json") if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "selfcheck", d["multi_gpu_selfcheck"], "clocks", d["clocks"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_2gpu.err").read()[-2500:])
PY
