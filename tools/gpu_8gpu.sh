#!/bin/bash
# Eight-GPU check (gpurun --gpus 8): the torchrun CLI test and bench.py --gpus 8 (weak scaling + the sharded-path self-check).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q > gpurun_out/test_gpu_cli_8gpu.log 2>&1; echo "cli tests rc=$?"; tail -2 gpurun_out/test_gpu_cli_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 1 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench 8gpu rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_8gpu.json") if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "value", round(d["value"], 2), "ms/pass", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2), "selfcheck", d["multi_gpu_selfcheck"], "clocks", d["clocks"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_8gpu.err").read()[-2500:])
PY
