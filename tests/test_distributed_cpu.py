"""N > 1 host logic on CPU: world_size-2 gloo run of the chain-sharding + single all-gather path."""
import os
import socket
import subprocess
import sys
import textwrap

import torch

from conftest import ROOT
from foldingdiff_b200 import distributed as fdist


def test_round_robin_shards_balance_the_length_sweep():
    lengths = [50 + (i % 78) for i in range(512)]
    shards = [fdist.shard_indices(512, r, 8) for r in range(8)]
    assert sorted(i for s in shards for i in s) == list(range(512))
    sums = [sum(lengths[i] for i in s) for s in shards]
    assert max(sums) - min(sums) <= 0.03 * max(sums)


def test_single_process_passthrough():
    noise = torch.arange(4 * 3 * 2, dtype=torch.float32).reshape(4, 3, 2)
    out = fdist.sharded_final_angles(lambda l, n: n * 2, [3, 3, 2, 1], noise)
    assert torch.equal(out, noise * 2)


def test_noise_shard_rows_of_the_global_draw():
    """Parity-mode RNG: each rank draws the whole batch's normals and keeps its rows; the ranks' rows
    stitched together are the single-process draws, step after step, and an empty shard still advances the stream."""
    from foldingdiff_b200 import sampling
    B, N, F, steps = 5, 7, 6, 3
    torch.manual_seed(3)
    ref = [torch.randn(B, N, F) for _ in range(steps)]
    for world in (2, 8):  # 8 > B: ranks 5..7 hold no chain
        got = [torch.zeros(B, N, F) for _ in range(steps)]
        for rank in range(world):
            rows = fdist.shard_indices(B, rank, world)
            shard = sampling.NoiseShard(B, rows)
            torch.manual_seed(3)
            for k in range(steps):
                z = torch.empty(len(rows), N, F)
                shard.draw(z)
                got[k][rows] = z
        for a, b in zip(ref, got):
            assert torch.equal(a, b)


WORKER = textwrap.dedent("""
    import os, sys, torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from conftest import mini_state_dict
    from foldingdiff_b200 import distributed as fdist
    from oracle import forward as ofwd, loop as oloop, schedules as osched
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    torch.set_num_threads(2)
    sd, cfg, _, _ = mini_state_dict()
    model = ofwd.OracleModel(sd, ofwd.OracleConfig(**cfg), [True] * 6).eval()
    T = 4
    betas = osched.betas_for("linear", T)
    lengths = [24, 17, 24, 9, 20]
    torch.manual_seed(5)                                   # same seed on every rank -> same noise
    noise = oloop.sample_noise(torch.zeros(5, 24, 6), [True] * 6)
    z_all = [torch.randn(5, 24, 6) for _ in range(T)]      # global per-step draws, sliced per rank
    def run(local_lengths, local_noise):
        idx = fdist.shard_indices(5, rank, dist.get_world_size())
        z = [zz[idx] for zz in z_all]
        return oloop.p_sample_loop(model, local_lengths, local_noise, T, betas, [True] * 6, z_list=z)[-1]
    got = fdist.sharded_final_angles(run, lengths, noise)
    ref = oloop.p_sample_loop(model, lengths, noise, T, betas, [True] * 6, z_list=z_all)[-1]
    ok = True
    for i, l in enumerate(lengths):                        # chains are independent: sharding is exact
        ok &= bool(torch.allclose(got[i, :l], ref[i, :l], atol=1e-6))
    print(f"rank {rank} ok={ok} shape={tuple(got.shape)}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
""")


def test_gloo_world2_matches_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % os.path.join(ROOT, "tests"))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "rank 0 ok=True" in res.stdout and "rank 1 ok=True" in res.stdout
