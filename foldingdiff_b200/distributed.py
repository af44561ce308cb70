"""
Chain sharding across the GPUs of one box.

Independent chains are the only parallel axis of the sampler: there is no exchange inside the
T-step loop.  Each rank (one process per GPU, torch.distributed over NCCL) holds a replica of the
weights (58 MB), runs the native loop on its share of the chains, and ONE all-gather of the finished
`(B_local, N, F)` angle tensors (1.5 MB per 512 chains) restores the reference's output order.
The reference itself samples on a single device (bin/sample.py:286,341-343).

Chains are dealt round-robin (`chain i -> rank i % world`) because `sampling.sample` emits lengths in
ascending order (reference sampling.py:168-175): every rank then sees the same length mix, i.e. equal
work per rank.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def sharded_final_angles(run_fn: Callable[[List[int], torch.Tensor], torch.Tensor],
                         lengths: Sequence[int], noise: torch.Tensor, group=None,
                         gather_device: Optional[torch.device] = None) -> torch.Tensor:
    """
    Run `run_fn(local_lengths, local_noise) -> (B_local, N, F)` final angles on this rank's share of
    the batch and return the full `(B, N, F)` tensor (CPU) in the original chain order on EVERY rank.
    `noise` is the full-batch initial noise, identical on all ranks (drawn from the same seeded CPU
    generator, which keeps the sharded run bit-identical to the single-device one at t = T).
    """
    rank, world = _world(group)
    B = len(lengths)
    mine = shard_indices(B, rank, world)
    local = run_fn([int(lengths[i]) for i in mine], noise[mine])  # (0, N, F) when this rank has no chain
    if world == 1:
        return local.cpu()
    per_rank = (B + world - 1) // world
    dev = gather_device if gather_device is not None else local.device
    send = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=torch.float32, device=dev)
    send[: len(mine)] = local.to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)  # the one collective of the whole job
    out = torch.empty((B,) + tuple(local.shape[1:]), dtype=torch.float32)
    for r in range(world):
        idx = shard_indices(B, r, world)
        out[idx] = recv[r][: len(idx)].cpu()
    return out


def sample_final_sharded(model, lengths: Sequence[int], noise: torch.Tensor, timesteps: int,
                         betas: torch.Tensor, is_angle, group=None, rng: str = "parity") -> torch.Tensor:
    """
    Multi-GPU `p_sample_loop(..., history="final")[-1]`: (B, N, F) final angles on every rank.

    rng="parity" (SURVEY.md section 8e): every rank draws each step's normals for the WHOLE batch from its device
    generator and keeps its own rows, so with the same device seed on every rank the result equals the single-GPU
    `p_sample_loop` on that seed bit for bit.  rng="perf": each rank draws only its own rows (seed the ranks apart).
    """
    from . import sampling

    if rng not in ("parity", "perf"):
        raise ValueError(f"rng must be 'parity' or 'perf', got {rng!r}")
    rank, world = _world(group)
    shard = sampling.NoiseShard(len(lengths), shard_indices(len(lengths), rank, world)) if rng == "parity" else None

    def run(local_lengths, local_noise):
        dev = next(model.parameters()).device
        out = sampling.p_sample_loop(model, local_lengths, local_noise, timesteps, betas, is_angle=is_angle,
                                     disable_pbar=True, history="final", noise_shard=shard)
        return out[-1].to(dev)

    return sharded_final_angles(run, lengths, noise, group=group)


def sample_sharded(model, train_dset, n: int = 10, sweep_lengths=(50, 128), batch_size: int = 512,
                   feature_key: str = "angles", seed: Optional[int] = None, group=None, rng: str = "parity"):
    """
    Multi-GPU counterpart of `sampling.sample(..., history="final")`: every rank builds the same length list and
    draws the same initial noise (same CPU seed), samples its round-robin share of each batch chunk, and one
    all-gather per chunk returns the final angles in the reference's order.  Returns, on every rank, the list of
    `(length, n_features)` arrays with the training mean offset added and angular columns re-wrapped, exactly as
    `sampling.sample` post-processes them (reference sampling.py:205-222).

    rng="parity" (default): `torch.manual_seed(seed)` on every rank and whole-batch step draws sliced per rank:
    `sample_sharded(world = N, seed)` == `sampling.sample(seed, history="final")` on one GPU, bit for bit.
    rng="perf": per-rank device streams `seed + 1000003 * rank`, each rank draws only its own rows.
    """
    import numpy as np

    from . import sampling, utils

    lo, hi = sweep_lengths
    if not lo < hi:
        raise ValueError(f"Minimum length {lo} must be less than maximum {hi}")
    lengths = [l for l in range(lo, hi) for _ in range(n)]
    rank, _ = _world(group)
    if seed is not None:
        torch.manual_seed(seed)  # identical initial noise on every rank (CPU generator); seeds the device too
        if rng == "perf" and torch.cuda.is_available():
            # an independent stream of per-step normals per rank: with the same device seed and per-rank draws,
            # chain k of every rank would be driven by the same z_t sequence
            torch.cuda.manual_seed(seed + 1000003 * rank)
    out = []
    for chunk in utils.seq_to_groups(lengths, batch_size):
        noise = train_dset.sample_noise(torch.zeros((len(chunk), train_dset.pad, model.n_inputs), dtype=torch.float32))
        noise = noise[:, : max(chunk), :]
        final = sample_final_sharded(model, chunk, noise, train_dset.timesteps, train_dset.alpha_beta_terms["betas"],
                                     train_dset.feature_is_angular[feature_key], group=group, rng=rng)
        out.extend(final[i, :l].numpy() for i, l in enumerate(chunk))
    inner = getattr(train_dset, "dset", None)
    if inner is not None and hasattr(inner, "get_masked_means") and inner.get_masked_means() is not None:
        means = inner.get_masked_means()
        out = [s + means for s in out]
        idx = np.where(train_dset.feature_is_angular[feature_key])[0]
        for s in out:
            s[..., idx] = utils.modulo_with_wrapped_range(s[..., idx], -np.pi, np.pi)
    return out
