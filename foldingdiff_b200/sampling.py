"""
Reverse-diffusion sampling: the reference's `sampling` surface on the native CUDA step.

Same call signatures and return types as /root/reference/foldingdiff/sampling.py:
  p_sample (:28), p_sample_loop (:79), sample (:135), sample_simple (:227),
  get_reconstruction_error (:288, denoising loop only - scoring needs TMalign/biotite).

How the loop is executed differs (this is the hot path the project replaces):
  * one `fd_p_sample_steps` call runs a whole window of reverse steps on the device: the
    noise-predictor forward, the posterior update, the per-column mod-2pi wrap and the history
    write are kernels enqueued back to back - no per-step `.item()` sync, no B-iteration mask loop,
    no `compute_alphas` per step, no per-step `img.cpu()` (reference sampling.py:42-58, :131);
  * only the valid residues of each chain are computed (the reference computes padded rows and
    throws them away, sampling.py:201-203).  Padded positions of returned tensors are 0 in the
    history and left at their input value in `p_sample`;
  * the per-step normals are still drawn with torch's generator for the model's device, one
    `(B, N, F)` draw per step with t > 0 in the reference's order (sampling.py:73), so a seeded run
    consumes RNG state exactly like the reference on the same device.
"""
from __future__ import annotations

import json
import logging
import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import pandas as pd
import torch
from torch import nn

from . import datasets as dsets
from . import modelling, utils

# reverse steps handed to the device per native call (bounds the pre-drawn noise buffer)
STEP_WINDOW = int(os.environ.get("FOLDINGDIFF_B200_STEP_WINDOW", "50"))

# Where the per-step normals come from (extension; the reference has exactly one way, torch.randn_like on the device
# generator, sampling.py:73 - that is "torch", the default and the parity mode).  "philox": the step kernel draws them
# from the library's counter-based stream (fd_p_sample_steps_philox) - no per-step host work and no noise tensor in
# HBM; seeded from torch's CPU generator once per loop, so torch.manual_seed still makes runs reproducible, but the
# values are NOT torch's.
NOISE_SOURCE = os.environ.get("FOLDINGDIFF_B200_NOISE", "torch")


def set_noise_source(source: str) -> None:
    global NOISE_SOURCE
    assert source in ("torch", "philox"), source
    NOISE_SOURCE = source


def _engine_for(model: nn.Module):
    if not hasattr(model, "native_engine"):
        raise TypeError(
            f"{type(model).__name__} is not a foldingdiff_b200 model: the native sampler runs the "
            "built-in noise predictor only (load one with BertForDiffusionBase.from_dir)")
    return model.native_engine()


def _draw_normal(out: torch.Tensor) -> None:
    """Fill `out` with standard normals from torch's generator for its device: the same stream
    consumption as the reference's `torch.randn_like(x)` (sampling.py:73).  Tests patch this hook to
    feed the CPU oracle's draws to the device."""
    torch.randn(out.shape, device=out.device, dtype=out.dtype, out=out)


def _wrap_mask(is_angle: Union[bool, Sequence[bool]], n_features: int) -> List[bool]:
    if isinstance(is_angle, bool):
        return [is_angle] * n_features
    assert len(is_angle) == n_features
    return [bool(a) for a in is_angle]


@torch.no_grad()
def p_sample(model: nn.Module, x: torch.Tensor, t: torch.Tensor, seq_lens: Sequence[int],
             t_index: torch.Tensor, betas: torch.Tensor) -> torch.Tensor:
    """One posterior step x_t -> x_{t-1} (no wrap; see p_sample_loop).  All entries of `t` must agree."""
    t_unique = torch.unique(t)
    assert len(t_unique) == 1, f"Got multiple values for t: {t_unique}"
    ti = int(t_unique.item())
    eng = _engine_for(model)
    eng.set_schedule(betas)
    lens = [int(l) for l in (seq_lens.reshape(-1).tolist() if torch.is_tensor(seq_lens) else seq_lens)]
    eng.set_batch(lens, x.shape[1])
    out = x.detach().to(torch.float32).contiguous().clone()
    z = None
    if ti > 0:
        z = torch.empty_like(out)
        _draw_normal(z)
    eng.p_sample_steps(out, ti + 1, ti, z, None, [False] * x.shape[-1])
    return out


class NoiseShard:
    """
    Parity-mode RNG of a chain-sharded run: this process holds rows `rows` of a global
    batch of `global_batch` chains.  Every step it draws the FULL `(global_batch, N, F)` tensor from the device
    generator - the very draw the single-GPU loop makes (reference sampling.py:73) - and keeps its own rows, so N
    ranks seeded alike reproduce the 1-GPU run bit for bit.  (The draw is ~3 M normals per step at 8 x 512 chains:
    microseconds.)
    """

    def __init__(self, global_batch: int, rows: Sequence[int]):
        self.global_batch = int(global_batch)
        self.rows = [int(r) for r in rows]
        self._idx = None
        self._full = None

    def draw(self, out: torch.Tensor) -> None:
        """Fill `out` (B_local, N, F) with this shard's rows of the next global draw."""
        shape = (self.global_batch,) + tuple(out.shape[1:])
        if self._full is None or self._full.shape != shape or self._full.device != out.device:
            self._full = torch.empty(shape, device=out.device, dtype=out.dtype)
            self._idx = torch.as_tensor(self.rows, device=out.device, dtype=torch.long)
        _draw_normal(self._full)
        if len(self.rows):
            torch.index_select(self._full, 0, self._idx, out=out)


def _run_steps(eng, x: torch.Tensor, t_start: int, wrap: Sequence[bool],
               history: Optional[torch.Tensor], shard: Optional[NoiseShard] = None) -> None:
    """
    Steps t = t_start-1 .. 0 in windows of STEP_WINDOW reverse steps; draws each step's normals like
    `torch.randn_like(x)`.  `history`, if given, is a HOST tensor (t_start, B, N, F), ideally pinned: each
    window's states are written by the tail kernel into one of two device staging buffers and copied out on a
    side stream while the next window computes, so the device->host transfer of the reference's full-history
    return value (1.56 GB at B = 512, T = 1000) hides behind the compute instead of following it.
    """
    B, N, F = x.shape
    dev = x.device
    stage, copied, side = None, None, None
    philox = NOISE_SOURCE == "philox" and shard is None
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if philox else 0  # one draw from torch's CPU generator
    if history is not None:
        w = min(STEP_WINDOW, t_start)
        stage = [torch.zeros((w, B, N, F), device=dev, dtype=torch.float32) for _ in range(2)]
        copied = [None, None]
        side = torch.cuda.Stream(device=dev)
    done, win = 0, 0
    while done < t_start:
        t_hi = t_start - done
        n = min(STEP_WINDOW, t_hi)
        z = None if philox else torch.empty((n, B, N, F), device=dev, dtype=torch.float32)
        for k in range(n if not philox else 0):
            if t_hi - 1 - k > 0:  # the reference draws nothing at t == 0
                if shard is None:
                    _draw_normal(z[k])
                else:
                    shard.draw(z[k])
        if B == 0:  # a rank without chains in this chunk only keeps its generator in step with the others
            done += n
            continue
        hist_dev = None
        if history is not None:
            buf = win & 1
            if copied[buf] is not None:
                torch.cuda.current_stream(dev).wait_event(copied[buf])  # its previous contents are on the host
            hist_dev = stage[buf][:n]
        if philox:
            eng.p_sample_steps_philox(x, t_hi, t_hi - n, seed, done * B * N * F, hist_dev, wrap)
        else:
            eng.p_sample_steps(x, t_hi, t_hi - n, z, hist_dev, wrap)
        if history is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                side.wait_event(ready)
                history[done:done + n].copy_(hist_dev, non_blocking=True)
                copied[buf] = torch.cuda.Event()
                copied[buf].record(side)
        done += n
        win += 1
    if side is not None:
        side.synchronize()
        eng.check_status()  # the history is complete on the host: a timed-out pipeline must not pass silently


@torch.no_grad()
def p_sample_loop(model: nn.Module, lengths: Sequence[int], noise: torch.Tensor, timesteps: int,
                  betas: torch.Tensor, is_angle: Union[bool, List[bool]] = [False, True, True, True],
                  disable_pbar: bool = False, history: str = "full",
                  noise_shard: Optional[NoiseShard] = None) -> torch.Tensor:
    """
    Returns a CPU tensor of shape (timesteps, batch_size, seq_len, n_ft): entry k is the state after
    the k-th reverse step, entry -1 is x_0.  `history="final"` (extension) returns only that last
    entry, shape (1, batch, seq_len, n_ft), so `result[-1]` means the same thing either way.
    `noise_shard` (extension, distributed.py): draw each step's normals for a global batch and keep this
    process's rows.
    """
    device = next(model.parameters()).device
    eng = _engine_for(model)
    eng.set_schedule(betas, timesteps)
    x = noise.detach().to(device=device, dtype=torch.float32).contiguous().clone()
    B, N, F = x.shape
    logging.info(f"Starting from noise {tuple(noise.shape)} with angularity {is_angle} using {device}")
    if B > 0:
        eng.set_batch([int(l) for l in lengths], N)
    wrap = _wrap_mask(is_angle, F)
    if history == "full":
        # pinned host memory (torch's caching host allocator recycles it across calls); padded rows stay 0
        hist = torch.empty((timesteps, B, N, F), dtype=torch.float32, pin_memory=True)  # every element is copied over
        _run_steps(eng, x, timesteps, wrap, hist, noise_shard)
        return hist
    assert history == "final", history
    _run_steps(eng, x, timesteps, wrap, None, noise_shard)
    valid = torch.arange(N, device=device)[None, :] < torch.as_tensor(list(lengths), device=device)[:, None]
    out = (x * valid[..., None]).unsqueeze(0).cpu()  # synchronises
    eng.check_status()
    return out


def sample(model: nn.Module, train_dset, n: int = 10, sweep_lengths: Optional[Tuple[int, int]] = (50, 128),
           batch_size: int = 512, feature_key: str = "angles", disable_pbar: bool = False,
           trim_to_length: bool = True, history: str = "full") -> List[np.ndarray]:
    """
    Sample `n` chains per length in [sweep_min, sweep_max) (upper bound exclusive, like the
    reference), or `n` chains with lengths from `train_dset.sample_length()` when
    `sweep_lengths` is None.  Returns arrays of shape (timesteps, seq_len, n_ft) in length order
    ((1, seq_len, n_ft) with history="final").  `train_dset` needs: sample_noise, timesteps,
    alpha_beta_terms, feature_is_angular, pad (and optionally dset.get_masked_means()).
    """
    if sweep_lengths is not None:
        lo, hi = sweep_lengths
        if not lo < hi:
            raise ValueError(f"Minimum length {lo} must be less than maximum {hi}")
        lengths = [l for l in range(lo, hi) for _ in range(n)]
    else:
        lengths = [train_dset.sample_length() for _ in range(n)]
    logging.info(f"Sampling {len(lengths)} items in batches of size {batch_size}")
    out: List[np.ndarray] = []
    for chunk in utils.seq_to_groups(lengths, batch_size):
        noise = train_dset.sample_noise(
            torch.zeros((len(chunk), train_dset.pad, model.n_inputs), dtype=torch.float32))
        if trim_to_length:
            noise = noise[:, : max(chunk), :]
        sampled = p_sample_loop(model=model, lengths=chunk, noise=noise, timesteps=train_dset.timesteps,
                                betas=train_dset.alpha_beta_terms["betas"],
                                is_angle=train_dset.feature_is_angular[feature_key],
                                disable_pbar=disable_pbar, history=history)
        # np.array(...): each chain's trimmed slice is copied into pageable memory, so the (pinned, up to ~2 GB)
        # history block goes back to torch's host allocator after every chunk instead of staying page-locked for
        # as long as the caller keeps the list (the reference returns views of a pageable tensor, sampling.py:201)
        out.extend(np.array(sampled[:, i, :l, :].numpy()) for i, l in enumerate(chunk))
        del sampled
    inner = getattr(train_dset, "dset", None)
    if inner is not None and hasattr(inner, "get_masked_means") and inner.get_masked_means() is not None:
        means = inner.get_masked_means()
        logging.info(f"Shifting predicted values by original offset: {means}")
        out = [s + means for s in out]
        angular_idx = np.where(train_dset.feature_is_angular[feature_key])[0]
        for s in out:  # the shift may cross the circle boundary
            s[..., angular_idx] = utils.modulo_with_wrapped_range(s[..., angular_idx], -np.pi, np.pi)
    return out


def sample_simple(model_dir: str, n: int = 10, sweep_lengths: Tuple[int, int] = (50, 128)) -> List[pd.DataFrame]:
    """Load a model directory and sample; one DataFrame of final angles per chain."""
    assert os.path.isdir(model_dir), f"{model_dir} is not a directory (hub ids need network access)"
    with open(os.path.join(model_dir, "training_args.json")) as f:
        targs = json.load(f)
    model = modelling.BertForDiffusionBase.from_dir(model_dir).to("cuda:0")
    dummy = dsets.AnglesEmptyDataset.from_dir(model_dir)
    noised = dsets.NoisedAnglesDataset(dset=dummy, dset_key="angles", timesteps=targs["timesteps"],
                                       exhaustive_t=False, beta_schedule=targs["variance_schedule"],
                                       nonangular_variance=1.0, angular_variance=targs["variance_scale"])
    sampled = sample(model, noised, n=n, sweep_lengths=sweep_lengths, disable_pbar=True, history="final")
    return [pd.DataFrame(s[-1], columns=noised.feature_names["angles"]) for s in sampled]


@torch.no_grad()
def denoise_from(model: nn.Module, corrupted: torch.Tensor, lengths: Sequence[int], noise_timesteps: int,
                 betas: torch.Tensor) -> torch.Tensor:
    """
    The partial-denoise loop of get_reconstruction_error (reference sampling.py:311-330): start from
    x_t at t = noise_timesteps, run t = noise_timesteps-1 .. 0, wrapping EVERY column with the
    default +-pi range after each step (the reference calls modulo_with_wrapped_range(img) there,
    not the per-feature wrap of p_sample_loop).  Returns the (B, N, F) result on `corrupted.device`.
    """
    device = next(model.parameters()).device
    eng = _engine_for(model)
    eng.set_schedule(betas)
    x = corrupted.detach().to(device=device, dtype=torch.float32).contiguous().clone()
    eng.set_batch([int(l) for l in lengths], x.shape[1])
    _run_steps(eng, x, noise_timesteps, [True] * x.shape[-1], None)
    torch.cuda.current_stream(device).synchronize()
    eng.check_status()
    return x


@torch.no_grad()
def get_reconstruction_error(model: nn.Module, dset, noise_timesteps: int = 250, bs: int = 512,
                             score_fn=None):
    """
    Noise every item of `dset` to t = noise_timesteps, denoise it back, and return
    (reconstructed, truth) lists of per-chain DataFrames - plus `score_fn(recon, truth, filename)`
    results when a scorer is given.  The reference scores with NeRF + the TMalign binary
    (sampling.py:267-284, 343-356), which this image does not have; that tail is out of scope.
    """
    model.eval()
    cols = dset.feature_names["angles"]
    recon, truth, files = [], [], []
    for idx_batch in utils.seq_to_groups(list(range(len(dset))), bs):
        items = [dset.__getitem__(i, use_t_val=noise_timesteps) for i in idx_batch]
        corrupted = torch.stack([it["corrupted"] for it in items])
        lengths = [int(it["lengths"]) for it in items]
        img = denoise_from(model, corrupted, lengths, noise_timesteps, dset.alpha_beta_terms["betas"]).cpu()
        for j, (it, i, l) in enumerate(zip(items, idx_batch, lengths)):
            recon.append(pd.DataFrame(img[j, :l].numpy(), columns=cols))
            truth.append(pd.DataFrame(it["angles"][:l].numpy(), columns=cols))
            files.append(dset.filenames[i] if hasattr(dset, "filenames") else None)
    if score_fn is None:
        return recon, truth
    return np.array([score_fn(r, t, f) for r, t, f in zip(recon, truth, files)])
