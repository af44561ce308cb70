"""
ORACLE (test infrastructure) - CPU restatement of the reverse-diffusion loop.

Restates, with the reference's fp32 op order:
  wrap()           foldingdiff/utils.py:87-121   modulo_with_wrapped_range
  sample_noise()   foldingdiff/datasets.py:772-799
  p_sample()       foldingdiff/sampling.py:28-75
  p_sample_loop()  foldingdiff/sampling.py:79-132
  denoise_from()   foldingdiff/sampling.py:311-330 (get_reconstruction_error inner loop)
Pinned against the reference's own functions in tests/golden/make_golden.py
(same model callable, same RNG state => bit-identical tensors on CPU).

The per-step normal draws come from ``torch.randn_like`` on the global CPU
generator exactly like the reference (sampling.py:73), or from an explicit
`z_list` so a GPU run can be fed the identical stream.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import schedules


def wrap(vals, lo: float = -np.pi, hi: float = np.pi):
    """utils.py:99-106: ((v - lo) % (hi - lo)) + lo with Python-float bounds."""
    assert lo <= 0.0 and lo < hi
    span = hi - lo
    return ((vals - lo) % span) + lo


def sample_noise(shape_like: torch.Tensor, is_angular: Sequence[bool],
                 angular_var: float = 1.0, nonangular_var: float = 1.0) -> torch.Tensor:
    """datasets.py:772-799 (global CPU generator)."""
    noise = torch.randn_like(shape_like)
    if angular_var != 1.0 or nonangular_var != 1.0:
        for j in range(noise.shape[-1]):
            noise[..., j] *= angular_var if is_angular[j] else nonangular_var
    idx = np.where(is_angular)[0]
    noise[..., idx] = wrap(noise[..., idx], -np.pi, np.pi)
    return noise


def step_coefficients(betas: torch.Tensor, t_index: int):
    """The three scalars sampling.py:43-53,72 selects for one step."""
    tab = schedules.alpha_tables(betas)
    c1 = (1.0 / torch.sqrt(tab["alphas"]))[t_index]
    beta_t = betas[t_index]
    s_t = tab["sqrt_one_minus_alphas_cumprod"][t_index]
    sigma_t = torch.sqrt(tab["posterior_variance"][t_index])
    return c1, beta_t, s_t, sigma_t


@torch.no_grad()
def p_sample(model, x: torch.Tensor, t: torch.Tensor, seq_lens: Sequence[int],
             betas: torch.Tensor, z: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sampling.py:28-75. `z` overrides the randn_like draw (same shape as x)."""
    uniq = torch.unique(t)
    assert len(uniq) == 1, f"Got multiple values for t: {uniq}"
    ti = int(uniq.item())
    c1, beta_t, s_t, sigma_t = step_coefficients(betas, ti)
    mask = torch.zeros(x.shape[:2])
    for i, n in enumerate(seq_lens):
        mask[i, : int(n)] = 1.0
    mean = c1 * (x - beta_t * model(x, t, attention_mask=mask) / s_t)  # :62-67
    if ti == 0:
        return mean
    if z is None:
        z = torch.randn_like(x)  # :73
    return mean + sigma_t * z  # :75


@torch.no_grad()
def p_sample_loop(model, lengths: Sequence[int], noise: torch.Tensor, timesteps: int,
                  betas: torch.Tensor, is_angle: Union[bool, List[bool]],
                  z_list: Optional[Sequence[torch.Tensor]] = None,
                  start_t: Optional[int] = None, wrap_all: bool = False) -> torch.Tensor:
    """
    sampling.py:79-132 -> (steps, B, N, F); index 0 is the state after the first
    reverse step, index -1 is x_0.  `start_t` (exclusive upper bound of t, default
    `timesteps`) and `wrap_all` give the get_reconstruction_error variant
    (sampling.py:319-330: every column wrapped with the default +-pi range).
    `z_list[k]` is the draw for the k-th executed step (unused at t == 0).
    """
    img = noise.clone()
    b = img.shape[0]
    hi = timesteps if start_t is None else start_t
    out = []
    for k, i in enumerate(reversed(range(hi))):
        z = None if z_list is None else (z_list[k] if i > 0 else None)
        img = p_sample(model, img, torch.full((b,), i, dtype=torch.long), lengths, betas, z=z)
        if wrap_all:
            img = wrap(img)
        elif isinstance(is_angle, bool):
            if is_angle:
                img = wrap(img, -torch.pi, torch.pi)
        else:
            assert len(is_angle) == img.shape[-1]
            for j in range(img.shape[-1]):
                if is_angle[j]:
                    img[:, :, j] = wrap(img[:, :, j], -torch.pi, torch.pi)
        out.append(img.clone())
    return torch.stack(out)


def circular_abs_diff(a: torch.Tensor, b: torch.Tensor, is_angle: Sequence[bool]) -> torch.Tensor:
    """|a-b| with angular columns compared on the circle (a +-pi crossing is not a 2*pi error)."""
    d = (a - b).abs()
    ang = torch.tensor(list(is_angle), dtype=torch.bool)
    dc = torch.minimum(d, (2 * np.pi - d).abs())
    return torch.where(ang, dc, d)
