"""
Variance schedules and the per-step coefficient table of the sampler.

Public names follow /root/reference/foldingdiff/beta_schedules.py (`cosine_beta_schedule` :20,
`linear_beta_schedule` :32, `quadratic_beta_schedule` :38, `compute_alphas` :45, `get_variance_schedule` :65)
so call sites keep working.  What differs is where the tables are used: the reference re-derives all of them
inside every `p_sample` call (sampling.py:42-43); here they are evaluated ONCE on the host - with the
reference's exact fp32 torch op order, because the first cosine reverse step multiplies by
1/sqrt(alpha_{T-1}) = 100 and a 1-ulp difference in a table is a visible difference in the sample - and handed
to the CUDA library as one (T, 4) array (`step_coefficients`).
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, Literal

import torch
from torch.nn.functional import pad

SCHEDULES = Literal["linear", "cosine", "quadratic"]
_CLIP = (0.0001, 0.9999)


def cosine_beta_schedule(timesteps: int, s: float = 8e-3) -> torch.Tensor:
    """Nichol & Dhariwal (arXiv:2102.09672): betas from a squared-cosine alpha-bar, clipped to [1e-4, 0.9999]."""
    grid = torch.linspace(0, timesteps, timesteps + 1)
    abar = torch.cos(((grid / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    abar = abar / abar[0]
    ratio = abar[1:] / abar[:-1]
    return torch.clip(1 - ratio, *_CLIP)


def linear_beta_schedule(timesteps: int, beta_start=1e-4, beta_end=0.02) -> torch.Tensor:
    return torch.linspace(beta_start, beta_end, timesteps)


def quadratic_beta_schedule(timesteps: int, beta_start=1e-4, beta_end=0.02) -> torch.Tensor:
    """Despite its name (kept from the reference) this is a sigmoid ramp between the two betas."""
    return torch.sigmoid(torch.linspace(-6, 6, timesteps)) * (beta_end - beta_start) + beta_start


_BY_NAME: Dict[str, Callable[..., torch.Tensor]] = {
    "cosine": cosine_beta_schedule,
    "linear": linear_beta_schedule,
    "quadratic": quadratic_beta_schedule,
}


def get_variance_schedule(keyword: SCHEDULES, timesteps: int, **kwargs) -> torch.Tensor:
    if keyword not in _BY_NAME:
        raise ValueError(f"Unrecognized variance schedule: {keyword}")
    logging.info(f"Getting {keyword} variance schedule with {timesteps} timesteps")
    return _BY_NAME[keyword](timesteps, **kwargs)


def compute_alphas(betas: torch.Tensor) -> Dict[str, torch.Tensor]:
    """alpha, alpha-bar and the posterior variance derived from betas; same six keys as the reference."""
    out: Dict[str, torch.Tensor] = {"betas": betas}
    out["alphas"] = 1.0 - betas
    abar = torch.cumprod(out["alphas"], dim=0)
    out["alphas_cumprod"] = abar
    out["sqrt_alphas_cumprod"] = torch.sqrt(abar)
    out["sqrt_one_minus_alphas_cumprod"] = torch.sqrt(1.0 - abar)
    abar_before = pad(abar[:-1], (1, 0), value=1.0)  # alpha-bar_{t-1}, with alpha-bar_{-1} := 1
    out["posterior_variance"] = betas * (1.0 - abar_before) / (1.0 - abar)
    return out


def step_coefficients(betas: torch.Tensor) -> torch.Tensor:
    """
    (T, 4) fp32 table {1/sqrt(alpha_t), beta_t, sqrt(1 - alphabar_t), sqrt(posterior_var_t)}: the scalars
    sampling.p_sample selects at sampling.py:43-53 and :72, for every t at once.
    """
    betas = betas.detach().to("cpu", torch.float32)
    tab = compute_alphas(betas)
    cols = (1.0 / torch.sqrt(tab["alphas"]), betas, tab["sqrt_one_minus_alphas_cumprod"],
            torch.sqrt(tab["posterior_variance"]))
    return torch.stack(cols, dim=1).contiguous()
