"""
Variance schedules and the per-step coefficient table of the sampler.

Same public names as /root/reference/foldingdiff/beta_schedules.py
(`cosine_beta_schedule` :20, `linear_beta_schedule` :32, `quadratic_beta_schedule` :38,
`compute_alphas` :45, `get_variance_schedule` :65).  The tables are evaluated ONCE on the
host with the reference's exact fp32 torch op order - the first cosine reverse step
multiplies by 1/sqrt(alpha_{T-1}) = 100, so a 1-ulp difference in a table is a visible
difference in the sample - and handed to the CUDA library as a (T, 4) array
(`step_coefficients`), instead of being recomputed on every step as the reference does
(sampling.py:42-43).
"""
from __future__ import annotations

import logging
from typing import Dict, Literal

import torch
import torch.nn.functional as F

SCHEDULES = Literal["linear", "cosine", "quadratic"]


def cosine_beta_schedule(timesteps: int, s: float = 8e-3) -> torch.Tensor:
    """Nichol & Dhariwal cosine schedule; betas clipped to [1e-4, 0.9999]."""
    t = torch.linspace(0, timesteps, timesteps + 1)
    abar = torch.cos(((t / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    abar = abar / abar[0]
    return torch.clip(1 - (abar[1:] / abar[:-1]), 0.0001, 0.9999)


def linear_beta_schedule(timesteps: int, beta_start=1e-4, beta_end=0.02) -> torch.Tensor:
    return torch.linspace(beta_start, beta_end, timesteps)


def quadratic_beta_schedule(timesteps: int, beta_start=1e-4, beta_end=0.02) -> torch.Tensor:
    ramp = torch.linspace(-6, 6, timesteps)
    return torch.sigmoid(ramp) * (beta_end - beta_start) + beta_start


def compute_alphas(betas: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The six derived tables, keyed exactly like the reference's dict."""
    alphas = 1.0 - betas
    abar = torch.cumprod(alphas, dim=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
    return {
        "betas": betas,
        "alphas": alphas,
        "alphas_cumprod": abar,
        "sqrt_alphas_cumprod": torch.sqrt(abar),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
        "posterior_variance": post_var,
    }


def get_variance_schedule(keyword: SCHEDULES, timesteps: int, **kwargs) -> torch.Tensor:
    logging.info(f"Getting {keyword} variance schedule with {timesteps} timesteps")
    table = {"cosine": cosine_beta_schedule, "linear": linear_beta_schedule,
             "quadratic": quadratic_beta_schedule}
    if keyword not in table:
        raise ValueError(f"Unrecognized variance schedule: {keyword}")
    return table[keyword](timesteps, **kwargs)


def step_coefficients(betas: torch.Tensor) -> torch.Tensor:
    """
    (T, 4) fp32 table {1/sqrt(alpha_t), beta_t, sqrt(1 - alphabar_t), sqrt(posterior_var_t)}:
    the scalars sampling.p_sample selects at sampling.py:43-53 and :72, for every t.
    """
    betas = betas.detach().to("cpu", torch.float32)
    tab = compute_alphas(betas)
    c1 = 1.0 / torch.sqrt(tab["alphas"])
    return torch.stack([c1, betas, tab["sqrt_one_minus_alphas_cumprod"],
                        torch.sqrt(tab["posterior_variance"])], dim=1).contiguous()
