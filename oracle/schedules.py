"""
ORACLE (test infrastructure) - variance schedules and derived alpha tables.

Restates /root/reference/foldingdiff/beta_schedules.py:
  cosine  :20-29   linear :32-35   quadratic :38-42
  compute_alphas :45-62   get_variance_schedule :65-78
The torch op sequence is kept identical (fp32, same order) because the tables
must be BIT-identical to the reference's: the first cosine reverse step
multiplies by 1/sqrt(alpha) = 100.  Checked against tests/golden/schedules.npz
(written by the reference's own code) and SURVEY.md A.3.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def betas_for(keyword: str, T: int) -> torch.Tensor:
    if keyword == "cosine":  # beta_schedules.py:24-29
        s = 8e-3
        grid = torch.linspace(0, T, T + 1)
        abar = torch.cos(((grid / T) + s) / (1 + s) * torch.pi * 0.5) ** 2
        abar = abar / abar[0]
        return torch.clip(1 - (abar[1:] / abar[:-1]), 0.0001, 0.9999)
    if keyword == "linear":  # :35
        return torch.linspace(1e-4, 0.02, T)
    if keyword == "quadratic":  # :41-42 (it is a sigmoid ramp, despite the name)
        return torch.sigmoid(torch.linspace(-6, 6, T)) * (0.02 - 1e-4) + 1e-4
    raise ValueError(f"Unrecognized variance schedule: {keyword}")


def alpha_tables(betas: torch.Tensor) -> Dict[str, torch.Tensor]:
    """beta_schedules.py:49-62, same six keys."""
    alphas = 1.0 - betas
    abar = torch.cumprod(alphas, dim=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.0)
    return {
        "betas": betas,
        "alphas": alphas,
        "alphas_cumprod": abar,
        "sqrt_alphas_cumprod": torch.sqrt(abar),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
        "posterior_variance": betas * (1.0 - abar_prev) / (1.0 - abar),
    }
