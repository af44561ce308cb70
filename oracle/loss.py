"""
ORACLE (test infrastructure) - the training objective of SURVEY.md section 8f rank 3, restated on the CPU.

No CUDA counterpart exists yet: this module and its tests fix WHAT a training-step kernel has to reproduce (values
and the gradient it has to hand to the backward pass) before one is written.  Nothing in the product imports it.

What it follows:
  wrapped smooth-L1 / L1 on angles   /root/reference/foldingdiff/losses.py:12-62
  per-feature loss terms              /root/reference/foldingdiff/modelling.py:553-604 (valid tokens only, one scalar
                                      per feature; angular features use the wrapped loss with beta = pi / 10,
                                      modelling.py:224-233; the others torch's smooth_l1_loss / l1_loss)
  the scalar that is back-propagated  modelling.py:684-691 (mean of the terms; the optional L1 weight penalty is not
                                      restated - every shipped training_args.json has l1 = 0)

Pinned by tests/test_oracle_loss.py: the known answers of the reference's own tests/test_losses.py and doctests, and -
where the stock package is installed in baseline/_ref - bit-identical values against `foldingdiff.losses` and against
`BertForDiffusion._get_loss_terms` itself, plus torch autograd through the reference's functions for the gradient.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch

from .loop import wrap

ANGULAR_BETA = torch.pi / 10  # modelling.py:230-232


def radian_l1(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """losses.py:12-26: both arguments reduced to [0, 2 pi) first, then the signed difference folded to [-pi, pi)."""
    two_pi = 2 * torch.pi
    d = target % two_pi - pred % two_pi
    d = (d + torch.pi) % two_pi - torch.pi
    return d.abs().mean()


def radian_smooth_l1(pred: torch.Tensor, target: torch.Tensor, beta: float = 1.0,
                     circle_penalty: float = 0.0) -> torch.Tensor:
    """losses.py:29-62: Huber on the wrapped difference, optional penalty on whole turns of the prediction."""
    assert pred.shape == target.shape and beta > 0
    d = wrap(target - pred, -torch.pi, torch.pi)
    a = d.abs()
    out = torch.where(a < beta, 0.5 * (d ** 2) / beta, a - 0.5 * beta).mean()
    if circle_penalty > 0:
        out = out + circle_penalty * torch.div(pred.abs(), torch.pi, rounding_mode="trunc").mean()
    return out


def loss_terms(pred: torch.Tensor, known: torch.Tensor, attn_mask: torch.Tensor, is_angular: Sequence[bool],
               loss: str = "smooth_l1", circle_lambda: float = 0.0) -> torch.Tensor:
    """
    modelling.py:571-604: (F,) tensor, entry i = the loss of feature i over the tokens with attn_mask != 0,
    taken in (batch, position) order as `torch.where(mask)` lists them.
    """
    assert pred.shape == known.shape and loss in ("smooth_l1", "l1")
    b, n = torch.where(attn_mask)
    terms = []
    for i, ang in enumerate(is_angular):
        p, k = pred[b, n, i], known[b, n, i]
        if ang and loss == "smooth_l1":
            terms.append(radian_smooth_l1(p, k, beta=ANGULAR_BETA, circle_penalty=circle_lambda))
        elif ang:
            terms.append(radian_l1(p, k))
        elif loss == "smooth_l1":
            terms.append(torch.nn.functional.smooth_l1_loss(p, k))
        else:
            terms.append(torch.nn.functional.l1_loss(p, k))
    return torch.stack(terms)


def training_loss(pred, known, attn_mask, is_angular, loss: str = "smooth_l1") -> torch.Tensor:
    """modelling.py:684-685: the mean of the per-feature terms."""
    return loss_terms(pred, known, attn_mask, is_angular, loss).mean()


def training_loss_grad(pred: torch.Tensor, known: torch.Tensor, attn_mask: torch.Tensor,
                       is_angular: Sequence[bool]) -> torch.Tensor:
    """
    d training_loss / d pred for loss = "smooth_l1", circle_lambda = 0, in closed form - what a fused loss kernel would
    write for the backward pass.  With M valid tokens and F features every valid element carries
        angular:      -clip(wrap(known - pred) / beta, -1, 1) / (M F)     (the wrap has slope 1 almost everywhere)
        non-angular:   clip(pred - known, -1, 1) / (M F)                   (torch's smooth_l1_loss, beta = 1)
    and padded positions carry 0.
    """
    F = pred.shape[-1]
    valid = attn_mask != 0
    M = int(valid.sum())
    g = torch.zeros_like(pred)
    for i, ang in enumerate(is_angular):
        if ang:
            d = wrap(known[..., i] - pred[..., i], -torch.pi, torch.pi)
            gi = -(d / ANGULAR_BETA).clamp(-1.0, 1.0)
        else:
            gi = (pred[..., i] - known[..., i]).clamp(-1.0, 1.0)
        g[..., i] = torch.where(valid, gi / (M * F), torch.zeros_like(gi))
    return g


def expected_loss_of_uninformed_predictor(angular_variance: float = 1.0) -> float:
    """
    Sanity anchor for a training run (not a reference function): a predictor that outputs 0 for a standard-normal
    target pays E[huber_beta(wrap(z))] per angular feature; for sigma = 1 that is ~0.65 with beta = pi / 10.
    Computed by quadrature so tests can bound a freshly initialised model's loss.
    """
    import numpy as np
    z = np.linspace(-12.0, 12.0, 480001) * angular_variance
    w = np.exp(-0.5 * (z / angular_variance) ** 2) / (angular_variance * math.sqrt(2 * math.pi))
    d = (z + math.pi) % (2 * math.pi) - math.pi
    a = np.abs(d)
    h = np.where(a < ANGULAR_BETA, 0.5 * d * d / ANGULAR_BETA, a - 0.5 * ANGULAR_BETA)
    return float(np.trapezoid(h * w, z))
