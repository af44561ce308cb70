"""Edge cases of the hot path against the CPU oracle: tiny / ragged / boundary shapes, argument errors."""
import numpy as np
import pytest
import torch

from gpu_util import FWD_TOL, GEMMS, mini_model, prefix_mask
from foldingdiff_b200 import _native, beta_schedules, sampling

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gemm", GEMMS)
@pytest.mark.parametrize("lengths,n_pad", [([1], 1), ([1, 2, 3], 3), ([16, 17, 15, 1], 17), ([128], 128), ([31, 32, 33, 64, 65, 96, 97], 97)])
def test_forward_boundary_shapes(mini_dir, mini_oracle, gemm, lengths, n_pad):
    model = mini_model(mini_dir, gemm)
    g = torch.Generator().manual_seed(n_pad * 7 + len(lengths))
    x = torch.randn(len(lengths), n_pad, 6, generator=g)
    t = torch.randint(0, 250, (len(lengths),), generator=g)
    mask = prefix_mask(lengths, n_pad)
    eps = model(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    ref = mini_oracle[0](x, t, attention_mask=mask)
    assert float((eps - ref).abs().max()) < FWD_TOL[gemm]


def test_timestep_outside_the_schedule(mini_dir, mini_oracle):
    model = mini_model(mini_dir, "fp32")
    x = torch.randn(2, 20, 6, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([999, 5000])  # the time embedding is defined for any integer t (modelling.py:59-71)
    mask = torch.ones(2, 20)
    eps = model(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    assert float((eps - mini_oracle[0](x, t, attention_mask=mask)).abs().max()) < 1e-5


def test_p_sample_first_and_last_step(mini_dir, mini_oracle, monkeypatch):
    from oracle import loop as oloop
    model = mini_model(mini_dir, "fp32")
    betas = beta_schedules.get_variance_schedule("cosine", 50)
    x = oloop.wrap(torch.randn(2, 24, 6, generator=torch.Generator().manual_seed(1)))
    z = torch.randn(2, 24, 6, generator=torch.Generator().manual_seed(2))
    monkeypatch.setattr(sampling, "_draw_normal", lambda out: out.copy_(z.to(out.device)))
    for t in (49, 0):
        ref = oloop.p_sample(mini_oracle[0], x, torch.full((2,), t), [24, 20], betas, z=z)
        got = sampling.p_sample(model, x.cuda(), torch.full((2,), t, device="cuda"), [24, 20], t, betas).cpu()
        tol = 2e-3 if t == 49 else 1e-5  # 1/sqrt(alpha_{T-1}) = 100 on the first cosine step
        assert float((got[0] - ref[0]).abs().max()) < tol and float((got[1, :20] - ref[1, :20]).abs().max()) < tol


def test_native_argument_errors(mini_dir):
    model = mini_model(mini_dir, "fp32")
    eng = model.native_engine()
    with pytest.raises(_native.NativeError, match="lengths"):
        eng.set_batch([0, 5], 8)
    with pytest.raises(_native.NativeError, match="lengths"):
        eng.set_batch([9], 8)
    with pytest.raises(_native.NativeError):
        eng.set_batch([4], 129)  # n_pad beyond max_position_embeddings
    eng.set_batch([4, 8], 8)
    eng.set_schedule(beta_schedules.get_variance_schedule("linear", 10))
    x = torch.zeros(2, 8, 6, device="cuda")
    with pytest.raises(_native.NativeError, match="t_lo"):
        eng.p_sample_steps(x, 11, 0, torch.zeros(11, 2, 8, 6, device="cuda"), None, [True] * 6)
    with pytest.raises(_native.NativeError, match="noise"):
        eng.p_sample_steps(x, 5, 0, None, None, [True] * 6)
    eng.p_sample_steps(x, 1, 0, None, None, [True] * 6)  # t = 0 alone needs no noise
    assert bool(torch.isfinite(x).all())
