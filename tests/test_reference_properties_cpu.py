"""
The properties the reference's own unit tests pin around this path ("must be re-asserted on both
oracle and kernel"), on the CPU side: the oracle and the host mirror.  The kernel side of the same list lives in
tests/test_gpu_forward.py::test_reference_invariances and tests/test_gpu_sampling.py.

  schedules strictly increasing           /root/reference/tests/test_variance_schedules.py:12-40
  time embedding: reproducible, permutes
  with its input, unique per timestep     /root/reference/tests/test_model_subparts.py:53-94
  sampling: same seed -> same chain,
  running on without reseeding -> another /root/reference/tests/test_sampling.py:26-47
"""
import numpy as np
import pytest
import torch

from foldingdiff_b200 import beta_schedules, engine
from oracle import forward as ofwd
from oracle import loop as oloop
from oracle import schedules as osched


@pytest.mark.parametrize("fn", ["linear_beta_schedule", "cosine_beta_schedule", "quadratic_beta_schedule"])
def test_product_schedules_strictly_increasing(fn):
    betas = getattr(beta_schedules, fn)(100)
    assert bool(torch.all(betas[1:] - betas[:-1] > 0))
    kw = fn.split("_")[0]
    assert torch.equal(betas, osched.betas_for(kw, 100))  # and the oracle's restatement is the same table


@pytest.mark.parametrize("embed", ["host", "oracle"])
def test_time_embedding_reproducible_permutes_and_unique(embed):
    torch.random.manual_seed(6489)
    W = torch.randn(2) * 2 * torch.pi  # embed_dim 4 as in the reference's test: two frequencies, [sin, cos]
    f = (lambda t: engine.gaussian_fourier_rows(W, t)) if embed == "host" else (lambda t: ofwd.time_embedding(W, t))
    t = torch.randint(low=0, high=250, size=(32,))
    x = f(t)
    assert x.shape == (32, 4) and torch.equal(x, f(t))
    idx = torch.randperm(32)
    assert torch.equal(x[idx], f(t[idx]))
    e = f(torch.arange(0, 1000))
    close = (e[:, None, :] - e[None, :, :]).abs() <= 1e-8 + 1e-5 * e[None, :, :].abs()  # torch.allclose, all pairs at once
    same = close.all(-1)
    assert int(same.sum()) == 1000 and bool(torch.equal(same, torch.eye(1000, dtype=torch.bool)))


def test_host_time_table_is_the_oracle_embedding_bit_for_bit():
    g = torch.Generator().manual_seed(0)
    W = torch.randn(192, generator=g) * 2 * torch.pi
    assert torch.equal(engine.gaussian_fourier_table(W, 1000), ofwd.time_embedding(W, torch.arange(1000)))


def test_oracle_sampling_seed_reproducibility_and_sensitivity(mini_oracle):
    model, _, _ = mini_oracle
    T, lens = 6, [20, 13]
    betas = osched.betas_for("cosine", T)

    def run():
        noise = oloop.sample_noise(torch.zeros(2, 20, 6), [True] * 6)
        return oloop.p_sample_loop(model, lens, noise, T, betas, [True] * 6)[-1]

    torch.manual_seed(1234)
    a = run()
    torch.manual_seed(1234)
    b = run()
    c = run()  # the generator has moved on
    valid = torch.arange(20)[None, :] < torch.tensor(lens)[:, None]
    assert torch.equal(a, b)
    assert not np.allclose(a[valid].numpy(), c[valid].numpy())
    assert float(a[valid].abs().max()) <= float(np.float32(np.pi))
