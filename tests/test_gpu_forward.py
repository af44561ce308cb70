"""Forward parity: fd_forward (through model(...)) against the fp32 CPU oracle and the golden eps."""
import pytest
import torch

from conftest import load_golden
from gpu_util import FWD_TOL, GEMMS, mini_model, prefix_mask, prod_model, prod_state_dict
from foldingdiff_b200 import synthetic
from oracle import forward as ofwd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gemm", GEMMS)
def test_mini_forward_matches_golden_and_oracle(mini_dir, mini_oracle, gemm):
    g = load_golden("mini_forward.npz")
    x, t, lengths = torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), g["lengths"].tolist()
    mask = prefix_mask(lengths, x.shape[1])
    model = mini_model(mini_dir, gemm)
    eps = model(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    ref = torch.from_numpy(g["eps_f32"])
    err_all = float((eps - ref).abs().max())  # the forward computes every row, like the reference
    err64 = float((eps.double() - torch.from_numpy(g["eps_f64"]))[mask.bool()].abs().max())
    print(f"[{gemm}] mini forward: max|eps - oracle_f32| = {err_all:.3e} (all rows), vs f64 = {err64:.3e}")
    assert err_all < FWD_TOL[gemm]
    live = mini_oracle[0](x, t, attention_mask=mask)
    assert float((eps - live).abs().max()) < FWD_TOL[gemm]


@pytest.mark.parametrize("gemm", GEMMS)
def test_production_shape_forward_matches_golden(gemm):
    g = load_golden("prod_forward.npz")
    sd = prod_state_dict()
    assert abs(float(sd["encoder.layer.11.output.dense.weight"][5, 7]) - float(g["weight_probe"][0])) < 1e-9
    x, t, lengths = torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), g["lengths"].tolist()
    mask = prefix_mask(lengths, 128)
    model = prod_model(gemm)
    eps = model(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    err = float((eps - torch.from_numpy(g["eps_f32"])).abs().max())
    err64 = float((eps.double() - torch.from_numpy(g["eps_f64"])).abs().max())
    print(f"[{gemm}] production-shape forward: max|eps - oracle_f32| = {err:.3e}, vs f64 = {err64:.3e}")
    assert err < FWD_TOL[gemm]


def test_large_ragged_batch_every_cta_walks_many_items():
    """400 chains x 12 heads = 4800 attention items on 148 persistent CTAs: every CTA pipelines ~32 items through its
    TMA ring / TMEM buffers / mbarrier phases (the small cases above give each CTA at most one).  The tensor-core path
    (tcgen05 GEMMs + tcgen05 attention) must agree with the fp32 CUDA-core path - independent kernels for every
    contraction - and with the CPU oracle on a sample of chains."""
    g = torch.Generator().manual_seed(4242)
    lengths = [50 + (i * 7) % 79 for i in range(400)]
    x = torch.randn(400, 128, 6, generator=g)
    t = torch.randint(0, 1000, (400,), generator=g)
    mask = prefix_mask(lengths, 128)
    valid = mask.bool()
    eps_tc = prod_model("tc3x")(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    eps_32 = prod_model("fp32")(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    err = float((eps_tc - eps_32)[valid].abs().max())
    print(f"large batch: max|eps_tc3x - eps_fp32| over valid rows = {err:.3e}")
    assert err < FWD_TOL["tc3x"]
    from foldingdiff_b200 import _native
    assert _native.lib().fd_debug_tc_status() == 0
    pick = [0, 137, 399]
    cfg = ofwd.OracleConfig(**synthetic.PRODUCTION)
    oracle = ofwd.OracleModel(prod_state_dict(), cfg, [True] * 6).eval()
    ref = oracle(x[pick], t[pick], attention_mask=mask[pick])
    assert float((eps_tc[pick] - ref)[valid[pick]].abs().max()) < FWD_TOL["tc3x"]


@pytest.mark.parametrize("gemm", GEMMS)
def test_reference_invariances(mini_dir, gemm):
    """tests/test_transformer.py:83-162 of the reference: determinism, mask invariance, batch order."""
    model = mini_model(mini_dir, gemm)
    g = torch.Generator().manual_seed(6489)
    x = torch.randn(32, 128, 6, generator=g)
    t = torch.randint(0, 250, (32,), generator=g)
    lengths = torch.randint(40, 129, (32,), generator=g).tolist()
    mask = prefix_mask(lengths, 128)
    a = model(x.cuda(), t.cuda(), attention_mask=mask.cuda())
    assert torch.equal(a, model(x.cuda(), t.cuda(), attention_mask=mask.cuda()))
    x2 = x.clone()
    x2[mask == 0] += torch.randn(int((mask == 0).sum()), 6, generator=g)
    b = model(x2.cuda(), t.cuda(), attention_mask=mask.cuda())
    assert torch.allclose(a[mask.bool()], b[mask.bool()], rtol=1e-3, atol=1e-6)
    perm = torch.randperm(32, generator=g)
    c = model(x[perm].cuda(), t[perm].cuda(), attention_mask=mask[perm].cuda())
    assert torch.allclose(a[perm][mask[perm].bool()], c[mask[perm].bool()], atol=2e-6)


def test_non_prefix_mask_and_timestep_shapes(mini_dir, mini_oracle):
    model = mini_model(mini_dir, "fp32")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 40, 6, generator=g)
    mask = (torch.rand(3, 40, generator=g) > 0.3).float()
    mask[:, 0] = 1.0
    t = torch.tensor([[3], [100], [249]])  # (B, 1) like the datasets' "t" entries
    eps = model(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    ref = mini_oracle[0](x, t.squeeze(-1), attention_mask=mask)
    assert float((eps - ref).abs().max()) < 1e-5
    with pytest.raises(AssertionError):
        model(x.cuda(), t.cuda(), attention_mask=mask[None].cuda())  # mask must be 2-D (modelling.py:447)


def test_to_device_roundtrip_and_state_dict(mini_dir):
    model = mini_model(mini_dir, "fp32")
    assert next(model.parameters()).device.type == "cuda"
    x = torch.zeros(1, 8, 6, device="cuda")
    e1 = model(x, torch.zeros(1, dtype=torch.long, device="cuda"), attention_mask=torch.ones(1, 8, device="cuda"))
    with torch.no_grad():
        model.token_decoder.dense2.bias.add_(1.0)  # parameter edits are picked up (engine rebuild)
    e2 = model(x, torch.zeros(1, dtype=torch.long, device="cuda"), attention_mask=torch.ones(1, 8, device="cuda"))
    assert torch.allclose(e2, e1 + 1.0, atol=1e-6)


@pytest.mark.parametrize("env", [{"FOLDINGDIFF_B200_TC_MODE": "mcast"}, {"FOLDINGDIFF_B200_TC_MODE": "single"},
                                 {"FOLDINGDIFF_B200_ATT": "pool"}])
def test_alternative_tensor_core_paths(env, tmp_path):
    """The A/B variants of the tensor-core path (multicast clusters, single CTA, fused GEMM+LayerNorm, the other
    attention kernels) are
    selected by environment at library load, so each runs in its own process: same forward parity gate."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import load_golden\n"
        "from gpu_util import FWD_TOL, prefix_mask, prod_model\n"
        "g = load_golden('prod_forward.npz')\n"
        "x, t, lengths = torch.from_numpy(g['x']), torch.from_numpy(g['t']), g['lengths'].tolist()\n"
        "eps = prod_model('tc3x')(x.cuda(), t.cuda(), attention_mask=prefix_mask(lengths, 128).cuda()).cpu()\n"
        "err = float((eps - torch.from_numpy(g['eps_f32'])).abs().max())\n"
        "from foldingdiff_b200 import _native\n"
        "print('err', err, 'tc_status', _native.lib().fd_debug_tc_status())\n"
        "assert err < FWD_TOL['tc3x'] and _native.lib().fd_debug_tc_status() == 0\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
