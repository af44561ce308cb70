"""NeRF row: oracle vs the reference's own output; CUDA kernel vs both."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nerf as onerf

NAMES = ["phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA"]


def test_oracle_reproduces_reference_nerf_bit_for_bit():
    g = load_golden("nerf.npz")
    for tag in ("a", "b", "c", "d"):
        ang = g[f"{tag}_angles"]
        assert np.array_equal(onerf.build_chain(ang, NAMES, center=True), g[f"{tag}_coords_centered"])
        assert np.array_equal(onerf.build_chain(ang, NAMES, center=False), g[f"{tag}_coords_raw"])
    assert np.array_equal(onerf.build_chain(g["r_angles"], NAMES, center=True), g["r_coords_centered"])


def test_oracle_geometry_properties():
    g = load_golden("nerf.npz")
    xyz = g["a_coords_raw"]
    assert np.allclose(xyz[:3], [onerf.N_INIT, onerf.CA_INIT, onerf.C_INIT])
    d = np.linalg.norm(np.diff(xyz, axis=0), axis=1)  # bonds cycle N-CA, CA-C, C-N
    assert np.allclose(d[3::3], 1.46, atol=1e-9) and np.allclose(d[4::3], 1.54, atol=1e-9) and np.allclose(d[2::3], 1.34, atol=1e-9)
    assert np.allclose(g["a_coords_centered"].mean(axis=0), 0, atol=1e-9)


@pytest.mark.gpu
def test_cuda_nerf_matches_reference_golden():
    from foldingdiff_b200 import nerf
    g = load_golden("nerf.npz")
    tags = ["a", "b", "c", "d", "r"]
    L = [g[f"{t}_angles"].shape[0] for t in tags]
    batch = torch.zeros(len(tags), 128, 6)
    for i, t in enumerate(tags):
        batch[i, : L[i]] = torch.from_numpy(g[f"{t}_angles"])
    out = nerf.build_backbone(batch.cuda(), L, NAMES, center=True).cpu().numpy()
    worst = 0.0
    for i, t in enumerate(tags):
        ref = g[f"{t}_coords_centered"]
        err = float(np.abs(out[i, : 3 * L[i]] - ref).max())
        worst = max(worst, err)
        assert np.all(out[i, 3 * L[i]:] == 0)
    print(f"CUDA NeRF vs reference NERFBuilder: max |dx| = {worst:.3e} A (PDB files carry 1e-3 A)")
    assert worst < 1e-3  # float32 trig in the reference vs fp64 here, accumulated over <= 381 placements
    raw = nerf.build_backbone(batch.cuda(), L, NAMES, center=False).cpu().numpy()
    assert float(np.abs(raw[0, :384] - g["a_coords_raw"]).max()) < 1e-3
    assert np.allclose(raw[:, :3], np.stack([onerf.N_INIT, onerf.CA_INIT, onerf.C_INIT]), atol=1e-5)
    # the reference's batched signature (nerf.py:207): uncentred, equal lengths
    a = torch.from_numpy(g["a_angles"])[None].cuda()
    b = nerf.nerf_build_batch(a[..., 0], a[..., 1], a[..., 2], a[..., 3], a[..., 4], a[..., 5]).cpu().numpy()
    assert b.shape == (1, 384, 3) and float(np.abs(b[0] - g["a_coords_raw"]).max()) < 1e-3
    # default bond angles when the columns are absent (canonical-minimal-angles has only phi/psi/omega/tau)
    sub = nerf.build_backbone(batch[:, :, :3].contiguous().cuda(), L, NAMES[:3], center=False).cpu().numpy()
    ref = onerf.build_chain(g["b_angles"][:, :3], NAMES[:3], center=False)
    assert float(np.abs(sub[1, :150] - ref).max()) < 1e-3


@pytest.mark.gpu
def test_cuda_nerf_full_batch_properties():
    from foldingdiff_b200 import nerf, synthetic
    lengths = synthetic.sweep_lengths(512)
    g = torch.Generator().manual_seed(3)
    ang = (torch.rand(512, 127, 6, generator=g) - 0.5) * 2 * np.pi
    ang[..., 3:] = ang[..., 3:].abs() * 0.2 + 1.7
    xyz = nerf.build_backbone(ang.cuda(), lengths, NAMES, center=True).cpu()
    for i in (0, 77, 511):
        n = 3 * lengths[i]
        d = (xyz[i, 1:n] - xyz[i, : n - 1]).norm(dim=1)
        # the first residue is the (non-ideal) 1CRN start frame; every placed bond has its ideal length
        assert torch.allclose(d[3::3], torch.tensor(1.46), atol=2e-4) and torch.allclose(d[4::3], torch.tensor(1.54), atol=2e-4)
        assert torch.allclose(d[2::3], torch.tensor(1.34), atol=2e-4)
        assert float(xyz[i, :n].mean(dim=0).abs().max()) < 1e-4
    assert bool(torch.isfinite(xyz).all())
