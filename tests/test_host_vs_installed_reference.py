"""
Live differential tests: the host-side mirror (foldingdiff_b200.{beta_schedules,utils,datasets}) against the STOCK
reference package installed in baseline/_ref (see baseline/reference_arm.py).  Skipped where that install is absent.
Rows a9 - a12 of SURVEY.md section 8: schedules, wrap, noise sampling, the dataset metadata loader, forward noising.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from foldingdiff_b200 import beta_schedules, datasets, utils

sys.path.insert(0, os.path.join(ROOT, "baseline"))
import reference_arm  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_arm.available(), reason="baseline/_ref not installed")


@pytest.fixture(scope="module")
def ref():
    sampling, rbeta, rutils, _ = reference_arm.load_reference()
    from foldingdiff import datasets as rdatasets  # type: ignore
    return dict(sampling=sampling, beta=rbeta, utils=rutils, datasets=rdatasets)


@pytest.mark.parametrize("kind", ["linear", "cosine", "quadratic"])
@pytest.mark.parametrize("T", [2, 100, 250, 1000])
def test_schedules_and_alpha_tables_bit_identical(ref, kind, T):
    ours, theirs = beta_schedules.get_variance_schedule(kind, T), ref["beta"].get_variance_schedule(kind, T)
    assert ours.dtype == theirs.dtype and torch.equal(ours, theirs)
    a, b = beta_schedules.compute_alphas(ours), ref["beta"].compute_alphas(theirs)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_wrap_bit_identical_on_random_and_edge_values(ref):
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(4096, generator=g) * 10, torch.tensor([0.0, -0.0, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 1e-30, 1e6, -1e6]),
                   torch.nextafter(torch.tensor([np.pi, -np.pi], dtype=torch.float32), torch.tensor([0.0, 0.0]))])
    assert torch.equal(utils.modulo_with_wrapped_range(x), ref["utils"].modulo_with_wrapped_range(x))
    for lo, hi in [(-2.0, 2.0), (0.0, 1.0), (-np.pi, np.pi)]:
        assert torch.equal(utils.modulo_with_wrapped_range(x, lo, hi), ref["utils"].modulo_with_wrapped_range(x, lo, hi))
    arr = x.numpy().astype(np.float64)
    assert np.array_equal(utils.modulo_with_wrapped_range(arr, -2, 2), ref["utils"].modulo_with_wrapped_range(arr, -2, 2))


def test_empty_dataset_and_noise_source_agree(ref, tmp_path):
    import json
    d = tmp_path / "m"
    d.mkdir()
    (d / "training_args.json").write_text(json.dumps({"angles_definitions": "canonical-full-angles", "max_seq_len": 128}))
    np.save(d / "training_mean_offset.npy", np.linspace(-1, 1, 6))
    ours, theirs = datasets.AnglesEmptyDataset.from_dir(str(d)), ref["datasets"].AnglesEmptyDataset.from_dir(str(d))
    assert ours.feature_names == theirs.feature_names and ours.feature_is_angular == theirs.feature_is_angular and ours.pad == theirs.pad
    assert np.array_equal(ours.get_masked_means(), theirs.get_masked_means())
    for kw in ({}, {"nonangular_variance": 2.0, "angular_variance": 0.5}):
        no = datasets.NoisedAnglesDataset(ours, timesteps=100, beta_schedule="cosine", **kw)
        nt = ref["datasets"].NoisedAnglesDataset(theirs, timesteps=100, beta_schedule="cosine", **kw)
        for k in no.alpha_beta_terms:
            assert torch.equal(no.alpha_beta_terms[k], nt.alpha_beta_terms[k]), k
        shape = torch.zeros(7, 128, 6)
        torch.manual_seed(7344)
        a = no.sample_noise(shape)
        torch.manual_seed(7344)
        b = nt.sample_noise(shape)
        assert torch.equal(a, b)


class _Inner:
    """Minimal inner dataset with the attributes NoisedAnglesDataset reads (both implementations)."""
    feature_names = {"angles": ["phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA"]}
    feature_is_angular = {"angles": [True] * 6}
    pad = 32

    def __init__(self):
        g = torch.Generator().manual_seed(3)
        self.x = utils.modulo_with_wrapped_range(torch.randn(5, 32, 6, generator=g))

    def __len__(self):
        return 5

    def __getitem__(self, index, ignore_zero_center=False):
        mask = torch.zeros(32); mask[:20] = 1
        return {"angles": self.x[index].clone(), "attn_mask": mask, "position_ids": torch.arange(32)}


@pytest.mark.parametrize("schedule", ["linear", "cosine"])
def test_forward_noising_item_equals_the_reference_item(ref, schedule):
    no = datasets.NoisedAnglesDataset(_Inner(), timesteps=50, beta_schedule=schedule)
    nt = ref["datasets"].NoisedAnglesDataset(_Inner(), timesteps=50, beta_schedule=schedule)
    for idx, tval in [(0, None), (3, 10), (4, 49), (1, 0)]:
        torch.manual_seed(100 + idx)
        a = no.__getitem__(idx, use_t_val=tval)
        torch.manual_seed(100 + idx)
        b = nt.__getitem__(idx, use_t_val=tval)
        assert set(a.keys()) == set(b.keys())
        for k in a:
            va, vb = a[k], b[k]
            assert torch.equal(torch.as_tensor(va), torch.as_tensor(vb)), (k, idx, tval)


def test_sample_orchestration_equals_stock_sample(ref, tmp_path, monkeypatch):
    """
    Row a1 (sampling.sample, sampling.py:135-224): length list, chunking, per-chunk noise draws, trimming, per-chain
    slicing, mean-offset shift + re-wrap.  Both sides run the SAME inner loop (the stock p_sample_loop around the
    reference-assembled model, on the CPU), so any difference comes from the orchestration: must be bit-identical.
    """
    import json
    from conftest import mini_state_dict
    from foldingdiff_b200 import sampling as ours
    sd, cfg, _, _ = mini_state_dict()
    model = reference_arm.build_model(sd, cfg)
    d = tmp_path / "m"
    d.mkdir()
    (d / "training_args.json").write_text(json.dumps({"angles_definitions": "canonical-full-angles", "max_seq_len": 128}))
    np.save(d / "training_mean_offset.npy", np.array([-1.47, 0.75, 3.1, 1.94, 2.03, 2.12]))  # the shift crosses +-pi for omega
    T = 3
    o_dset = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset.from_dir(str(d)), timesteps=T, beta_schedule="cosine")
    r_dset = ref["datasets"].NoisedAnglesDataset(ref["datasets"].AnglesEmptyDataset.from_dir(str(d)), timesteps=T, beta_schedule="cosine")
    stock = ref["sampling"]

    def stock_loop(model, lengths, noise, timesteps, betas, is_angle, disable_pbar=False, history="full"):
        return stock.p_sample_loop(model, lengths, noise, timesteps, betas, is_angle=is_angle, disable_pbar=True)
    monkeypatch.setattr(ours, "p_sample_loop", stock_loop)
    for kw in (dict(n=2, sweep_lengths=(10, 14), batch_size=3), dict(n=1, sweep_lengths=(20, 23), batch_size=512)):
        torch.manual_seed(77)
        a = ours.sample(model, o_dset, **kw)
        torch.manual_seed(77)
        b = stock.sample(model, r_dset, disable_pbar=True, **kw)
        assert len(a) == len(b) == kw["n"] * (kw["sweep_lengths"][1] - kw["sweep_lengths"][0])
        for x, y in zip(a, b):
            assert x.shape == y.shape == (T, x.shape[1], 6) and np.array_equal(x, y)
    with pytest.raises(ValueError):
        ours.sample(model, o_dset, n=1, sweep_lengths=(30, 30))
    with pytest.raises(ValueError):
        stock.sample(model, r_dset, n=1, sweep_lengths=(30, 30))
