"""Host-side mirror of the reference interface (CPU): schedules, wrap, noise, loader, error behaviour."""
import json
import os

import numpy as np
import pytest
import torch

PI32 = float(np.float32(np.pi))  # the reference wraps in fp32: -float32(pi) is a legal value and |it| > math.pi

from conftest import load_golden, mini_state_dict
from foldingdiff_b200 import _native, beta_schedules, datasets, engine, modelling, sampling, utils

SEED = 7344


def test_product_schedules_match_reference_golden():
    g = load_golden("schedules.npz")
    for kw in ("cosine", "linear", "quadratic"):
        for T in (1000, 250, 100):
            tab = beta_schedules.compute_alphas(beta_schedules.get_variance_schedule(kw, T))
            assert set(tab) == {"betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod",
                                "sqrt_one_minus_alphas_cumprod", "posterior_variance"}
            for k, v in tab.items():
                assert np.array_equal(v.numpy(), g[f"{kw}_{T}_{k}"]), (kw, T, k)
    with pytest.raises(ValueError):
        beta_schedules.get_variance_schedule("sigmoid", 10)


def test_step_coefficients_are_the_scalars_p_sample_selects():
    betas = beta_schedules.get_variance_schedule("cosine", 1000)
    coef = beta_schedules.step_coefficients(betas)
    tab = beta_schedules.compute_alphas(betas)
    assert coef.shape == (1000, 4) and coef.dtype == torch.float32
    assert torch.equal(coef[:, 0], 1.0 / torch.sqrt(tab["alphas"]))
    assert torch.equal(coef[:, 1], betas)
    assert torch.equal(coef[:, 2], tab["sqrt_one_minus_alphas_cumprod"])
    assert torch.equal(coef[:, 3], torch.sqrt(tab["posterior_variance"]))
    assert abs(float(coef[999, 0]) - 99.9917068) < 1e-3 and float(coef[0, 3]) == 0.0


def test_wrap_known_answers():
    # the reference's tests/test_utils.py known answers
    assert utils.modulo_with_wrapped_range(3, -2, 2) == -1
    assert utils.modulo_with_wrapped_range(-3, -2, 2) == 1
    assert np.allclose(utils.modulo_with_wrapped_range(np.array([3, -3, 0, 2, -2]), -2, 2), [-1, 1, 0, -2, -2])
    assert utils.modulo_with_wrapped_range(5, 0, 4) == 1
    g = load_golden("wrap.npz")
    vals = torch.from_numpy(g["vals"])
    assert np.array_equal(utils.modulo_with_wrapped_range(vals.clone(), -np.pi, np.pi).numpy(), g["wrapped"])
    assert np.array_equal(utils.modulo_with_wrapped_range(vals.clone()).numpy(), g["wrapped_default"])


def test_sample_noise_matches_reference_golden():
    g = load_golden("noise.npz")
    d = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset("canonical-full-angles", pad=128),
                                     timesteps=100, beta_schedule="cosine")
    torch.manual_seed(SEED)
    assert np.array_equal(d.sample_noise(torch.zeros(4, 128, 6)).numpy(), g["noise"])
    d2 = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset("canonical-full-angles", pad=128),
                                      timesteps=100, beta_schedule="cosine", angular_variance=0.5)
    torch.manual_seed(SEED)
    assert np.array_equal(d2.sample_noise(torch.zeros(4, 128, 6)).numpy(), g["noise_var05"])
    assert d.timesteps == 100 and d.pad == 128 and d.feature_is_angular["angles"] == [True] * 6
    assert d.feature_names["angles"][:2] == ["phi", "psi"]


def test_empty_dataset_mean_offset_contract(mini_dir, tmp_path):
    d = datasets.AnglesEmptyDataset.from_dir(mini_dir)
    assert d.pad == 128
    with pytest.raises(NotImplementedError):  # the fixture has no training_mean_offset.npy (reference quirk kept)
        d.get_masked_means()
    off = np.arange(6, dtype=np.float32)
    d2 = datasets.AnglesEmptyDataset("canonical-full-angles", mean_offset=off)
    got = d2.get_masked_means()
    got[0] = 99
    assert d2.get_masked_means()[0] == 0  # returns a copy


def test_forward_noising_getitem():
    base = datasets.SyntheticAnglesDataset(n=3, length=20, pad=32, seed=1)
    d = datasets.NoisedAnglesDataset(base, timesteps=50, beta_schedule="cosine")
    torch.manual_seed(0)
    it = d.__getitem__(1, use_t_val=10)
    assert it["corrupted"].shape == (32, 6) and int(it["t"]) == 10
    expect = it["sqrt_alphas_cumprod_t"] * it["angles"] + it["sqrt_one_minus_alphas_cumprod_t"] * it["known_noise"]
    assert torch.allclose(it["corrupted"], utils.modulo_with_wrapped_range(expect), atol=1e-6)
    assert float(it["corrupted"].abs().max()) <= PI32


def test_from_dir_loads_checkpoint_strictly(mini_dir, tmp_path):
    sd, cfg, targs, _ = mini_state_dict()
    m = modelling.BertForDiffusionBase.from_dir(mini_dir)
    assert m.n_inputs == 6 and m.ft_is_angular == [True] * 6 and len(m.ft_names) == 6
    assert m.config.hidden_size == 192 and m.config.position_embedding_type == "relative_key"
    got = m.state_dict()
    assert set(got) == set(sd)  # same 113 keys as the reference checkpoint
    for k in sd:
        assert torch.equal(got[k], sd[k]), k
    # copy_to writes a loadable minimal snapshot (reference tests/test_transformer.py:165-236)
    snap = tmp_path / "snapshot"
    modelling.BertForDiffusion.from_dir(mini_dir, copy_to=str(snap))
    assert sorted(os.listdir(snap)) == ["config.json", "models", "training_args.json"]
    m2 = modelling.BertForDiffusionBase.from_dir(str(snap))
    for k in sd:
        assert torch.equal(m2.state_dict()[k], sd[k])
    m3 = modelling.BertForDiffusionBase.from_dir(mini_dir, load_weights=False)
    assert not torch.equal(m3.state_dict()["token_decoder.dense1.weight"], sd["token_decoder.dense1.weight"])
    with pytest.raises(IndexError):  # no checkpoint in best_by_train, like the reference
        modelling.BertForDiffusionBase.from_dir(mini_dir, best_by="train")


def test_unsupported_configs_fail_loudly():
    sd, cfg, targs, _ = mini_state_dict()
    bad = dict(cfg, position_embedding_type="absolute")
    with pytest.raises(NotImplementedError):
        modelling.BertForDiffusionBase(modelling.BertConfig(**bad), ft_is_angular=[True] * 6)
    with pytest.raises(NotImplementedError):
        modelling.BertForDiffusionBase(modelling.BertConfig(**cfg), ft_is_angular=[True] * 6, decoder="linear")


def test_no_cpu_fallback(mini_dir):
    m = modelling.BertForDiffusionBase.from_dir(mini_dir)  # parameters on CPU
    x = torch.zeros(2, 16, 6)
    with pytest.raises(_native.NativeError, match="no CPU fallback|CUDA"):
        m(x, torch.zeros(2, dtype=torch.long), attention_mask=torch.ones(2, 16))
    d = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset.from_dir(mini_dir), timesteps=5, beta_schedule="linear")
    with pytest.raises(_native.NativeError):
        sampling.p_sample_loop(m, [16, 16], x, 5, d.alpha_beta_terms["betas"], is_angle=[True] * 6)
    with pytest.raises(TypeError):
        sampling.p_sample_loop(torch.nn.Linear(6, 6), [16], x[:1], 5, d.alpha_beta_terms["betas"])


def test_sample_argument_errors(mini_dir):
    m = modelling.BertForDiffusionBase.from_dir(mini_dir)
    d = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset.from_dir(mini_dir), timesteps=5, beta_schedule="linear")
    with pytest.raises(ValueError):
        sampling.sample(m, d, n=1, sweep_lengths=(60, 60))


def test_time_table_op_order():
    W = torch.randn(96, generator=torch.Generator().manual_seed(0)) * 2 * torch.pi
    tab = engine.gaussian_fourier_table(W, 250)
    t = torch.arange(250)
    proj = t[:, None] * W[None, :] * 2 * torch.pi  # modelling.py:69
    assert torch.equal(tab, torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1))
    # the reference's tests/test_model_subparts.py properties: deterministic, unique over t
    assert torch.equal(tab, engine.gaussian_fourier_table(W, 250))
    assert len({tuple(r.tolist()) for r in tab}) == 250
    perm = torch.randperm(250, generator=torch.Generator().manual_seed(1))
    assert torch.equal(engine.gaussian_fourier_rows(W, perm), tab[perm])


def test_weight_key_order_matches_header():
    keys = engine.weight_key_order(6)
    sd, _, _, _ = mini_state_dict()
    assert len(keys) == _native.W_HEAD + 6 * _native.W_PER_LAYER + _native.W_TAIL == 112
    assert set(keys) == set(sd) - {"time_embed.W"}
    assert keys[4 + 17 + 6] == "encoder.layer.1.attention.self.distance_embedding.weight"


def test_cli_argument_checks_run_before_any_device_or_process_group_work(tmp_path):
    """bin/sample.py: a bad invocation ends at once with the reference's messages - before CUDA, before torch.distributed
    (under torchrun a rank-0 failure after init_process_group would leave the other ranks in a collective)."""
    import subprocess
    import sys
    from conftest import ROOT, mini_state_dict, write_model_dir
    cli = os.path.join(ROOT, "bin", "sample.py")
    res = subprocess.run([sys.executable, cli, "-m", str(tmp_path / "missing"), "-o", str(tmp_path / "o1")], capture_output=True, text=True)
    assert res.returncode != 0 and "is not a directory" in res.stderr
    sd, cfg, targs, ckpt = mini_state_dict()
    mdir = write_model_dir(str(tmp_path / "model"), sd, cfg, targs, ckpt)
    busy = tmp_path / "busy"
    busy.mkdir()
    (busy / "leftover.txt").write_text("x")
    res = subprocess.run([sys.executable, cli, "-m", mdir, "-o", str(busy)], capture_output=True, text=True)
    assert res.returncode != 0 and "to be empty" in res.stderr
    res = subprocess.run([sys.executable, cli, "-m", mdir, "-o", str(tmp_path / "o2"), "--testcomparison"], capture_output=True, text=True)
    assert res.returncode != 0 and "CATH" in res.stderr
    env = dict(os.environ, WORLD_SIZE="2", RANK="0", LOCAL_RANK="0")
    res = subprocess.run([sys.executable, cli, "-m", mdir, "-o", str(tmp_path / "o3"), "--fullhistory"], capture_output=True, text=True, env=env)
    assert res.returncode != 0 and "single-GPU only" in res.stderr
