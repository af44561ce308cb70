"""
Chain-level parity at PRODUCTION shape (L 12, H 384, I 768; BASELINE configs 2, 3 and 5), SURVEY.md section 8c protocol
(2)-(4).  Two references are used:
  * the CPU oracle (oracle/forward.py + oracle/loop.py, fp32) on small subsamples - what the oracle finishes in seconds;
  * the library's own fp32 CUDA-core arithmetic (gemm="fp32": sgemm_tn_kernel + attention_simt_kernel), itself held to
    the oracle by the forward / teacher-forced tests, for whole batches at full T - what only a GPU finishes.
The default (benchmarked) arithmetic is gemm="tc3x".
"""
import numpy as np
import pytest
import torch

PI32 = float(np.float32(np.pi))  # the reference wraps in fp32: -float32(pi) is a legal value and |it| > math.pi

from gpu_util import prod_model, prod_state_dict
from foldingdiff_b200 import beta_schedules, datasets, sampling, synthetic
from oracle import forward as ofwd
from oracle import loop as oloop

pytestmark = pytest.mark.gpu
ANG = [True] * 6
SEED = 7344


def circ(a, b):
    return oloop.circular_abs_diff(a, b, ANG)


@pytest.fixture(scope="module")
def prod_oracle():
    return ofwd.OracleModel(prod_state_dict(), ofwd.OracleConfig(**synthetic.PRODUCTION), ANG).eval()


def config2_inputs(T=1000):
    lengths = synthetic.sweep_lengths(512)
    d = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset("canonical-full-angles"), timesteps=T, beta_schedule="cosine")
    torch.manual_seed(SEED)
    noise = d.sample_noise(torch.zeros(512, 128, 6))[:, :127].contiguous()
    return lengths, noise, d.alpha_beta_terms["betas"]


def run_chain(model, lengths, x0, T_total, t_start, betas, z_seed, wrap_all=False):
    """t = t_start-1 .. 0 on the device with normals from a fixed device seed (identical for every arithmetic)."""
    eng = model.native_engine()
    eng.set_schedule(betas, T_total)
    eng.set_batch(lengths, x0.shape[1])
    x = x0.cuda().contiguous().clone()
    g = torch.Generator(device="cuda").manual_seed(z_seed)
    done = 0
    while done < t_start:
        n = min(50, t_start - done)
        z = torch.randn((n,) + tuple(x.shape), device="cuda", generator=g)
        eng.p_sample_steps(x, t_start - done, t_start - done - n, z, None, ANG)
        done += n
    torch.cuda.synchronize()
    eng.check_status()
    return x.cpu()


def valid_diffs(a, b, lengths):
    return torch.cat([circ(a[i, :l], b[i, :l]).reshape(-1) for i, l in enumerate(lengths)])


def test_config2_full_cosine_chain_tc3x_against_fp32_arithmetic():
    """BASELINE config 2 (512 chains, len 50-127, T = 1000 cosine, from t = T): tc3x against the fp32 CUDA-core mode on
    identical noise, next to the floor of the problem itself: the fp32 mode against ITSELF with the initial noise
    perturbed by 1e-7 relative (protocol (4): a full cosine chain from t = T is ill-conditioned for any fp32
    implementation; 1/sqrt(alpha_T) = 100 at the first step)."""
    lengths, noise, betas = config2_inputs()
    T = 1000
    x_tc = run_chain(prod_model("tc3x"), lengths, noise, T, T, betas, 11)
    m32 = prod_model("fp32")
    x_32 = run_chain(m32, lengths, noise, T, T, betas, 11)
    jitter = noise * (1.0 + 1e-7 * torch.randn(noise.shape, generator=torch.Generator().manual_seed(1)))
    x_j = run_chain(m32, lengths, jitter, T, T, betas, 11)
    d, f = valid_diffs(x_tc, x_32, lengths), valid_diffs(x_j, x_32, lengths)
    stats = lambda v: (float(v.median()), float((v < 1e-4).float().mean()), float(v.max()))
    print(f"[config2 T=1000] tc3x vs fp32: median {stats(d)[0]:.3e}, frac<1e-4 {stats(d)[1]:.3f}, max {stats(d)[2]:.3e} | "
          f"floor (fp32 vs fp32 + 1e-7 input jitter): median {stats(f)[0]:.3e}, frac<1e-4 {stats(f)[1]:.3f}, max {stats(f)[2]:.3e}")
    assert bool(torch.isfinite(x_tc).all()) and float(x_tc.abs().max()) <= PI32
    # the tensor-core arithmetic must sit at the problem's own floor: no worse than a few times the self-divergence
    assert stats(d)[0] <= max(4.0 * stats(f)[0], 2e-5)
    assert stats(d)[1] >= min(0.9, stats(f)[1] - 0.1)
    # distributional equality of the finished structures (per-feature circular mean and dispersion)
    def circ_stats(a):
        rows = torch.cat([a[i, :l] for i, l in enumerate(lengths)])
        zc = torch.exp(1j * rows.to(torch.complex64)).mean(dim=0)
        return torch.angle(zc), 1.0 - zc.abs()
    (m1, v1), (m2, v2), (m3, v3) = circ_stats(x_tc), circ_stats(x_32), circ_stats(x_j)
    cd = lambda a, b: float(((a - b + np.pi) % (2 * np.pi) - np.pi).abs().max())
    dm, dv, dm_floor, dv_floor = cd(m1, m2), float((v1 - v2).abs().max()), cd(m3, m2), float((v3 - v2).abs().max())
    print(f"[config2 T=1000] per-feature circular mean / dispersion, tc3x vs fp32: {dm:.2e} / {dv:.2e}; floor: {dm_floor:.2e} / {dv_floor:.2e}")
    assert dm <= max(4.0 * dm_floor, 2e-2) and dv <= max(4.0 * dv_floor, 2e-2)


@pytest.mark.parametrize("gemm", ["tc3x", "fp32"])
def test_config2_teacher_forced_steps_against_cpu_oracle(prod_oracle, gemm):
    """Protocol (2) at production shape: 8 chains of the config-2 mix, 24 reverse steps at the start (t = T-1 ..), in the
    middle and at the end of the cosine schedule; the device is fed the ORACLE's x_t and the same z at every step."""
    lengths_all, noise_all, betas = config2_inputs()
    sub = list(range(5, 512, 64))  # 8 chains, lengths 55 .. 113
    lengths = [lengths_all[i] for i in sub]
    model = prod_model(gemm)
    eng = model.native_engine()
    eng.set_schedule(betas, 1000)
    eng.set_batch(lengths, 127)
    g = torch.Generator().manual_seed(3)
    worst, worst_first = 0.0, 0.0
    for t_hi in (1000, 500, 8):
        x = noise_all[sub].clone() if t_hi == 1000 else oloop.wrap(torch.randn(8, 127, 6, generator=g))
        for t in range(t_hi - 1, t_hi - 9, -1):
            z = torch.randn(8, 127, 6, generator=g)
            ref = oloop.wrap(oloop.p_sample(prod_oracle, x, torch.full((8,), t, dtype=torch.long), lengths, betas, z=z if t > 0 else None))
            got = x.cuda().contiguous().clone()
            eng.p_sample_steps(got, t + 1, t, z.cuda()[None].contiguous(), None, ANG)
            dmax = float(valid_diffs(got.cpu(), ref, lengths).max())
            if t == 999:
                worst_first = max(worst_first, dmax)
            else:
                worst = max(worst, dmax)
            x = ref
    print(f"[{gemm}] config-2 teacher-forced, 24 steps x 8 chains: first step (x100 gain) {worst_first:.3e}, others {worst:.3e}")
    assert worst < 1e-5 if gemm == "fp32" else worst < 2e-5
    assert worst_first < 2e-4


def test_config5_partial_denoise_b512(prod_oracle):
    """BASELINE config 5: 512 chains x 128 residues denoised from t = 250, every column wrapped (sampling.py:330).
    Whole batch: tc3x vs the fp32 CUDA-core mode; 4 chains of it: both against the CPU oracle over all 250 steps."""
    T, t0, B = 1000, 250, 512
    betas = beta_schedules.get_variance_schedule("cosine", T)
    terms = beta_schedules.compute_alphas(betas)
    g = torch.Generator().manual_seed(5)
    x0 = oloop.wrap(torch.randn(B, 128, 6, generator=g) * 0.5)
    eps = torch.randn(B, 128, 6, generator=g)
    corrupted = oloop.wrap(terms["sqrt_alphas_cumprod"][t0] * x0 + terms["sqrt_one_minus_alphas_cumprod"][t0] * eps)
    lengths = [128] * B
    x_tc = run_chain(prod_model("tc3x"), lengths, corrupted, T, t0, betas, 17)
    x_32 = run_chain(prod_model("fp32"), lengths, corrupted, T, t0, betas, 17)
    d = circ(x_tc, x_32)
    print(f"[config5 B=512 from t=250] tc3x vs fp32 arithmetic: max {float(d.max()):.3e}, median {float(d.median()):.3e}")
    assert float(d.max()) < 1e-4
    # 4 chains against the CPU oracle with the same normals (drawn on the device, copied to the oracle)
    sub = [0, 170, 341, 511]
    gz = torch.Generator(device="cuda").manual_seed(17)
    z_all = []
    done = 0
    while done < t0:
        n = min(50, t0 - done)
        z_all.append(torch.randn((n, B, 128, 6), device="cuda", generator=gz)[:, sub].cpu())
        done += n
    z_list = list(torch.cat(z_all))
    ref = oloop.p_sample_loop(prod_oracle, [128] * 4, corrupted[sub], T, betas, ANG, z_list=z_list, start_t=t0, wrap_all=True)[-1]
    for name, x in (("tc3x", x_tc), ("fp32", x_32)):
        e = float(circ(x[sub], ref).max())
        print(f"[config5] {name} vs CPU oracle, 4 chains x 250 steps: max {e:.3e}")
        assert e < 1e-4


def test_config3_forward_4096_chains(prod_oracle):
    """BASELINE config 3 shape: one forward over 4096 chains of 128 residues (524 288 tokens): tc3x vs the fp32 CUDA-core
    mode on every token, and both against the CPU oracle on 6 chains spread over the batch."""
    B = 4096
    g = torch.Generator().manual_seed(9)
    x = oloop.wrap(torch.randn(B, 128, 6, generator=g))
    t = torch.randint(0, 1000, (B,), generator=g)
    mask = torch.ones(B, 128)
    betas = beta_schedules.get_variance_schedule("cosine", 1000)
    out = {}
    for gemm in ("tc3x", "fp32"):
        m = prod_model(gemm)
        m.native_engine().set_schedule(betas, 1000)
        out[gemm] = m(x.cuda(), t.cuda(), attention_mask=mask.cuda()).cpu()
    d = (out["tc3x"] - out["fp32"]).abs()
    print(f"[config3 forward, 4096 x 128] tc3x vs fp32 arithmetic: max {float(d.max()):.3e}, rms {float(d.pow(2).mean().sqrt()):.3e}")
    assert float(d.max()) < 1e-5
    sub = [0, 811, 1622, 2433, 3244, 4095]
    ref = prod_oracle(x[sub], t[sub], attention_mask=mask[sub])
    for gemm in ("tc3x", "fp32"):
        e = float((out[gemm][sub] - ref).abs().max())
        print(f"[config3 forward] {gemm} vs CPU oracle on 6 chains: max {e:.3e}")
        assert e < 1e-5
