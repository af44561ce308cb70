#!/usr/bin/env python
"""
Generate the committed golden fixtures under tests/golden/.

Run in the AUTHORING container only (needs /root/reference):
    python tests/golden/make_golden.py

What is produced by the REFERENCE'S OWN CODE (imported from /root/reference
under oracle/ref_shims.py):
  schedules.npz   beta_schedules.get_variance_schedule + compute_alphas tables
  wrap.npz        utils.modulo_with_wrapped_range on fp32 tensors
  noise.npz       NoisedAnglesDataset.sample_noise after torch.manual_seed(7344)
  mini_chain.npz  sampling.p_sample / p_sample_loop (the reference's loop, unmodified)
                  driving the restated oracle forward on the mini fixture's real weights
What is produced by the oracle restatement (HF 4.11.3 encoder is not installable,
"parity unpinned" at that boundary - see oracle/__init__.py):
  mini_forward.npz, prod_forward.npz   eps_hat for seeded inputs, fp32 and fp64
Data (not code) carried over from the reference's test fixture:
  mini_model.npz  tests/mini_model_for_testing/results: config.json, training_args.json
                  and the checkpoint's 113 fp32 tensors
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import forward as ofwd  # noqa: E402
from oracle import loop as oloop  # noqa: E402
from oracle import ref_shims  # noqa: E402
from foldingdiff_b200 import synthetic  # noqa: E402

MINI_DIR = os.path.join(ref_shims.REFERENCE_ROOT, "tests", "mini_model_for_testing", "results")
SEED = 7344  # bin/sample.py:34-37


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez(path, **arrays)
    print(f"wrote {name}: {os.path.getsize(path) / 1e3:.1f} kB")


def main():
    torch.set_num_threads(os.cpu_count())
    ref = ref_shims.load()

    # ---- mini fixture: data files -> one npz -------------------------------------------
    with open(os.path.join(MINI_DIR, "config.json")) as f:
        cfg_json = f.read()
    with open(os.path.join(MINI_DIR, "training_args.json")) as f:
        targs_json = f.read()
    ckpt = os.path.join(MINI_DIR, "models", "best_by_valid", "epoch=19-step=1840.ckpt")
    sd = torch.load(ckpt, map_location="cpu", weights_only=True)["state_dict"]
    save("mini_model.npz", __config_json__=np.array(cfg_json), __training_args_json__=np.array(targs_json),
         __ckpt_name__=np.array(os.path.basename(ckpt)),
         **{k: v.numpy() for k, v in sd.items()})
    cfg = ofwd.OracleConfig(**json.loads(cfg_json))
    ft_ang = ref.datasets.FEATURE_SET_NAMES_TO_ANGULARITY[json.loads(targs_json)["angles_definitions"]]

    # ---- schedules (reference code) ----------------------------------------------------
    sched = {}
    for kw in ("cosine", "linear", "quadratic"):
        for T in (1000, 250, 100):
            betas = ref.beta_schedules.get_variance_schedule(kw, T)
            tab = ref.beta_schedules.compute_alphas(betas)
            for k, v in tab.items():
                sched[f"{kw}_{T}_{k}"] = v.numpy()
    save("schedules.npz", **sched)

    # ---- wrap (reference code) ---------------------------------------------------------
    g = torch.Generator().manual_seed(1)
    vals = torch.cat([torch.tensor([3.5, -3.5, np.pi, -np.pi, 7.0, -7.0, 100.0, 0.0, 3.1415925, -3.1415925],
                                   dtype=torch.float32),
                      torch.randn(4096, generator=g) * 4.0,
                      torch.randn(512, generator=g) * 300.0])
    save("wrap.npz", vals=vals.numpy(),
         wrapped=ref.utils.modulo_with_wrapped_range(vals.clone(), -np.pi, np.pi).numpy(),
         wrapped_default=ref.utils.modulo_with_wrapped_range(vals.clone()).numpy())

    # ---- initial noise (reference code) ------------------------------------------------
    def ref_dset(T, schedule, var_scale=1.0):
        empty = ref.datasets.AnglesEmptyDataset("canonical-full-angles", pad=128, mean_offset=None)
        return ref.datasets.NoisedAnglesDataset(empty, dset_key="angles", timesteps=T, exhaustive_t=False,
                                                beta_schedule=schedule, nonangular_variance=1.0,
                                                angular_variance=var_scale)

    torch.manual_seed(SEED)
    n1 = ref_dset(100, "cosine").sample_noise(torch.zeros(4, 128, 6))
    torch.manual_seed(SEED)
    n2 = ref_dset(100, "cosine", var_scale=0.5).sample_noise(torch.zeros(4, 128, 6))
    save("noise.npz", noise=n1.numpy(), noise_var05=n2.numpy())

    # ---- forward goldens (oracle restatement) ------------------------------------------
    sd32 = {k: v.float() for k, v in sd.items()}
    sd64 = {k: v.double() for k, v in sd.items()}
    g = torch.Generator().manual_seed(1234)  # SURVEY.md A.4 inputs
    x = oloop.wrap(torch.randn(4, 64, 6, generator=g))
    lengths = [64, 50, 33, 64]
    mask = torch.zeros(4, 64)
    for i, n in enumerate(lengths):
        mask[i, :n] = 1.0
    t = torch.tensor([0, 17, 100, 249])
    e32 = ofwd.forward(sd32, cfg, x, t, mask)
    tt32 = ofwd.time_embedding(sd32["time_embed.W"], torch.arange(250))
    e64 = ofwd.forward(sd64, cfg, x.double(), t, mask.double(), time_table=tt32.double())
    print("A.4 check eps[0,0,:]", e32[0, 0].numpy())
    print("A.4 check eps[2,32,:]", e32[2, 32].numpy())
    valid = mask.bool()
    print("A.4 sums", float(e32[valid].sum()), float(e32[valid].abs().sum()))
    save("mini_forward.npz", x=x.numpy(), lengths=np.array(lengths), t=t.numpy(),
         eps_f32=e32.numpy(), eps_f64=e64.numpy())

    pcfg = ofwd.OracleConfig(**synthetic.PRODUCTION)
    psd = synthetic.synthetic_state_dict(synthetic.PRODUCTION, seed=0)
    g = torch.Generator().manual_seed(4321)
    px = oloop.wrap(torch.randn(6, 128, 6, generator=g))
    plen = [128, 127, 50, 77, 1, 100]
    pmask = torch.zeros(6, 128)
    for i, n in enumerate(plen):
        pmask[i, :n] = 1.0
    pt = torch.tensor([999, 0, 500, 3, 250, 998])
    pe32 = ofwd.forward(psd, pcfg, px, pt, pmask)
    ptt32 = ofwd.time_embedding(psd["time_embed.W"], torch.arange(1000))
    pe64 = ofwd.forward({k: v.double() for k, v in psd.items()}, pcfg, px.double(), pt, pmask.double(),
                        time_table=ptt32.double())
    print("prod fp32 vs fp64 max abs", float((pe32.double() - pe64)[pmask.bool()].abs().max()))
    save("prod_forward.npz", x=px.numpy(), lengths=np.array(plen), t=pt.numpy(),
         eps_f32=pe32.numpy(), eps_f64=pe64.numpy(),
         weight_probe=np.array([float(psd["encoder.layer.11.output.dense.weight"][5, 7]),
                                float(psd["time_embed.W"][3])]))

    # ---- loop goldens: the REFERENCE'S loop driving the oracle forward -----------------
    model = ofwd.OracleModel(sd32, cfg, ft_ang).eval()
    chain = {}

    # single reference p_sample step, SURVEY.md A.4 (T=250 cosine, all t=100, seed right before)
    dset250 = ref_dset(250, "cosine")
    torch.manual_seed(SEED)
    y = ref.sampling.p_sample(model, x.clone(), torch.full((4,), 100, dtype=torch.long), lengths, 100,
                              dset250.alpha_beta_terms["betas"])
    print("A.4 check p_sample y[0,0,:]", y[0, 0].numpy(), float(y[valid].sum()))
    chain["step_x"], chain["step_y"], chain["step_lengths"] = x.numpy(), y.numpy(), np.array(lengths)

    def run_ref_chain(tag, T, schedule, lens, pad_to):
        d = ref_dset(T, schedule)
        torch.manual_seed(SEED)
        noise = d.sample_noise(torch.zeros(len(lens), 128, 6))[:, :pad_to]
        hist = ref.sampling.p_sample_loop(model, lens, noise, T, d.alpha_beta_terms["betas"],
                                          is_angle=d.feature_is_angular["angles"], disable_pbar=True)
        # the oracle's own loop must reproduce the reference loop bit for bit
        torch.manual_seed(SEED)
        noise2 = oloop.sample_noise(torch.zeros(len(lens), 128, 6), ft_ang)[:, :pad_to]
        hist2 = oloop.p_sample_loop(model, lens, noise2, T, d.alpha_beta_terms["betas"], ft_ang)
        assert torch.equal(noise, noise2), tag
        assert torch.equal(hist, hist2), tag
        chain[f"{tag}_noise"] = noise.numpy()
        chain[f"{tag}_hist"] = hist.numpy()
        chain[f"{tag}_lengths"] = np.array(lens)
        print(f"{tag}: hist {tuple(hist.shape)} bit-identical between reference loop and oracle loop")

    run_ref_chain("c1_cosine100", 100, "cosine", [64, 64, 64, 64], 64)  # BASELINE config 1
    run_ref_chain("linear100", 100, "linear", [64, 50, 33, 64], 64)  # well-conditioned, ragged
    save("mini_chain.npz", **chain)


if __name__ == "__main__":
    main()
