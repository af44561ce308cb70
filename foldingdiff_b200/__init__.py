"""
foldingdiff_b200 - B200-native reverse-diffusion sampler for microsoft/foldingdiff.

One hot path, rebuilt from scratch for sm_100a behind the reference's own Python surface:

    from foldingdiff_b200 import modelling, sampling, datasets
    model = modelling.BertForDiffusionBase.from_dir(model_dir).to("cuda:0")
    dset = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset.from_dir(model_dir), ...)
    angles = sampling.sample(model, dset, n=10, sweep_lengths=(50, 128), batch_size=512)

Modules: `modelling` (loader + forward), `sampling` (p_sample / p_sample_loop / sample),
`beta_schedules`, `datasets` (noise + schedule plumbing), `utils` (angle wrap),
`engine` / `_native` (ctypes binding of include/foldingdiff_b200.h), `distributed`
(chain sharding over GPUs), `synthetic` (seeded production-shape weights for benches).
The CUDA library lives in `csrc/`; there is no CPU or PyTorch-eager fallback.
"""
__version__ = "0.1.0"
