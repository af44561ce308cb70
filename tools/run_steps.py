#!/usr/bin/env python
"""Run a few reverse steps of BASELINE config 2 (for ncu captures; numbers printed here are not bench values)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foldingdiff_b200 import beta_schedules, modelling, synthetic  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--steps", type=int, default=2)
p.add_argument("--warm", type=int, default=1)
p.add_argument("--gemm", default="tc3x")
p.add_argument("--batch", type=int, default=512)
a = p.parse_args()
cfg = modelling.BertConfig(**synthetic.PRODUCTION)
m = modelling.BertForDiffusionBase(cfg, ft_is_angular=[True] * 6, gemm=a.gemm)
m.load_state_dict(synthetic.synthetic_state_dict(synthetic.PRODUCTION, seed=0))
m = m.to("cuda:0")
eng = m.native_engine()
lengths = synthetic.sweep_lengths(a.batch)
T = 1000
eng.set_schedule(beta_schedules.get_variance_schedule("cosine", T), T)
eng.set_batch(lengths, max(lengths))
x = torch.randn(a.batch, max(lengths), 6, device="cuda")
n = a.warm + a.steps
z = torch.randn(n, a.batch, max(lengths), 6, device="cuda")
eng.p_sample_steps(x, T, T - n, z, None, [True] * 6)
torch.cuda.synchronize()
print("done", float(x.abs().max()))
