#!/usr/bin/env python
"""Is the 3-pass tensor-core GEMM error a systematic shrink (RZ accumulate) or noise?  (experiment)"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foldingdiff_b200 import _native
lib = _native.lib()
def run(mode, a, w):
    c = torch.empty(a.shape[0], w.shape[0], device="cuda")
    _native.check(lib.fd_debug_gemm(mode, a.data_ptr(), w.data_ptr(), None, c.data_ptr(), a.shape[0], w.shape[0], a.shape[1], None), "gemm")
    torch.cuda.synchronize()
    return c.cpu().double()
for K in (64, 192, 384, 768):
    g = torch.Generator().manual_seed(K)
    a = torch.randn(1024, K, generator=g); w = torch.randn(384, K, generator=g) * 0.02
    # positive-mean variant: partial sums keep their sign
    a2 = a.abs(); w2 = w.abs()
    for tag, (x, y) in (("zero-mean", (a, w)), ("same-sign", (a2, w2))):
        ref = x.double() @ y.double().T
        for mode, name in ((0, "fp32"), (1, "tc3x")):
            c = run(mode, x.cuda(), y.cuda())
            rel = (c - ref) / ref.abs().clamp_min(ref.abs().median())
            big = ref.abs() > ref.abs().median()
            signed = ((c - ref) * torch.sign(ref))[big] / ref.abs()[big]
            print(f"K={K:4d} {tag:9s} {name}: mean signed rel err {signed.mean():+.3e}  std {signed.std():.3e}  max abs rel {rel.abs().max():.3e}")
