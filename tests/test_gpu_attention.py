"""Relative-key attention in isolation (fd_debug_attention): CUDA-core and mma.sync kernels vs torch fp64."""
import math

import numpy as np
import pytest
import torch

from foldingdiff_b200 import _native

pytestmark = pytest.mark.gpu


def reference(qkv, lengths, n_pad, dist, heads, all_rows):
    """fp64 restatement of HF 4.11.3 relative_key attention on packed rows."""
    H = heads * 32
    out, r0 = [], 0
    for n in lengths:
        nr = n_pad if all_rows else n
        blk = qkv[r0:r0 + nr].double()
        q = blk[:, :H].view(nr, heads, 32).permute(1, 0, 2)
        k = blk[:n, H:2 * H].view(n, heads, 32).permute(1, 0, 2)
        v = blk[:n, 2 * H:].view(n, heads, 32).permute(1, 0, 2)
        idx = torch.arange(nr)[:, None] - torch.arange(n)[None, :] + 127
        E = dist.double()[idx]  # (nr, n, 32)
        s = q @ k.transpose(-1, -2) + torch.einsum("hld,lrd->hlr", q, E)
        p = torch.softmax(s / math.sqrt(32), dim=-1)
        out.append((p @ v).permute(1, 0, 2).reshape(nr, H))
        r0 += nr
    return torch.cat(out)


def run(mode, qkv, lengths, n_pad, dist, heads, all_rows):
    lens = np.asarray(lengths, dtype=np.int32)
    rows = qkv.shape[0]
    ctx = torch.zeros(rows, heads * 32, device="cuda")
    _native.check(_native.lib().fd_debug_attention(mode, qkv.data_ptr(), len(lengths), n_pad, lens.ctypes.data,
                                                   int(all_rows), dist.data_ptr(), heads, ctx.data_ptr(), None),
                  "fd_debug_attention")
    return ctx.cpu()


CASES = [([128], 128, False), ([127, 50, 64, 65, 1, 17, 33, 100], 127, False), ([64, 64], 64, False),
         ([50, 128, 16], 128, True), ([80], 80, False), ([113, 97], 128, True),
         # 340 chains x 6 heads = 2040 items on 148 persistent CTAs: multi-item pipelines (ring slots, TMEM reuse, phases)
         ([32] * 40 + [96] * 300, 128, False)]


@pytest.mark.parametrize("lengths,n_pad,all_rows", CASES)
@pytest.mark.parametrize("mode,tol", [(0, 3e-6), (1, 8e-6), (2, 4e-3)])
def test_attention_kernels(lengths, n_pad, all_rows, mode, tol):
    heads = 6
    g = torch.Generator().manual_seed(sum(lengths) + n_pad)
    rows = sum(n_pad if all_rows else l for l in lengths)
    qkv = torch.randn(rows, 3 * heads * 32, generator=g)
    qkv[:, :heads * 32] *= 1.5  # sharper softmax
    dist = torch.randn(255, 32, generator=g) * 0.3
    ref = reference(qkv, lengths, n_pad, dist, heads, all_rows)
    got = run(mode, qkv.cuda(), lengths, n_pad, dist.cuda(), heads, all_rows).double()
    err = float((got - ref).abs().max())
    print(f"attention mode {mode} lengths {lengths[:8]}{'...' if len(lengths) > 8 else ''} all_rows={all_rows}: max abs err {err:.3e}")
    assert err < tol
