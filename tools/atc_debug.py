#!/usr/bin/env python
"""Stage-by-stage check of the tcgen05 attention kernel (attention_tc.cuh) against fp64, on a GPU box.

Runs fd_debug_attention with FOLDINGDIFF_B200_ATT=tc and the debug dump switched on, and compares, per
(chain, head) item:  S raw = Q K^T,  S + relative-key term,  O = P V (unnormalised),  ctx.  The first stage
that disagrees names the broken operand layout.  Dumps and inputs of the first case are saved under
gpurun_out/ so a wrong layout can be analysed offline.  Not a test and not a benchmark.
"""
import math
import os
import sys

os.environ.setdefault("FOLDINGDIFF_B200_ATT", "tc")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foldingdiff_b200 import _native  # noqa: E402

DBG_ROW = 290
lib = _native.lib()


def stages(qkv, lengths, n_pad, dist, heads, all_rows):
    H = heads * 32
    out, r0 = [], 0
    for n in lengths:
        nr = n_pad if all_rows else n
        blk = qkv[r0:r0 + nr].double()
        q = blk[:, :H].view(nr, heads, 32).permute(1, 0, 2)
        k = blk[:n, H:2 * H].view(n, heads, 32).permute(1, 0, 2)
        v = blk[:n, 2 * H:].view(n, heads, 32).permute(1, 0, 2)
        idx = torch.arange(nr)[:, None] - torch.arange(n)[None, :] + 127
        E = dist.double()[idx]
        s_raw = q @ k.transpose(-1, -2)
        s_rel = s_raw + torch.einsum("hld,lrd->hlr", q, E)
        t = s_rel / math.sqrt(32) * math.log2(math.e)
        m = t.max(-1, keepdim=True).values
        p = torch.exp2(t - m)
        o = p @ v
        ctx = (o / p.sum(-1, keepdim=True)).permute(1, 0, 2).reshape(nr, H)
        out.append(dict(s_raw=s_raw, s_rel=s_rel, o=o, m=m.squeeze(-1), sum=p.sum(-1), ctx=ctx, nr=nr, n=n))
        r0 += nr
    return out


def run_case(lengths, n_pad, all_rows, heads=6, save=None):
    g = torch.Generator().manual_seed(sum(lengths) + n_pad)
    rows = sum(n_pad if all_rows else l for l in lengths)
    qkv = torch.randn(rows, 3 * heads * 32, generator=g)
    qkv[:, :heads * 32] *= 1.5
    dist = torch.randn(255, 32, generator=g) * 0.3
    ref = stages(qkv, lengths, n_pad, dist, heads, all_rows)
    items = len(lengths) * heads
    dump = torch.full((items, 128, DBG_ROW), float("nan"), device="cuda")
    ctx = torch.zeros(rows, heads * 32, device="cuda")
    lens = np.asarray(lengths, dtype=np.int32)
    qd, dd = qkv.cuda(), dist.cuda()
    lib.fd_debug_attention_dump(dump.data_ptr())
    rc = lib.fd_debug_attention(1, qd.data_ptr(), len(lengths), n_pad, lens.ctypes.data, int(all_rows), dd.data_ptr(), heads,
                                ctx.data_ptr(), None)
    lib.fd_debug_attention_dump(None)
    status = lib.fd_debug_tc_status()
    torch.cuda.synchronize()
    dump, ctx = dump.cpu().double(), ctx.cpu().double()
    errs = dict(s_raw=0.0, s_rel=0.0, o=0.0, m=0.0, sum=0.0, ctx=0.0)
    r0 = 0
    for b, st in enumerate(ref):
        nr, n = st["nr"], st["n"]
        for h in range(heads):
            d = dump[b * heads + h, :nr]
            errs["s_raw"] = max(errs["s_raw"], float((d[:, :n] - st["s_raw"][h]).abs().max()))
            errs["s_rel"] = max(errs["s_rel"], float((d[:, 128:128 + n] - st["s_rel"][h]).abs().max()))
            # O is unnormalised with the kernel's own row max: compare after normalising by the kernel's sum
            errs["o"] = max(errs["o"], float((d[:, 256:288] / d[:, 289:290] - st["o"][h] / st["sum"][h][:, None]).abs().max()))
            errs["m"] = max(errs["m"], float((d[:, 288] - st["m"][h]).abs().max()))
            errs["sum"] = max(errs["sum"], float((d[:, 289] / st["sum"][h] - 1).abs().max()))
        errs["ctx"] = max(errs["ctx"], float((ctx[r0:r0 + nr] - st["ctx"]).abs().max()))
        r0 += nr
    print(f"case {lengths} n_pad={n_pad} all_rows={all_rows}: rc={rc} tc_status={status} " +
          " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
    if save:
        np.savez_compressed(save, qkv=qkv.numpy(), dist=dist.numpy(), dump=dump.float().numpy(), ctx=ctx.float().numpy(),
                            lengths=lens, n_pad=n_pad, all_rows=all_rows)
    return rc == 0 and status == 0 and errs["ctx"] < 1e-5


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    ok = run_case([128], 128, False, save="gpurun_out/atc_case0.npz")
    for lengths, n_pad, all_rows in [([64, 64], 64, False), ([80], 80, False), ([127, 50, 64, 65, 1, 17, 33, 100], 127, False),
                                     ([50, 128, 16], 128, True), ([113, 97], 128, True), ([32] * 40 + [96] * 300, 128, False)]:
        ok = run_case(lengths, n_pad, all_rows) and ok
    print("ATC_DEBUG", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
