"""
The algebra behind the LayerNorm folded into the projections (foldingdiff_b200/csrc/gemm_tc.cuh: TcLn), restated in numpy
float64 - what the CUDA epilogues implement, independent of them:

    LN(v) W^T + b  ==  rstd * (v W'^T - mean * c) + d,     W' = W diag(gamma),  c = W gamma,  d = W beta + b

with mean / rstd from the partial sums (sum v, sum v^2) the producing epilogue writes, and the residual of the next site
being LN(v) recomputed from v and the same sums.  Also the accumulation de-bias factors (tc_split_weight_kernel): a per-K-chunk
scale of the weights is undone exactly by no other operand, i.e. it is a pure (tiny) reweighting of the chunk products.
"""
import numpy as np


def layer_norm(v, g, b, eps):
    m = v.mean(-1, keepdims=True)
    var = ((v - m) ** 2).mean(-1, keepdims=True)
    return (v - m) / np.sqrt(var + eps) * g + b


def test_consumer_identity_and_partial_sum_statistics():
    rng = np.random.default_rng(0)
    rows, H, N, eps = 37, 384, 1152, 1e-12
    v = rng.normal(0.3, 1.7, (rows, H))
    g, be = 1 + 0.1 * rng.normal(size=H), 0.1 * rng.normal(size=H)
    W, b = 0.02 * rng.normal(size=(N, H)), 0.02 * rng.normal(size=N)
    ref = layer_norm(v, g, be, eps) @ W.T + b
    # what tc_fold_vectors_kernel / tc_split_weight_kernel prepare at create time
    Wp, c, d = W * g[None, :], W @ g, W @ be + b
    # what the producing epilogue leaves: 4 partial (sum, sum of squares) per row (2 column blocks x 2 column groups)
    parts = np.stack([np.stack([v[:, 96 * p:96 * p + 96].sum(-1), (v[:, 96 * p:96 * p + 96] ** 2).sum(-1)], -1) for p in range(4)], 1)
    s1, s2 = parts[..., 0].sum(1), parts[..., 1].sum(1)
    mean = s1 / H
    rstd = 1.0 / np.sqrt(np.maximum(s2 / H - mean * mean, 0.0) + eps)
    got = rstd[:, None] * (v @ Wp.T - mean[:, None] * c[None, :]) + d[None, :]
    assert np.abs(got - ref).max() < 1e-11
    # the residual of the next LayerNorm site: LN(v) recomputed from the raw rows and the same statistics
    res = ((v - mean[:, None]) * rstd[:, None]) * g + be
    assert np.abs(res - layer_norm(v, g, be, eps)).max() < 1e-12


def test_query_row_scale_carries_into_both_fold_vectors():
    """The query rows of the fused QKV weight are scaled by s_n (attention de-bias): W', c and d all carry s_n, so the
    epilogue's rstd * (acc - mean c) + d is s_n times the unscaled query - exactly what scaling Q means."""
    rng = np.random.default_rng(1)
    H = 64
    v = rng.normal(size=(5, H))
    g, be = 1 + 0.1 * rng.normal(size=H), 0.1 * rng.normal(size=H)
    W, b = rng.normal(size=(8, H)), rng.normal(size=8)
    s = 1 + 1e-7 * np.arange(1, 9)
    ref = (layer_norm(v, g, be, 1e-12) @ W.T + b) * s[None, :]
    m = v.mean(-1)
    r = 1 / np.sqrt(v.var(-1) + 1e-12)
    Wp, c, d = W * g[None, :] * s[:, None], s * (W @ g), s * (W @ be + b)
    got = r[:, None] * (v @ Wp.T - m[:, None] * c[None, :]) + d[None, :]
    assert np.abs(got - ref).max() < 1e-11


def test_chunk_prescale_is_a_linear_reweighting_of_the_chunk_products():
    """tc_rz: scaling K chunk i of the weights by 1 + beta (nk - i) changes the product by beta * sum_i (nk - i) p_i - the very
    functional a per-step relative shrink eps of the running sum removes (sum_j acc_j = sum_i (nk - i) p_i for one
    accumulate per chunk), so with beta = eps the two cancel to first order."""
    rng = np.random.default_rng(2)
    K, nk, beta = 384, 24, 1.0e-7
    a, w = rng.normal(size=K), rng.normal(size=K)
    p = (a * w).reshape(nk, 16).sum(-1)                       # per-chunk products
    scale = 1 + beta * (nk - np.arange(nk))
    pres = (a * (w.reshape(nk, 16) * scale[:, None]).reshape(-1)).sum()
    assert abs(pres - (p.sum() + beta * ((nk - np.arange(nk)) * p).sum())) < 1e-12
    acc, lost = 0.0, 0.0
    for i in range(nk):                                       # a running sum that loses eps of itself at every accumulate
        acc += p[i]
        lost += beta * acc
    assert abs(lost - beta * ((nk - np.arange(nk)) * p).sum()) < 1e-12
