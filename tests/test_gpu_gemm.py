"""The projection GEMMs through fd_debug_gemm: CUDA-core fp32, tcgen05 3-pass, tcgen05 1-pass."""
import pytest
import torch

from foldingdiff_b200 import _native

pytestmark = pytest.mark.gpu

# rows % 256 == 0 -> 2-CTA multicast clusters; otherwise the single-CTA variant
SHAPES = [(256, 384, 384), (384, 1152, 384), (128, 768, 384), (256, 384, 768), (128, 576, 192), (512, 192, 384),
          (128, 64, 64), (1280, 1152, 384), (2560, 768, 384), (37888, 384, 768)]


def run(mode, a, w, bias):
    lib = _native.lib()
    c = torch.empty(a.shape[0], w.shape[0], device="cuda")
    _native.check(lib.fd_debug_gemm(mode, a.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(),
                                    c.data_ptr(), a.shape[0], w.shape[0], a.shape[1],
                                    torch.cuda.current_stream().cuda_stream), "fd_debug_gemm")
    torch.cuda.synchronize()
    return c


def make(rows, n, k, seed=0, wscale=0.02):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(rows, k, generator=g)
    w = torch.randn(n, k, generator=g) * wscale
    b = torch.randn(n, generator=g) * 0.02
    ref = a.double() @ w.double().T + b.double()
    return a.cuda(), w.cuda(), b.cuda(), ref


@pytest.mark.parametrize("rows,n,k", SHAPES)
def test_simt_fp32_gemm(rows, n, k):
    a, w, b, ref = make(rows, n, k)
    c = run(_native.GEMM_FP32_SIMT, a, w, b).cpu().double()
    assert float((c - ref).abs().max() / ref.abs().max()) < 2e-6


@pytest.mark.parametrize("rows,n,k", SHAPES)
def test_tc_3x_gemm(rows, n, k):
    a, w, b, ref = make(rows, n, k, seed=1)
    c = run(_native.GEMM_TC_3X, a, w, b).cpu().double()
    assert _native.lib().fd_debug_tc_status() == 0
    err = float((c - ref).abs().max() / ref.abs().max())
    print(f"tc3x {rows}x{n}x{k}: rel err {err:.3e}")
    assert err < 5e-6  # fp16 hi/lo split: ~2^-21; plain fp16 would be ~5e-4


@pytest.mark.parametrize("rows,n,k", SHAPES[:4])
def test_tc_1x_gemm(rows, n, k):
    a, w, b, ref = make(rows, n, k, seed=2)
    c = run(_native.GEMM_TC_1X, a, w, b).cpu().double()
    assert _native.lib().fd_debug_tc_status() == 0
    err = float((c - ref).abs().max() / ref.abs().max())
    print(f"tc1x {rows}x{n}x{k}: rel err {err:.3e}")
    assert err < 3e-3


def test_tc_3x_weight_scaling_extremes():
    for wscale in (1e-4, 0.02, 3.0):  # the per-matrix power-of-two scale keeps `lo` out of fp16 subnormals
        a, w, b, ref = make(256, 384, 384, seed=3, wscale=wscale)
        c = run(_native.GEMM_TC_3X, a, w, b).cpu().double()
        assert float((c - ref).abs().max() / ref.abs().max()) < 5e-6
    assert _native.lib().fd_debug_tc_status() == 0


def test_tc_matches_simt_on_a_large_problem():
    a, w, b, _ = make(128 * 40, 1152, 384, seed=4)
    c0 = run(_native.GEMM_FP32_SIMT, a, w, b)
    c1 = run(_native.GEMM_TC_3X, a, w, b)
    assert _native.lib().fd_debug_tc_status() == 0
    assert float((c0 - c1).abs().max()) < 2e-5
