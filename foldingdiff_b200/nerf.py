"""
Batched NeRF on the GPU: sampled angles -> backbone N / CA / C coordinates.

The reference builds every chain with a Python loop of 3 (L-1) `place_dihedral` calls
(/root/reference/foldingdiff/nerf.py:79-122) fanned out over a CPU process pool
(/root/reference/bin/sample.py:105-128); once sampling runs at tens of backbones per second that loop is the
end-to-end bottleneck.  Here one CUDA thread walks one chain in fp64
(csrc/nerf.cuh) behind `fd_nerf_build`.

`build_backbone` takes the sampler's `(B, N, F)` angle tensor directly; `nerf_build_batch` keeps the signature of
the reference's batched torch version (nerf.py:207-292) for the fixed-bond-length case.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native

N_CA_LENGTH, CA_C_LENGTH, C_N_LENGTH = 1.46, 1.54, 1.34
_ORDER = ("phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA")
_ALIASES = {"N:CA:C": "tau"}


def build_backbone(angles: torch.Tensor, lengths: Sequence[int], feature_names: Sequence[str],
                   center: bool = True) -> torch.Tensor:
    """
    angles (B, N, F) float32 on a CUDA device, `feature_names` naming the F columns (must include phi, psi, omega;
    tau / CA:C:1N / C:1N:1CA are used when present, else the reference's default bond angles) ->
    (B, 3 N, 3) float32 coordinates, atoms in N, CA, C order, rows beyond 3 * length zero.
    """
    if not angles.is_cuda:
        raise _native.NativeError("build_backbone runs on CUDA tensors only (no CPU fallback)")
    assert angles.dim() == 3 and angles.dtype == torch.float32
    names = [_ALIASES.get(n, n) for n in feature_names]
    assert len(names) == angles.shape[-1]
    cols = np.asarray([names.index(k) if k in names else -1 for k in _ORDER], dtype=np.int32)
    if (cols[:3] < 0).any():
        raise ValueError("phi, psi and omega columns are required")
    B, N, F = angles.shape
    lens = np.ascontiguousarray(np.asarray([int(l) for l in lengths], dtype=np.int32))
    assert lens.shape == (B,)
    x = angles.contiguous()
    out = torch.empty((B, 3 * N, 3), device=angles.device, dtype=torch.float32)
    with torch.cuda.device(angles.device):
        _native.check(_native.lib().fd_nerf_build(x.data_ptr(), B, N, F, lens.ctypes.data, cols.ctypes.data, int(center),
                                                  out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                      "fd_nerf_build")
    return out


def nerf_build_batch(phi: torch.Tensor, psi: torch.Tensor, omega: torch.Tensor, bond_angle_n_ca_c: torch.Tensor,
                     bond_angle_ca_c_n: torch.Tensor, bond_angle_c_n_ca: torch.Tensor,
                     bond_len_n_ca: float = N_CA_LENGTH, bond_len_ca_c: float = CA_C_LENGTH,
                     bond_len_c_n: float = C_N_LENGTH) -> torch.Tensor:
    """(batch, seq) angle tensors -> (batch, seq * 3, 3) uncentred coordinates (reference nerf.py:207-292)."""
    if (bond_len_n_ca, bond_len_ca_c, bond_len_c_n) != (N_CA_LENGTH, CA_C_LENGTH, C_N_LENGTH):
        raise NotImplementedError("only the reference's default bond lengths are implemented natively")
    assert phi.ndim == psi.ndim == omega.ndim == 2 and phi.shape == psi.shape == omega.shape
    ang = torch.stack([phi, psi, omega, bond_angle_n_ca_c, bond_angle_ca_c_n, bond_angle_c_n_ca], dim=-1).to(torch.float32)
    return build_backbone(ang, [phi.shape[1]] * phi.shape[0], _ORDER, center=False)
