"""
Output writers of the sampling CLI: angle tables as pandas-compatible csv.gz and
backbones as PDB files, formatted and compressed by host C++ threads behind the C ABI
(`fd_write_angles_csv_gz`, `fd_write_backbone_pdb`, `fd_write_batch`; csrc/writers.hpp).

Mirrors the reference's entry points for this step:
  * `s.to_csv(sampled_angles_folder / f"generated_{i}.csv.gz")`      /root/reference/bin/sample.py:365-370
  * `angles_and_coords.write_coords_to_pdb(coords, out_fname)`       /root/reference/foldingdiff/angles_and_coords.py:187-253
  * `angles_and_coords.create_new_chain_nerf(out_fname, df, ...)`    :112-184   (NeRF on the GPU: foldingdiff_b200.nerf)
  * `write_preds_pdb_folder(final_sampled, outdir, ...)`             /root/reference/bin/sample.py:105-128
The reference pays a pandas / gzip / biotite round trip per chain inside a multiprocessing pool; `write_batch` does the
whole batch in one native call.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _native

GZ_LEVEL = 9  # what pandas' to_csv(..., compression="gzip") uses


def _cstrs(items: Sequence[str]):
    arr = (C.c_char_p * len(items))(*[os.fsencode(str(s)) for s in items])
    return arr


def write_angles_csv_gz(angles: np.ndarray, feature_names: Sequence[str], path, gz_level: int = GZ_LEVEL) -> str:
    """`pd.DataFrame(angles, columns=feature_names).to_csv(path)` for a float32 (rows, F) array and a .csv.gz path."""
    a = np.ascontiguousarray(angles, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == len(feature_names)
    names = _cstrs(feature_names)
    _native.check(_native.lib().fd_write_angles_csv_gz(a.ctypes.data, a.shape[0], a.shape[1], a.shape[1], names,
                                                       os.fsencode(str(path)), gz_level), "fd_write_angles_csv_gz")
    return str(path)


def write_coords_to_pdb(coords: np.ndarray, out_fname: str) -> str:
    """Same name and signature as the reference's angles_and_coords.write_coords_to_pdb: (3N, 3) N/CA/C coordinates."""
    c = np.ascontiguousarray(coords, dtype=np.float32)
    assert c.ndim == 2 and c.shape[1] == 3
    assert len(c) % 3 == 0, f"Expected 3N coords, got {len(c)}"
    _native.check(_native.lib().fd_write_backbone_pdb(c.ctypes.data, len(c), os.fsencode(str(out_fname))), "fd_write_backbone_pdb")
    return out_fname


def write_batch(lengths: Sequence[int], angles: Optional[np.ndarray] = None, feature_names: Optional[Sequence[str]] = None,
                csv_paths: Optional[Sequence[str]] = None, coords: Optional[np.ndarray] = None,
                pdb_paths: Optional[Sequence[str]] = None, threads: Optional[int] = None, gz_level: int = GZ_LEVEL) -> None:
    """
    angles (B, N, F) float32 + csv_paths -> one csv.gz per chain (first lengths[i] rows);
    coords (B, 3N', 3) float32 + pdb_paths -> one PDB per chain (first 3 * lengths[i] atoms).
    """
    lens = np.ascontiguousarray(np.asarray(list(lengths), dtype=np.int32))
    B = len(lens)
    a_ptr, n_pad, F, names, c_ptr, atoms_pad, csv, pdb = None, 0, 0, None, None, 0, None, None
    keep = []
    if csv_paths is not None:
        a = np.ascontiguousarray(angles, dtype=np.float32)
        assert a.ndim == 3 and a.shape[0] == B and len(csv_paths) == B and len(feature_names) == a.shape[2]
        a_ptr, n_pad, F, names, csv = a.ctypes.data, a.shape[1], a.shape[2], _cstrs(feature_names), _cstrs(csv_paths)
        keep.append(a)
    if pdb_paths is not None:
        c = np.ascontiguousarray(coords, dtype=np.float32)
        assert c.ndim == 3 and c.shape[0] == B and c.shape[2] == 3 and len(pdb_paths) == B
        c_ptr, atoms_pad, pdb = c.ctypes.data, c.shape[1], _cstrs(pdb_paths)
        keep.append(c)
    if B == 0 or (csv is None and pdb is None):
        return
    n_threads = threads if threads else min(32, os.cpu_count() or 1)
    _native.check(_native.lib().fd_write_batch(B, a_ptr, n_pad, F, names, c_ptr, atoms_pad, lens.ctypes.data, csv, pdb,
                                               n_threads, gz_level), "fd_write_batch")


def create_new_chain_nerf(out_fname: str, dists_and_angles, angles_to_set: Optional[List[str]] = None,
                          dists_to_set: Optional[List[str]] = None, center_coords: bool = True, device: str = "cuda:0") -> str:
    """
    Reference signature (angles_and_coords.py:112-184): angles DataFrame -> NeRF -> PDB file; returns the path.
    NeRF runs on the GPU (foldingdiff_b200.nerf.build_backbone); bond lengths are the reference's defaults, so a
    non-empty `dists_to_set` is rejected like any other unsupported column.
    """
    import torch

    from . import nerf as fnerf
    cols = list(dists_and_angles.columns)
    if angles_to_set is None and dists_to_set is None:
        angles_to_set = [c for c in cols if c.count(":") != 1]
        dists_to_set = [c for c in cols if c.count(":") == 1]
    assert angles_to_set is not None and dists_to_set is not None
    assert all(a in angles_to_set for a in ("phi", "psi", "omega"))
    if dists_to_set:
        raise NotImplementedError("per-residue bond lengths are not implemented natively (no shipped model predicts them)")
    for a in angles_to_set:
        if a not in ("phi", "psi", "omega", "tau", "N:CA:C", "CA:C:1N", "C:1N:1CA"):
            raise ValueError(f"Unrecognized angle: {a}")
    ang = torch.from_numpy(np.ascontiguousarray(dists_and_angles[angles_to_set].to_numpy(dtype=np.float32)))[None]
    xyz = fnerf.build_backbone(ang.to(device), [ang.shape[1]], angles_to_set, center=center_coords)[0].cpu().numpy()
    return write_coords_to_pdb(xyz, out_fname)


def write_preds_pdb_folder(final_sampled, outdir: str, basename_prefix: str = "generated_", threads: Optional[int] = None,
                           device: str = "cuda:0") -> List[str]:
    """Reference signature (bin/sample.py:105-128): a list of angle DataFrames -> `outdir/generated_{i}.pdb`, batched."""
    import torch

    from . import nerf as fnerf
    os.makedirs(outdir, exist_ok=True)
    if not len(final_sampled):
        return []
    cols = list(final_sampled[0].columns)
    lens = [len(df) for df in final_sampled]
    packed = np.zeros((len(lens), max(lens), len(cols)), dtype=np.float32)
    for i, df in enumerate(final_sampled):
        packed[i, : lens[i]] = df.to_numpy(dtype=np.float32)
    xyz = fnerf.build_backbone(torch.from_numpy(packed).to(device), lens, cols, center=True).cpu().numpy()
    paths = [os.path.join(outdir, f"{basename_prefix}{i}.pdb") for i in range(len(lens))]
    write_batch(lens, coords=xyz, pdb_paths=paths, threads=threads)
    return paths
