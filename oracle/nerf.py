"""
ORACLE (test infrastructure) - CPU restatement of the reference's NeRF backbone builder.

Restates /root/reference/foldingdiff/nerf.py: `NERFBuilder.cartesian_coords` (:79-122),
`centered_cartesian_coords` (:124-128), `place_dihedral` (:145-204, numpy branch) as driven by
`angles_and_coords.create_new_chain_nerf` (/root/reference/foldingdiff/angles_and_coords.py:112-184):
residue i+1's N, CA, C are placed from (psi_i, omega_i, phi_{i+1}) with bond angles
(CA:C:1N)_i, (C:1N:1CA)_i, tau_i - the reference indexes ALL three per-residue bond angles with i, kept as is.
Arithmetic follows the reference's mixed precision: angles arrive as float32 (DataFrame columns), so
cos / sin are float32; coordinates and the frame algebra are float64.
Pinned against the reference module itself (it imports as-is) in tests/golden/make_golden_nerf.py.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

N_CA_LENGTH, CA_C_LENGTH, C_N_LENGTH = 1.46, 1.54, 1.34  # nerf.py:17-19
N_INIT = np.array([17.047, 14.099, 3.625])  # nerf.py:22-24 (first residue of 1CRN)
CA_INIT = np.array([16.967, 12.784, 4.338])
C_INIT = np.array([15.685, 12.755, 5.133])
DEFAULT_ANGLES = {"tau": 109 / 180 * np.pi, "CA:C:1N": 115 / 180 * np.pi, "C:1N:1CA": 121 / 180 * np.pi}  # nerf.py:40-42


def place(a, b, c, bond_angle, bond_length, torsion):
    """nerf.py:173-188."""
    unit = lambda v: v / np.linalg.norm(v, axis=-1)
    ab = b - a
    bc = unit(c - b)
    n = unit(np.cross(ab, bc))
    nbc = np.cross(n, bc)
    m = np.stack([bc, nbc, n], axis=-1)
    d = np.stack([-bond_length * np.cos(bond_angle),
                  bond_length * np.cos(torsion) * np.sin(bond_angle),
                  bond_length * np.sin(torsion) * np.sin(bond_angle)], axis=a.ndim - 1)
    return m.dot(d) + c


def build_chain(angles: np.ndarray, names: Sequence[str], center: bool = True) -> np.ndarray:
    """
    angles (L, F) float32 with columns `names` (must contain phi, psi, omega; tau / CA:C:1N / C:1N:1CA optional)
    -> (3L, 3) float64 coordinates in N, CA, C order, like create_new_chain_nerf builds them.
    """
    col = {n: angles[:, i] for i, n in enumerate(names)}
    L = angles.shape[0]
    get = lambda key, i: (col[key][i] if key in col else DEFAULT_ANGLES[key])
    coords = [N_INIT.copy(), CA_INIT.copy(), C_INIT.copy()]
    for i in range(L - 1):
        dih = (col["psi"][i], col["omega"][i], col["phi"][i + 1])
        steps = (("CA:C:1N", C_N_LENGTH), ("C:1N:1CA", N_CA_LENGTH), ("tau", CA_C_LENGTH))
        for j, (akey, blen) in enumerate(steps):
            coords.append(place(coords[-3], coords[-2], coords[-1], get(akey, i), blen, dih[j]))
    out = np.array(coords)
    return out - out.mean(axis=0) if center else out
