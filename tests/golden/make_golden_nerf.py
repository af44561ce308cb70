#!/usr/bin/env python
"""Golden vectors for the NeRF row, written by the REFERENCE's own nerf.py.

    python tests/golden/make_golden_nerf.py      # authoring container only (/root/reference)

nerf.npz: float32 angle sets (canonical-full-angles column order) and the float64 coordinates that
foldingdiff.nerf.NERFBuilder produces for them, centred and uncentred, exactly the way
angles_and_coords.create_new_chain_nerf calls it (angles_and_coords.py:140-166).
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nerf as onerf  # noqa: E402
from oracle import ref_shims  # noqa: E402

ref_shims.install()
from foldingdiff import nerf as rnerf  # noqa: E402  (imports as-is: numpy + torch only)

NAMES = ["phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA"]


def reference_build(df: pd.DataFrame, center: bool) -> np.ndarray:
    b = rnerf.NERFBuilder(phi_dihedrals=df["phi"], psi_dihedrals=df["psi"], omega_dihedrals=df["omega"],
                          bond_angle_ca_c=df["tau"], bond_angle_c_n=df["CA:C:1N"], bond_angle_n_ca=df["C:1N:1CA"])
    return np.asarray(b.centered_cartesian_coords if center else b.cartesian_coords)


def main():
    rng = np.random.default_rng(20260923)
    out = {}
    # (a) realistic angles: helix/strand-like dihedrals + jitter, bond angles near their ideal values
    for tag, L in (("a", 128), ("b", 50), ("c", 77), ("d", 2)):
        ang = np.zeros((L, 6), dtype=np.float32)
        ang[:, 0] = rng.normal(-1.2, 0.4, L); ang[:, 1] = rng.normal(-0.7, 0.6, L)
        ang[:, 2] = rng.normal(np.pi, 0.05, L); ang[:, 2] = (ang[:, 2] + np.pi) % (2 * np.pi) - np.pi
        ang[:, 3] = rng.normal(1.94, 0.05, L); ang[:, 4] = rng.normal(2.03, 0.03, L); ang[:, 5] = rng.normal(2.12, 0.03, L)
        df = pd.DataFrame(ang, columns=NAMES)
        out[f"{tag}_angles"] = ang
        for center in (True, False):
            ref = reference_build(df, center)
            mine = onerf.build_chain(ang, NAMES, center=center)
            assert ref.shape == (3 * L, 3) and np.array_equal(ref, mine), (tag, center, np.abs(ref - mine).max())
            out[f"{tag}_coords_{'centered' if center else 'raw'}"] = ref
    # (b) fully random wrapped angles (what an untrained sampler emits)
    ang = ((rng.normal(0, 1.5, (64, 6)) + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    ang[:, 3:] = np.abs(ang[:, 3:]) * 0.3 + 1.6
    out["r_angles"] = ang
    ref = reference_build(pd.DataFrame(ang, columns=NAMES), True)
    assert np.array_equal(ref, onerf.build_chain(ang, NAMES, center=True))
    out["r_coords_centered"] = ref
    np.savez(os.path.join(HERE, "nerf.npz"), **out)
    print("wrote nerf.npz; oracle == reference bit for bit on", [k for k in out if k.endswith('angles')])


if __name__ == "__main__":
    main()
