"""
ORACLE (test infrastructure) - what the reference writes for the sampler's outputs, restated on the CPU.

  * angle tables: the reference's own call, `DataFrame.to_csv(path)` on a float32 frame
    (/root/reference/bin/sample.py:360-370).  pandas is a dependency of the reference and is installed here, so this
    half of the oracle IS the reference behaviour (pinned).
  * PDB files: the reference goes through biotite (`struc.Atom(...)` -> bond list -> `PDBFile.set_structure` -> `write`,
    /root/reference/foldingdiff/angles_and_coords.py:187-253), which is not installed here.  PINNED instead on the
    file the reference's own writer produced and committed, plots/pdb_structures/noising_visualization/fully_noised.pdb
    (copied to tests/golden/ref_fully_noised.pdb): 80-column ATOM records (GLY, chain A, res_id from 1, atom_id from 1,
    occupancy 1.00, b_factor 5.00, elements N / C / C) followed by CONECT records for the inter-residue bonds only
    (C_i - N_{i+1}, both directions); no TER / END.  (clean.pdb in the same folder was written by an older revision of
    the function - res_id from 0, b_factor 0.00, no bonds - and is not a fixture.)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import io
from typing import Sequence

import numpy as np
import pandas as pd


def angles_csv_text(angles: np.ndarray, feature_names: Sequence[str]) -> str:
    buf = io.StringIO()
    pd.DataFrame(np.asarray(angles, dtype=np.float32), columns=list(feature_names)).to_csv(buf)
    return buf.getvalue()


def backbone_pdb_text(coords: np.ndarray) -> str:
    coords = np.asarray(coords, dtype=np.float32)
    assert coords.ndim == 2 and coords.shape[1] == 3 and len(coords) % 3 == 0
    names, elems = (" N  ", " CA ", " C  "), (" N", " C", " C")
    lines = []
    for i, (x, y, z) in enumerate(coords.astype(np.float64)):
        lines.append("ATOM  %5d %s %3s %s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s  \n"
                     % (i + 1, names[i % 3], "GLY", "A", i // 3 + 1, x, y, z, 1.0, 5.0, elems[i % 3]))
    for c in range(3, len(coords), 3):
        lines.append("CONECT%5d%5d\n" % (c, c + 1))
        lines.append("CONECT%5d%5d\n" % (c + 1, c))
    return "".join(lines)
