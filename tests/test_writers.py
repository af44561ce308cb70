"""Output writers: native csv.gz / PDB writers against the oracle (pandas = the reference's own call)."""
import gzip
import io
import os

import numpy as np
import pandas as pd
import pytest

from foldingdiff_b200 import writers
from oracle import writers as owriters

NAMES = ["phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA"]


def _read_gz(path):
    with gzip.open(path, "rt") as f:
        return f.read()


def test_csv_gz_matches_pandas_text_on_sampler_like_values(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.uniform(-np.pi, np.pi, size=(128, 6)).astype(np.float32)
    a[:, 3:] = rng.uniform(1.8, 2.3, size=(128, 3)).astype(np.float32)
    p = tmp_path / "generated_0.csv.gz"
    writers.write_angles_csv_gz(a, NAMES, p)
    assert _read_gz(p) == owriters.angles_csv_text(a, NAMES)
    back = pd.read_csv(p, index_col=0)
    assert list(back.columns) == NAMES and np.array_equal(back.to_numpy(dtype=np.float32), a)


def test_csv_number_formatting_edge_cases(tmp_path):
    """numpy str(float32): positional for 1e-4 <= |x| < 1e6, scientific outside, shortest round-trip digits."""
    specials = [0.0, -0.0, 1.0, -1.0, 0.1, 1e-4, 9.9e-5, 0.0001001, 1e-5, 3e-39, 1e-45, 999999.94, 999999.0, 1e6, 1e7,
                1.5e7, 16777216.0, 1e10, 3.4e38, 1 / 3, 123456.78, 100000.0, np.pi, -np.pi, np.inf, -np.inf, np.nan, 2.5]
    rng = np.random.default_rng(1)
    logu = (10.0 ** rng.uniform(-12, 12, size=5000) * rng.choice([-1, 1], size=5000)).astype(np.float32)
    bits = rng.integers(0, 2**32, size=5000, dtype=np.uint64).astype(np.uint32).view(np.float32)  # any bit pattern
    vals = np.concatenate([np.asarray(specials, dtype=np.float32), logu, bits])
    vals = np.concatenate([vals, np.zeros((-len(vals)) % 6, dtype=np.float32)]).reshape(-1, 6)
    p = tmp_path / "edge.csv.gz"
    writers.write_angles_csv_gz(vals, NAMES, p)
    got, want = _read_gz(p).splitlines(), owriters.angles_csv_text(vals, NAMES).splitlines()
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]


def test_csv_empty_and_ragged(tmp_path):
    p = tmp_path / "empty.csv.gz"
    writers.write_angles_csv_gz(np.zeros((0, 6), dtype=np.float32), NAMES, p)
    assert _read_gz(p) == owriters.angles_csv_text(np.zeros((0, 6), dtype=np.float32), NAMES)
    with pytest.raises(AssertionError):
        writers.write_angles_csv_gz(np.zeros((4, 5), dtype=np.float32), NAMES, p)


def test_pdb_matches_oracle_and_parses_back(tmp_path):
    rng = np.random.default_rng(2)
    xyz = (rng.normal(size=(3 * 57, 3)) * 20).astype(np.float32)
    xyz[0] = [-999.9994, 9999.9994, 0.0005]  # field-width edges of %8.3f
    p = str(tmp_path / "generated_0.pdb")
    assert writers.write_coords_to_pdb(xyz, p) == p
    text = open(p).read()
    assert text == owriters.backbone_pdb_text(xyz)
    lines = [l for l in text.splitlines() if l.startswith("ATOM")]
    assert len(lines) == 171 and all(len(l) == 80 for l in lines)
    assert text.splitlines()[171:173] == ["CONECT    3    4", "CONECT    4    3"] and len(text.splitlines()) == 171 + 2 * 56
    parsed = np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in lines])
    assert np.abs(parsed - xyz).max() <= 5.1e-4
    assert [l[12:16] for l in lines[:3]] == [" N  ", " CA ", " C  "] and lines[3][22:26] == "   2" and lines[170][6:11] == "  171"
    with pytest.raises(AssertionError):
        writers.write_coords_to_pdb(xyz[:5], p)


def test_batch_writer_equals_per_chain_calls(tmp_path):
    rng = np.random.default_rng(3)
    lengths = [50, 128, 1, 77, 64, 99, 3]
    ang = rng.uniform(-np.pi, np.pi, size=(len(lengths), 128, 6)).astype(np.float32)
    xyz = (rng.normal(size=(len(lengths), 3 * 128, 3)) * 15).astype(np.float32)
    csvs = [str(tmp_path / f"generated_{i}.csv.gz") for i in range(len(lengths))]
    pdbs = [str(tmp_path / f"generated_{i}.pdb") for i in range(len(lengths))]
    writers.write_batch(lengths, angles=ang, feature_names=NAMES, csv_paths=csvs, coords=xyz, pdb_paths=pdbs, threads=4)
    for i, n in enumerate(lengths):
        assert _read_gz(csvs[i]) == owriters.angles_csv_text(ang[i, :n], NAMES)
        assert open(pdbs[i]).read() == owriters.backbone_pdb_text(xyz[i, : 3 * n])
    # only one kind of output, single thread
    only = [str(tmp_path / f"only_{i}.pdb") for i in range(len(lengths))]
    writers.write_batch(lengths, coords=xyz, pdb_paths=only, threads=1)
    assert all(open(a).read() == open(b).read() for a, b in zip(only, pdbs))
    with pytest.raises(Exception):
        writers.write_batch([200], angles=ang[:1], feature_names=NAMES, csv_paths=[csvs[0]])  # longer than the padded array
    with pytest.raises(Exception):
        writers.write_batch([5], coords=xyz[:1], pdb_paths=[str(tmp_path / "no_such_dir" / "x.pdb")])


def _atoms(path):
    lines = open(path).read().splitlines()
    atoms = [l for l in lines if l.startswith("ATOM")]
    return np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in atoms], dtype=np.float32), lines


def test_pdb_pinned_on_files_written_by_the_reference(tmp_path):
    """tests/golden/ref_fully_noised.pdb was written by the reference's own write_coords_to_pdb (biotite) and is
    committed in its repository (plots/pdb_structures/noising_visualization/fully_noised.pdb; res_id from 1 and
    b_factor 5.00 as in angles_and_coords.py:200-232 - the sibling clean.pdb predates that code and is not used):
    feeding its coordinates back through the native writer and the oracle must reproduce it byte for byte."""
    from conftest import ROOT
    gold = os.path.join(ROOT, "tests", "golden", "ref_fully_noised.pdb")
    xyz, _ = _atoms(gold)
    assert xyz.shape == (264, 3)
    p = str(tmp_path / "a.pdb")
    writers.write_coords_to_pdb(xyz, p)
    assert open(p).read() == open(gold).read() == owriters.backbone_pdb_text(xyz)
