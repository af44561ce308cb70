"""bin/sample.py end to end on the mini fixture (with a mean-offset file): flags and output tree of the reference CLI."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, mini_state_dict, write_model_dir
from foldingdiff_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_sample_cli(tmp_path):
    sd, cfg, targs, ckpt = mini_state_dict()
    targs = dict(targs, timesteps=20, variance_schedule="linear")
    mdir = write_model_dir(str(tmp_path / "model"), sd, cfg, targs, ckpt, mean_offset=synthetic.CATH_MEAN_OFFSET)
    out = tmp_path / "out"
    cmd = [sys.executable, os.path.join(ROOT, "bin", "sample.py"), "-m", mdir, "-o", str(out), "-n", "2", "-l", "50", "53",
           "-b", "4", "--seed", "11", "--fullhistory"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    files = sorted(os.listdir(out / "sampled_angles"))
    assert [f for f in files if f.endswith(".csv.gz")] == [f"generated_{i}.csv.gz" for i in range(6)]
    df = pd.read_csv(out / "sampled_angles" / "generated_5.csv.gz", index_col=0)
    assert df.shape == (52, 6) and list(df.columns) == ["phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA"]
    assert np.abs(df.to_numpy()).max() <= np.pi + 1e-6
    assert len(os.listdir(out / "sampled_angles" / "sample_history" / "generated_0")) == 20
    assert sorted(os.listdir(out / "model_snapshot")) == ["config.json", "models", "training_args.json"]
    xyz = np.load(out / "sampled_coords.npz")
    assert xyz["generated_0"].shape == (150, 3) and np.isfinite(xyz["generated_5"]).all()
    # sampled_pdb/generated_{i}.pdb (native writer): ATOM records of the same coordinates, 3 decimals
    assert sorted(os.listdir(out / "sampled_pdb")) == sorted(f"generated_{i}.pdb" for i in range(6))
    lines = open(out / "sampled_pdb" / "generated_5.pdb").read().splitlines()
    lines = [l for l in lines if l.startswith("ATOM")]
    assert len(lines) == 3 * 52 and all(len(l) == 80 for l in lines)
    parsed = np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in lines])
    assert np.abs(parsed - xyz["generated_5"]).max() <= 5.1e-4
    # a non-empty output directory is refused, like the reference (bin/sample.py:299)
    again = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert again.returncode != 0 and "to be empty" in again.stderr


def test_sample_cli_torchrun_two_gpus(tmp_path):
    """torchrun --nproc-per-node 2 bin/sample.py: chains sharded over two GPUs, rank 0 writes the same output tree."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sd, cfg, targs, ckpt = mini_state_dict()
    targs = dict(targs, timesteps=20, variance_schedule="linear")
    mdir = write_model_dir(str(tmp_path / "model"), sd, cfg, targs, ckpt, mean_offset=synthetic.CATH_MEAN_OFFSET)
    out = tmp_path / "out"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "bin", "sample.py"), "-m", mdir, "-o", str(out), "-n", "3",
           "-l", "50", "53", "-b", "5", "--seed", "11"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    files = sorted(f for f in os.listdir(out / "sampled_angles") if f.endswith(".csv.gz"))
    assert len(files) == 9
    lens = [len(pd.read_csv(out / "sampled_angles" / f"generated_{i}.csv.gz", index_col=0)) for i in range(9)]
    assert lens == [50, 50, 50, 51, 51, 51, 52, 52, 52]  # the reference's order survives the round-robin sharding
    vals = np.stack([pd.read_csv(out / "sampled_angles" / f"generated_{i}.csv.gz", index_col=0).to_numpy()[:50] for i in range(9)])
    assert np.isfinite(vals).all() and np.abs(vals).max() <= np.pi + 1e-6
    # nine distinct chains: no two ranks replayed the same noise
    flat = vals.reshape(9, -1)
    assert min(np.abs(flat[i] - flat[j]).max() for i in range(9) for j in range(i)) > 1e-3
    assert sorted(os.listdir(out / "model_snapshot")) == ["config.json", "models", "training_args.json"]
