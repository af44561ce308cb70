#!/usr/bin/env python
"""
Sample protein backbones with the B200-native sampler.

Same command line and output tree as the reference's bin/sample.py (/root/reference/bin/sample.py:237-287,
README.md:90-96): `-m/--model -o/--outdir -n/--num -l/--lengths -b/--batchsize --fullhistory
--testcomparison --nopsea --seed --device`.  Native here: the hot path (load -> sampling.sample), the batched GPU
NeRF (angles -> N / CA / C coordinates) and the output writers (sampled_angles/*.csv.gz, sampled_pdb/*.pdb,
sampled_coords.npz), SURVEY.md section 8f rows 1-2.  Plots and the PSEA / test-set comparison of the reference need
biotite + matplotlib + its CATH pipeline and are out of scope.
"""
import argparse
import json
import logging
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foldingdiff_b200 import modelling, sampling  # noqa: E402
from foldingdiff_b200.datasets import AnglesEmptyDataset, NoisedAnglesDataset  # noqa: E402

SEED = 7344  # the reference CLI default (bin/sample.py:34-37)


def build_datasets(model_dir: Path):
    """The data-free dataset shell of the reference's build_datasets(load_actual=False) (bin/sample.py:80-103)."""
    with open(model_dir / "training_args.json") as f:
        targs = json.load(f)
    mean_file = model_dir / "training_mean_offset.npy"
    shell = AnglesEmptyDataset(targs["angles_definitions"], pad=targs["max_seq_len"],
                               mean_offset=np.load(mean_file) if mean_file.exists() else None)
    key = "coords" if targs["angles_definitions"] == "cart-coords" else "angles"
    return NoisedAnglesDataset(shell, dset_key=key, timesteps=targs["timesteps"], exhaustive_t=False,
                               beta_schedule=targs["variance_schedule"], nonangular_variance=1.0,
                               angular_variance=targs["variance_scale"])


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(usage=__doc__, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-m", "--model", type=str, default="wukevin/foldingdiff_cath",
                   help="Path to model directory (config.json, training_args.json, models/)")
    p.add_argument("--outdir", "-o", type=str, default=os.getcwd(), help="Path to output directory")
    p.add_argument("--num", "-n", type=int, default=10, help="Number of examples to generate *per length*")
    p.add_argument("-l", "--lengths", type=int, nargs=2, default=[50, 128], help="Range of lengths to sample from")
    p.add_argument("-b", "--batchsize", type=int, default=512, help="Batch size to use when sampling")
    p.add_argument("--fullhistory", action="store_true", help="Store full history, not just final structure")
    p.add_argument("--testcomparison", action="store_true", help="Run comparison against test set (needs the reference's CATH data pipeline)")
    p.add_argument("--nopsea", action="store_true", help="Skip PSEA calculations")
    p.add_argument("--seed", type=int, default=SEED, help="Random seed")
    p.add_argument("--device", type=str, default="cuda:0", help="Device to use (CUDA only: there is no CPU path)")
    return p


def main() -> None:
    args = build_parser().parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    outdir = Path(args.outdir)
    # every argument / output-directory check runs BEFORE the process group exists, so a bad invocation ends the job at
    # once instead of leaving ranks blocked in a collective until the NCCL timeout
    if not os.path.isdir(args.model):
        raise SystemExit(f"{args.model} is not a directory; hub ids need network access, which this build does not assume")
    if args.testcomparison:
        raise SystemExit("--testcomparison needs the reference's CATH dataset pipeline (out of scope here)")
    if world > 1 and args.fullhistory:
        raise SystemExit("--fullhistory is single-GPU only")
    if rank == 0 and os.path.isdir(outdir) and os.listdir(outdir):  # (torchrun tears the other ranks down when rank 0 exits)
        raise SystemExit(f"Expected {outdir} to be empty!")  # the reference asserts the same (bin/sample.py:300)
    if world > 1:
        # launched with torchrun: one process per GPU, chains sharded round-robin over the ranks, one all-gather
        # of the finished angles per batch, rank 0 writes the outputs (foldingdiff_b200/distributed.py)
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        args.device = f"cuda:{local}"
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(args.device))
        rank = dist.get_rank()
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
        os.makedirs(outdir / "plots", exist_ok=True)
    train_dset = build_datasets(Path(args.model))
    model = modelling.BertForDiffusionBase.from_dir(
        args.model, copy_to=outdir / "model_snapshot" if rank == 0 else None).to(torch.device(args.device))
    lo, hi = args.lengths
    assert lo < hi and hi <= train_dset.dset.pad

    if world > 1:
        from foldingdiff_b200 import distributed as fdist
        finals = fdist.sample_sharded(model, train_dset, n=args.num, sweep_lengths=(lo, hi), batch_size=args.batchsize,
                                      seed=args.seed)
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
        sampled = [f[None] for f in finals]
    else:
        torch.manual_seed(args.seed)
        sampled = sampling.sample(model, train_dset, n=args.num, sweep_lengths=(lo, hi), batch_size=args.batchsize,
                                  history="full" if args.fullhistory else "final")
    cols = list(train_dset.feature_names[train_dset.dset_key])
    from foldingdiff_b200 import nerf as fnerf
    from foldingdiff_b200 import writers
    # outputs (reference tree, README.md:90-96): sampled_angles/generated_{i}.csv.gz (what DataFrame.to_csv writes),
    # sampled_pdb/generated_{i}.pdb, plus sampled_coords.npz.  Formatting, gzip and file IO run in native threads
    # (fd_write_batch) instead of one pandas / biotite round trip per chain in a process pool (bin/sample.py:105-128).
    finals = [np.ascontiguousarray(s[-1], dtype=np.float32) for s in sampled]
    lens = [len(f) for f in finals]
    n_max = max(lens)
    packed = np.zeros((len(finals), n_max, len(cols)), dtype=np.float32)
    for i, f in enumerate(finals):
        packed[i, : lens[i]] = f
    angles_dir, pdb_dir = outdir / "sampled_angles", outdir / "sampled_pdb"
    os.makedirs(angles_dir, exist_ok=True)
    os.makedirs(pdb_dir, exist_ok=True)
    csv_paths = [str(angles_dir / f"generated_{i}.csv.gz") for i in range(len(lens))]
    writers.write_batch(lens, angles=packed, feature_names=cols, csv_paths=csv_paths)
    if train_dset.dset_key == "angles":
        # backbone coordinates on the GPU (fd_nerf_build): (3 * length, 3) N/CA/C per chain.  Like the reference's
        # create_new_chain_nerf (angles_and_coords.py:112-184) a chain whose coordinates are not finite gets no PDB file.
        xyz = fnerf.build_backbone(torch.from_numpy(packed).to(args.device), lens, cols, center=True).cpu().numpy()
        ok = [i for i in range(len(lens)) if np.isfinite(xyz[i, : 3 * lens[i]]).all()]
        if len(ok) < len(lens):
            logging.warning(f"{len(lens) - len(ok)} chains have non-finite coordinates: no PDB written for them")
        writers.write_batch([lens[i] for i in ok], coords=np.ascontiguousarray(xyz[ok]),
                            pdb_paths=[str(pdb_dir / f"generated_{i}.pdb") for i in ok])
        np.savez_compressed(outdir / "sampled_coords.npz", **{f"generated_{i}": xyz[i, : 3 * lens[i]] for i in ok})
    if args.fullhistory:
        hist_dir = angles_dir / "sample_history"
        for i, series in enumerate(sampled):
            d = hist_dir / f"generated_{i}"
            os.makedirs(d, exist_ok=True)
            snaps = np.ascontiguousarray(series, dtype=np.float32)  # (T, len, F)
            writers.write_batch([snaps.shape[1]] * len(snaps), angles=snaps, feature_names=cols,
                                csv_paths=[str(d / f"generated_{i}_timestep_{t}.csv.gz") for t in range(len(snaps))])
    dfs = finals
    logging.info(f"Wrote {len(dfs)} sampled angle sets to {angles_dir}")


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    main()
