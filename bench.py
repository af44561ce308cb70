#!/usr/bin/env python
"""
bench.py - backbones/sec of the reverse-diffusion sampler (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus 1 ...             # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1

One "step" = one pass of the hot path over one batch: the full T-step p_sample loop (T = 1000,
cosine schedule) on 512 chains of lengths 50..127 (BASELINE config 2, "foldingdiff_cath BERT,
batch=512, len 50-128, T=1000") with synthetic weights of the production architecture (the real
checkpoint is not available offline) and synthetic wrapped-Gaussian noise.

  value  whole-job backbones/s with the initial noise already resident in HBM
  e2e    the same through the public API with HOST buffers: sampling.p_sample_loop(host noise) ->
         host tensor; H2D of the noise and D2H of the result inside the timed region
  roofline / kernels  per-kernel device times measured live with CUDA events (fd_profile_*)
  cpu_baseline  the oracle port (torch CPU fp32 restatement + the reference's loop arithmetic)
         timed on this box's host cores on a bounded sample, extrapolated to T steps
Multi-GPU: chains are independent -> each rank runs 512 chains (weak scaling), no collective inside
the loop, one NCCL all-gather of the final angles per pass (inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from foldingdiff_b200 import beta_schedules, datasets, synthetic  # noqa: E402

SEED = 7344
PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")
# From the committed `ncu --set full` capture of this configuration (profiles/r01_gemm_ncu.md, config 2,
# FD_GEMM_TC_3X, CTA-pair mode): dram__bytes_read.sum + dram__bytes_write.sum per launch, and
# sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active.  Static annotations, not re-measured here.
NCU_GEMM = {
    "dram_bytes_per_launch": {"gemm_qkv": 226.2e6, "gemm_attn_out": 82.4e6, "gemm_ffn1": 155.6e6, "gemm_ffn2": 153.1e6},
    "tensor_pipe_active_pct": {"gemm_qkv": 75.8, "gemm_attn_out": 66.2, "gemm_ffn1": 62.1, "gemm_ffn2": 79.5,
                               "attention_tc": 18.6},
    "source": "profiles/r01_gemm_ncu.md, profiles/r01_attention_tc_ncu.md",
}
FALLBACK_PEAK_TFLOPS = 1590.0  # /opt/skills/guides/B200_PROFILING.md fallback (burst)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--workload", choices=["config2", "config3", "config5"], default="config2")
    p.add_argument("--timesteps", type=int, default=1000)
    p.add_argument("--batch", type=int, default=None, help="chains per GPU (default: the workload's)")
    p.add_argument("--gemm", choices=["tc3x", "fp32", "tc1x"], default=os.environ.get("FOLDINGDIFF_B200_GEMM", "tc3x"))
    p.add_argument("--e2e-history", choices=["full", "final"], default="full")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-chains", type=int, default=64)
    p.add_argument("--cpu-steps", type=int, default=3)
    p.add_argument("--profile-steps", type=int, default=20, help="reverse steps of the CUDA-event kernel profile")
    return p.parse_args()


def workload(args):
    """-> (lengths per GPU, n_pad, start_t, wrap_all, name)"""
    T = args.timesteps
    if args.workload == "config2":
        B = args.batch or 512
        lengths = synthetic.sweep_lengths(B)
        return lengths, max(lengths), T, False, f"foldingdiff_cath BERT (synthetic weights), batch={B}/GPU, len 50-127, T={T}"
    if args.workload == "config3":
        B = args.batch or 4096
        return [128] * B, 128, T, False, f"foldingdiff_cath (synthetic weights), batch={B}/GPU, len=128, T={T}"
    B = args.batch or 512
    return [128] * B, 128, min(250, T), True, f"partial_noise_reconstruct (synthetic weights/inputs), batch={B}/GPU, len=128, from t={min(250, T)}"


def initial_noise(B, n_pad, T):
    d = datasets.NoisedAnglesDataset(datasets.AnglesEmptyDataset("canonical-full-angles", pad=128),
                                     timesteps=T, beta_schedule="cosine")
    torch.manual_seed(SEED)
    return d, d.sample_noise(torch.zeros(B, 128, 6))[:, :n_pad].contiguous()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
_CPU_MODEL = None
_REF = None  # (sampling, beta_schedules, utils) of the reference installed under baseline/_ref, or False


def _reference():
    """The UNMODIFIED reference from baseline/_ref (baseline/reference_arm.py), or False where it is not installed."""
    global _REF
    if _REF is None:
        try:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import reference_arm
            _REF = reference_arm.load_reference()[:3] + (reference_arm,) if reference_arm.available() else False
        except Exception as e:  # noqa: BLE001
            print(f"[bench] baseline/_ref not usable ({e}); timing the oracle port instead", file=sys.stderr)
            _REF = False
    return _REF


def cpu_kind():
    return "reference" if _reference() else "port"


def _cpu_model():
    global _CPU_MODEL
    if _CPU_MODEL is None:
        sd = synthetic.synthetic_state_dict(synthetic.PRODUCTION, seed=0)
        ref = _reference()
        if ref:
            _CPU_MODEL = ref[3].build_model(sd, synthetic.PRODUCTION)
        else:
            from oracle import forward as ofwd  # the one place bench.py may execute oracle/
            _CPU_MODEL = ofwd.OracleModel(sd, ofwd.OracleConfig(**synthetic.PRODUCTION), [True] * 6).eval()
    return _CPU_MODEL


def _cpu_steps(sub, n_pad, T, start_t, steps, threads):
    """
    seconds per reverse step of the reference's CPU path on `threads` host threads (1 untimed warm-up step).
    With baseline/_ref: the stock `sampling.p_sample` + per-column `utils.modulo_with_wrapped_range` + `.cpu()` of the
    reference's loop body (sampling.py:111-131) around the reference-assembled model; otherwise the oracle port.
    """
    torch.set_num_threads(threads)
    model = _cpu_model()
    ref = _reference()
    g = torch.Generator().manual_seed(SEED)
    if ref:
        sampling, beta_schedules, utils = ref[:3]
        betas = beta_schedules.get_variance_schedule("cosine", T)
        x = torch.randn(len(sub), n_pad, 6, generator=g)

        def step(img, i):
            with torch.no_grad():
                img = sampling.p_sample(model=model, x=img, t=torch.full((len(sub),), i, dtype=torch.long), seq_lens=sub,
                                        t_index=i, betas=betas)
            for j in range(img.shape[2]):
                img[:, :, j] = utils.modulo_with_wrapped_range(img[:, :, j], range_min=-torch.pi, range_max=torch.pi)
            img.cpu()
            return img
        x = step(x, start_t - 1)
        t0 = time.perf_counter()
        for k in range(steps):  # per-step cost does not depend on t
            x = step(x, start_t - 2 - k)
        return (time.perf_counter() - t0) / steps
    from oracle import loop as oloop
    from oracle import schedules as osched
    betas = osched.betas_for("cosine", T)
    x = oloop.wrap(torch.randn(len(sub), n_pad, 6, generator=g))
    x = oloop.wrap(oloop.p_sample(model, x, torch.full((len(sub),), start_t - 1, dtype=torch.long), sub, betas))
    t0 = time.perf_counter()
    for k in range(steps):  # per-step cost does not depend on t
        x = oloop.wrap(oloop.p_sample(model, x, torch.full((len(sub),), start_t - 2 - k, dtype=torch.long), sub, betas))
    return (time.perf_counter() - t0) / steps


_BEST_THREADS = None


def cpu_reference_rate(lengths, n_pad, T, chains, steps, start_t, wrap_all):
    """
    backbones/s of the reference's CPU path from a bounded sample: `chains` chains taken evenly from the
    workload's length mix, `steps` reverse steps timed after one warm-up, extrapolated to the full loop.
    torch's CPU kernels do not scale to every core count, so the thread count is chosen by a one-step
    probe over {all cores, 64, 32, 16} and the best one is used ("all the host threads it can use").
    """
    global _BEST_THREADS, _REF, _CPU_MODEL
    cores = os.cpu_count() or 1
    stride = max(1, len(lengths) // chains)
    sub = [lengths[i] for i in range(0, len(lengths), stride)][:chains]
    if _reference():
        try:  # a reference install that imports but does not run must not cost the measurement: fall back to the port
            _cpu_steps(sub[:2], n_pad, T, start_t, 1, min(cores, 16))
        except Exception as e:  # noqa: BLE001
            print(f"[bench] reference arm failed ({type(e).__name__}: {e}); timing the oracle port instead", file=sys.stderr)
            _REF, _CPU_MODEL = False, None
    if _BEST_THREADS is None:
        cands = sorted({c for c in (cores, 64, 32, 16) if c <= cores}, reverse=True)
        probe = {c: _cpu_steps(sub, n_pad, T, start_t, 1, c) for c in cands}
        _BEST_THREADS = min(probe, key=probe.get)
    dt = _cpu_steps(sub, n_pad, T, start_t, steps, _BEST_THREADS)
    rate = len(sub) / (dt * start_t)
    what = ("stock sampling.p_sample + wrap of the reference installed in baseline/_ref around its own GaussianFourierProjection / "
            "BertEmbeddings / AnglesPredictor and the installed transformers' relative_key attention + BERT blocks (4.11.3 is "
            "not installable); " if _reference() else "oracle port of the reference loop and forward (baseline/_ref absent); ")
    sample = (what + f"{len(sub)} chains (every {stride}-th of the workload's lengths, sum len {sum(sub)}), {steps} reverse steps "
              f"timed after 1 warm-up at {dt:.3f} s/step on {_BEST_THREADS} threads (best of a 1-step probe; box has {cores} "
              f"cores), extrapolated x{start_t} steps")
    return rate, _BEST_THREADS, sample, dt


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lengths, n_pad, start_t, wrap_all, wl_name = workload(args)
    T = args.timesteps
    B = len(lengths)
    flops_step = synthetic.algorithmic_flops(synthetic.PRODUCTION, lengths)

    if args.impl == "reference":
        if rank != 0:
            return 0
        # each bench "step" of this arm is one bounded sample (cpu_steps reverse steps over cpu_chains chains of the
        # workload's length mix, extrapolated to the full loop); W warm-up samples, then K timed ones
        reps, dts = [], []
        sample, cores = "", os.cpu_count() or 1
        for i in range(args.warmup + args.steps):
            r, cores, sample, d = cpu_reference_rate(lengths, n_pad, T, args.cpu_chains, max(1, args.cpu_steps), start_t, wrap_all)
            if i >= args.warmup:
                reps.append(r)
                dts.append(d)
        value = float(np.mean(reps))
        dt = float(np.mean(dts))
        line = {
            "impl": "reference", "metric": "backbones/sec", "value": value, "unit": "backbones/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt * max(1, args.cpu_steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "timesteps": T,
                       "note": ("CPU path: the unmodified reference installed in baseline/_ref - its own loop (sampling.p_sample), "
                                "schedules, wrap and model sub-modules; encoder layers from the installed transformers' modules "
                                "because transformers==4.11.3 is not installable (baseline/reference_arm.py)") if _reference() else
                               ("CPU path: fp32 torch restatement of the reference forward (HF 4.11.3 encoder not "
                                "installable) + the reference's loop arithmetic; baseline/_ref absent")},
            "cpu_baseline": {"value": value, "unit": "backbones/s", "cores": cores, "kind": cpu_kind(), "sample": sample},
            "e2e": {"value": value, "unit": "backbones/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------------
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU fallback"
    from foldingdiff_b200 import distributed as fdist
    from foldingdiff_b200 import modelling, sampling
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    cfg = modelling.BertConfig(**synthetic.PRODUCTION)
    model = modelling.BertForDiffusionBase(cfg, ft_is_angular=[True] * 6, gemm=args.gemm)
    model.load_state_dict(synthetic.synthetic_state_dict(synthetic.PRODUCTION, seed=0))
    model = model.to(dev)
    eng = model.native_engine()
    dset, noise_host = initial_noise(B, n_pad, T)
    noise_host = noise_host.pin_memory()
    betas = dset.alpha_beta_terms["betas"]
    wrap = [True] * 6
    eng.set_schedule(betas, T)
    eng.set_batch(lengths, n_pad)
    noise_dev = noise_host.to(dev)
    gather = [torch.empty((B, n_pad, 6), device=dev) for _ in range(world)] if world > 1 else None

    def one_pass_device():
        x = noise_dev.clone()
        sampling._run_steps(eng, x, start_t, wrap, None)
        if world > 1:
            dist.all_gather(gather, x)  # the job's only collective: finished angles
        return x

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        one_pass_device()
    launches0 = eng.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = timed(one_pass_device, args.steps)
    launches = eng.launch_count() - launches0
    clocks = clk.summary()
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step / 1000.0)

    # ---- e2e: host noise in, host result out, through the public API ------------------------------
    e2e = None
    if not args.no_e2e:
        hist_mode = args.e2e_history

        def one_pass_e2e():
            if wrap_all:
                out = sampling.denoise_from(model, noise_host, lengths, start_t, betas).cpu()
            else:
                out = sampling.p_sample_loop(model, lengths, noise_host, T, betas, is_angle=wrap, history=hist_mode)
            return out

        one_pass_e2e()  # allocator warm-up for the history buffers
        ms_e = timed(one_pass_e2e, args.steps) / args.steps
        steps_out = start_t if (hist_mode == "full" and not wrap_all) else 1
        e2e = {"value": world * B / (ms_e / 1000.0), "unit": "backbones/s",
               "h2d_bytes_per_step": int(noise_host.numel() * 4),
               "d2h_bytes_per_step": int(steps_out * B * n_pad * 6 * 4),
               "ms_per_step": ms_e, "api": "sampling.p_sample_loop(host noise) -> host tensor, history=" + hist_mode}

    # ---- per-kernel device times (CUDA events on the launch stream), rank 0 only -------------------
    roofline, kernels = None, None
    if rank == 0:
        nprof = min(args.profile_steps, start_t)
        x = noise_dev.clone()
        z = torch.randn((nprof, B, n_pad, 6), device=dev)
        torch.cuda.synchronize()
        eng.profile_begin()
        eng.p_sample_steps(x, start_t, start_t - nprof, z, None, wrap)
        prof = eng.profile_end()
        total = sum(v[0] for v in prof.values()) or 1.0
        kernels = {k: {"ms_per_reverse_step": v[0] / nprof, "launches_per_reverse_step": v[1] / nprof,
                       "share": v[0] / total} for k, v in prof.items() if v[1]}
        n = np.asarray(lengths, dtype=np.float64)
        H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        gemm_flops = {"gemm_qkv": L * 6 * H * H * n.sum(), "gemm_attn_out": L * 2 * H * H * n.sum(),
                      "gemm_ffn1": L * 2 * H * I * n.sum(), "gemm_ffn2": L * 2 * H * I * n.sum(),
                      "gemm_head": 2 * H * H * n.sum()}
        att_flops = L * 6 * H * (n * n).sum()
        peak_tf, peak_src = FALLBACK_PEAK_TFLOPS, "fallback (B200_PROFILING.md)"
        if os.path.isfile(PEAKS_FILE):
            with open(PEAKS_FILE) as f:
                pk = json.load(f)
            peak_tf, peak_src = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
        gemm_ms = sum(prof[k][0] for k in gemm_flops) / nprof
        gemm_tf = sum(gemm_flops.values()) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        att_ms = prof["attention"][0] / nprof
        att_tf = att_flops / (att_ms * 1e-3) / 1e12 if att_ms > 0 else 0.0
        dominant = "projection GEMMs (tc_gemm_kernel / sgemm_tn_kernel)" if gemm_ms >= att_ms else "relative-key attention (attention_tc_kernel / attention_simt_kernel)"
        dom_tf = gemm_tf if gemm_ms >= att_ms else att_tf
        roofline = {"bound": "tensor", "kernel": dominant, "achieved": dom_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": dom_tf / peak_tf, "peak_source": peak_src,
                    "traffic": (float(np.mean(list(NCU_GEMM["dram_bytes_per_launch"].values())))
                                if (args.workload == "config2" and args.gemm == "tc3x" and gemm_ms >= att_ms) else None),
                    "traffic_note": "ncu DRAM bytes per GEMM launch, mean over the four projections; " + NCU_GEMM["source"],
                    "tensor_pipe_active_pct_ncu": NCU_GEMM["tensor_pipe_active_pct"],
                    "gemm_tflops_algorithmic": gemm_tf, "attention_tflops_algorithmic": att_tf,
                    "whole_step_tflops_algorithmic": flops_step * start_t / (ms_per_step * 1e-3) / 1e12,
                    "note": "algorithmic FLOPs (valid tokens, 2 flop/MAC, SURVEY 8d) / CUDA-event kernel time; "
                            "the 3-pass split issues 3x these MMAs"}

    # ---- section 8f rank-1 row: batched NeRF (angles -> backbone coordinates), reported beside the headline ----
    nerf_line = None
    if rank == 0:
        from foldingdiff_b200 import nerf as fnerf
        names = ["phi", "psi", "omega", "tau", "CA:C:1N", "C:1N:1CA"]
        ang = noise_dev.clone()
        ang[..., 3:] = ang[..., 3:].abs() * 0.2 + 1.7
        for _ in range(3):
            fnerf.build_backbone(ang, lengths, names)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            xyz = fnerf.build_backbone(ang, lengths, names)
        e1.record()
        torch.cuda.synchronize()
        gpu_rate = reps * B / (e0.elapsed_time(e1) / 1000.0)
        nerf_line = {"gpu_structures_per_s": gpu_rate, "atoms_per_structure": "3 x length (N, CA, C)",
                     "api": "foldingdiff_b200.nerf.build_backbone (fd_nerf_build)", "chains": B}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import nerf as onerf  # CPU baseline leg: the reference's per-chain Python loop, restated
            host = ang[:8].cpu().numpy()
            t0 = time.perf_counter()
            for i in range(8):
                onerf.build_chain(host[i, : lengths[i]], names)
            nerf_line["cpu_structures_per_s_1core"] = 8 / (time.perf_counter() - t0)
            nerf_line["cpu_sample"] = "8 chains, oracle restatement of nerf.NERFBuilder (single core, as one pool worker)"

    # ---- section 8f rank-2 row: output writers (csv.gz + PDB per chain), native batch call vs the reference's way ----
    if rank == 0 and nerf_line is not None:
        import shutil
        import tempfile
        from foldingdiff_b200 import writers as fwriters
        tmp = tempfile.mkdtemp(prefix="fd_writers_")
        try:
            ang_h, xyz_h = ang.cpu().numpy(), xyz.cpu().numpy()
            t0 = time.perf_counter()
            fwriters.write_batch(lengths, angles=ang_h, feature_names=names, csv_paths=[f"{tmp}/g{i}.csv.gz" for i in range(B)],
                                 coords=xyz_h, pdb_paths=[f"{tmp}/g{i}.pdb" for i in range(B)])
            dt = time.perf_counter() - t0
            wline = {"native_chains_per_s": B / dt, "files_per_chain": 2, "chains": B,
                     "api": "foldingdiff_b200.writers.write_batch (fd_write_batch)", "host_threads": min(32, os.cpu_count() or 1)}
            if world == 1 and not args.no_cpu_baseline:
                import pandas as pd
                from oracle import writers as owriters  # CPU baseline leg: DataFrame.to_csv + the PDB text restatement
                t0 = time.perf_counter()
                for i in range(16):
                    pd.DataFrame(ang_h[i, : lengths[i]], columns=names).to_csv(f"{tmp}/ref{i}.csv.gz")
                    with open(f"{tmp}/ref{i}.pdb", "w") as f:
                        f.write(owriters.backbone_pdb_text(xyz_h[i, : 3 * lengths[i]]))
                wline["cpu_chains_per_s_1core"] = 16 / (time.perf_counter() - t0)
                wline["cpu_sample"] = "16 chains, DataFrame.to_csv(.csv.gz) + Python PDB text (single core, as one pool worker)"
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    else:
        wline = None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, sample, _ = cpu_reference_rate(lengths, n_pad, T, args.cpu_chains, args.cpu_steps, start_t, wrap_all)
        cpu_baseline = {"value": rate, "unit": "backbones/s", "cores": cores, "kind": cpu_kind(), "sample": sample}

    if rank == 0:
        line = {
            "metric": "backbones/sec", "value": value, "unit": "backbones/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"tc3x": "f32 (fp16 hi/lo 3-pass tensor-core GEMMs, fp32 accumulate)",
                                                              "fp32": "f32", "tc1x": "f16 (single-pass, NOT parity mode)"}[args.gemm],
            "data": "synthetic",
            "config": {"workload": wl_name, "timesteps": T, "reverse_steps_per_pass": start_t, "chains_per_gpu": B,
                       "gemm": args.gemm, "parallelism": f"dp{world} (independent chains, one all-gather of final angles)",
                       "l2": "inputs larger than L2: ~1 GB of activations per reverse step, 1000 steps per pass",
                       "algorithmic_tflop_per_pass": flops_step * start_t / 1e12},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
            "cpu_baseline": cpu_baseline, "next_rows": {"nerf": nerf_line, "writers": wline},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
