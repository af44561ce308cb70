"""
Python owner of one native sampler handle (include/foldingdiff_b200.h).

PyTorch is used here only for what it is good at on the host side: device memory
(`tensor.data_ptr()`), the current CUDA stream and the device RNG.  Every FLOP of the
forward and of the reverse step runs in the hand-written CUDA behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native, beta_schedules

LAYER_KEYS = [
    "attention.self.query.weight", "attention.self.query.bias",
    "attention.self.key.weight", "attention.self.key.bias",
    "attention.self.value.weight", "attention.self.value.bias",
    "attention.self.distance_embedding.weight",
    "attention.output.dense.weight", "attention.output.dense.bias",
    "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias",
    "intermediate.dense.weight", "intermediate.dense.bias",
    "output.dense.weight", "output.dense.bias",
    "output.LayerNorm.weight", "output.LayerNorm.bias",
]
HEAD_KEYS = ["inputs_to_hidden_dim.weight", "inputs_to_hidden_dim.bias",
             "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]
TAIL_KEYS = ["token_decoder.dense1.weight", "token_decoder.dense1.bias",
             "token_decoder.layer_norm.weight", "token_decoder.layer_norm.bias",
             "token_decoder.dense2.weight", "token_decoder.dense2.bias"]


def weight_key_order(n_layers: int):
    """State-dict keys in the order fd_create expects them (header comment above FD_W_HEAD)."""
    keys = list(HEAD_KEYS)
    for l in range(n_layers):
        keys += [f"encoder.layer.{l}.{k}" for k in LAYER_KEYS]
    return keys + list(TAIL_KEYS)


def gaussian_fourier_table(W: torch.Tensor, timesteps: int) -> torch.Tensor:
    """
    (T, H) table of GaussianFourierProjection(t), t = 0..T-1, in the reference's fp32 op order
    (/root/reference/foldingdiff/modelling.py:69-70): t[:, None] * W[None, :] * 2 * pi, then
    cat[sin, cos].  Evaluated on the CPU: the arguments reach ~1e4 rad where one fp32 ulp is
    ~1e-3 rad, so the table must be produced exactly once, the reference's way.
    """
    return gaussian_fourier_rows(W, torch.arange(timesteps))


def gaussian_fourier_rows(W: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    W = W.detach().to("cpu", torch.float32)
    t = t.detach().to("cpu")
    proj = t[:, None] * W[None, :] * 2 * torch.pi
    return torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1).contiguous()


def sinusoidal_rows(dim: int, t: torch.Tensor) -> torch.Tensor:
    """SinusoidalPositionEmbeddings (/root/reference/foldingdiff/modelling.py:83-93)."""
    import math
    half = dim // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    ang = t.detach().to("cpu")[:, None] * freq[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1).to(torch.float32).contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Engine:
    """One native handle bound to one CUDA device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], *, hidden: int, layers: int, heads: int,
                 intermediate: int, max_pos: int, n_features: int, ln_eps: float,
                 time_rows_fn, device: torch.device, gemm: str = "tc3x", head_ln_eps: float = 1e-12):
        if device.type != "cuda":
            raise _native.NativeError("foldingdiff_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _native.lib()
        self.device = device
        self.hidden, self.n_features = hidden, n_features
        self._time_rows_fn = time_rows_fn
        self._schedule_key = None
        self._batch_key = None
        self.gemm = gemm
        # placeholder schedule (1 step); p_sample_loop installs the real one via set_schedule
        dims = _native.FdDims(hidden, layers, heads, intermediate, max_pos, n_features, 1, ln_eps, head_ln_eps)
        keys = weight_key_order(layers)
        missing = [k for k in keys if k not in state_dict]
        if missing:
            raise KeyError(f"state_dict is missing {missing[:3]}... ({len(missing)} keys)")
        host = [state_dict[k].detach().to("cpu", torch.float32).contiguous() for k in keys]
        arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
        tt = time_rows_fn(torch.arange(1))
        coef = torch.zeros(1, 4)
        handle = C.c_void_p()
        index = device.index if device.index is not None else torch.cuda.current_device()
        _native.check(self.lib.fd_create(C.byref(dims), arr, len(host), tt.data_ptr(), coef.data_ptr(),
                                         index, _native.GEMM_MODES[gemm], C.byref(handle)), "fd_create")
        self._h = handle
        # the C entry points restore the caller's CUDA device themselves (include/foldingdiff_b200.h, Conventions);
        # `with torch.cuda.device(...)` below only makes torch's current STREAM the one of the handle's device

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.fd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ----------------------------------------------------------------------------
    def set_gemm(self, gemm: str):
        _native.check(self.lib.fd_set_gemm_mode(self._h, _native.GEMM_MODES[gemm]), "fd_set_gemm_mode")
        self.gemm = gemm

    def set_schedule(self, betas: torch.Tensor, timesteps: Optional[int] = None):
        betas = betas.detach().to("cpu", torch.float32).contiguous()
        T = int(betas.shape[0]) if timesteps is None else int(timesteps)
        key = (T, betas.numpy().tobytes())
        if key == self._schedule_key:
            return
        coef = beta_schedules.step_coefficients(betas)
        assert coef.shape[0] >= T, f"betas has {coef.shape[0]} entries, need {T}"
        coef = coef[:T].contiguous()
        table = self._time_rows_fn(torch.arange(T))
        with torch.cuda.device(self.device):
            _native.check(self.lib.fd_set_schedule(self._h, T, table.data_ptr(), coef.data_ptr()), "fd_set_schedule")
        self._schedule_key = key
        self.timesteps = T

    def time_rows(self, t: torch.Tensor) -> torch.Tensor:
        return self._time_rows_fn(t)

    def set_batch(self, lengths: Sequence[int], n_pad: int, all_rows: bool = False,
                  key_mask: Optional[torch.Tensor] = None):
        lens = np.ascontiguousarray(np.asarray([int(l) for l in lengths], dtype=np.int32))
        km = None
        if key_mask is not None:
            km = key_mask.detach().to("cpu", torch.float32).contiguous()
        key = (lens.tobytes(), int(n_pad), bool(all_rows), None if km is None else km.numpy().tobytes())
        if key == self._batch_key:
            return
        with torch.cuda.device(self.device):
            _native.check(self.lib.fd_set_batch(self._h, len(lens), int(n_pad), lens.ctypes.data, int(all_rows),
                                                _ptr(km), _stream()), "fd_set_batch")
        self._batch_key = key

    # -- compute ----------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert temb.is_cuda and temb.dtype == torch.float32 and temb.is_contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _native.check(self.lib.fd_forward(self._h, x.data_ptr(), temb.data_ptr(), out.data_ptr(), _stream()),
                          "fd_forward")
        return out

    def p_sample_steps(self, x: torch.Tensor, t_hi: int, t_lo: int, noise: Optional[torch.Tensor],
                       history: Optional[torch.Tensor], wrap_mask: Sequence[bool]):
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        for t in (noise, history):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
        wm = (C.c_uint8 * self.n_features)(*[1 if w else 0 for w in wrap_mask])
        with torch.cuda.device(self.device):
            _native.check(self.lib.fd_p_sample_steps(self._h, x.data_ptr(), int(t_hi), int(t_lo), _ptr(noise),
                                                     _ptr(history), wm, _stream()), "fd_p_sample_steps")

    def p_sample_steps_philox(self, x: torch.Tensor, t_hi: int, t_lo: int, seed: int, offset: int,
                              history: Optional[torch.Tensor], wrap_mask: Sequence[bool]):
        """Throughput mode: the step kernel draws its own normals (library Philox stream `seed`, element `offset`
        onwards) - no per-step host work, no noise tensor in HBM."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert history is None or (history.is_cuda and history.dtype == torch.float32 and history.is_contiguous())
        wm = (C.c_uint8 * self.n_features)(*[1 if w else 0 for w in wrap_mask])
        with torch.cuda.device(self.device):
            _native.check(self.lib.fd_p_sample_steps_philox(self._h, x.data_ptr(), int(t_hi), int(t_lo), int(seed),
                                                            int(offset), _ptr(history), wm, _stream()),
                          "fd_p_sample_steps_philox")

    def check_status(self):
        """Raise NativeError if a tensor-core pipeline of an earlier asynchronous call timed out.  Call after the
        stream has been synchronised (the loop does, wherever it hands results to the caller)."""
        _native.check(self.lib.fd_status(self._h), "fd_status")

    def sample_host(self, lengths, x0: np.ndarray, t_start: int, noise: Optional[np.ndarray], seed: int,
                    wrap_mask, full_history: bool) -> np.ndarray:
        """fd_sample_host: numpy in, numpy out (the pure C-ABI path a non-Python caller would take)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float32)
        B, N, F = x0.shape
        lens = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
        out = np.empty((t_start, B, N, F) if full_history else (B, N, F), dtype=np.float32)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
        wm = (C.c_uint8 * self.n_features)(*[1 if w else 0 for w in wrap_mask])
        self._batch_key = None
        _native.check(self.lib.fd_sample_host(self._h, B, N, lens.ctypes.data, x0.ctypes.data, int(t_start),
                                              None if nz is None else nz.ctypes.data, int(seed), wm,
                                              int(full_history), out.ctypes.data), "fd_sample_host")
        return out

    def profile_begin(self):
        _native.check(self.lib.fd_profile_begin(self._h), "fd_profile_begin")

    def profile_end(self):
        """-> {category: (total_ms, launches)} for the kernels launched since profile_begin()."""
        n = self.lib.fd_profile_num_categories()
        ms = (C.c_float * n)()
        cnt = (C.c_int64 * n)()
        _native.check(self.lib.fd_profile_end(self._h, ms, cnt), "fd_profile_end")
        return {self.lib.fd_profile_category_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def launch_count(self) -> int:
        return int(self.lib.fd_launch_count(self._h))
