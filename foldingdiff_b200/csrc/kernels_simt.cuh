// CUDA-core (fp32 FMA) kernels of the reverse-diffusion step.
//
// These are (a) every non-GEMM stage of the step - embedding, relative-key attention,
// LayerNorm, decoder tail + posterior update + mod-2pi wrap - and (b) an fp32 SIMT GEMM that
// is the in-GPU reference arithmetic (FD_GEMM_FP32_SIMT) the tensor-core path is checked
// against.  Data layout: activations are PACKED rows - only the computed tokens of every chain,
// chain after chain, `row_src[r] = b * n_pad + n` maps a packed row back to (chain, residue).
//
// Reference arithmetic being restated (paths under /root/reference):
//   embed_kernel      modelling.py:464-472 (+ BertEmbeddings :168, GaussianFourierProjection table)
//   attention_simt    transformers 4.11.3 BertSelfAttention, relative_key (call site modelling.py:473)
//   layernorm_kernel  BertSelfOutput / BertOutput LayerNorm
//   tail_kernel       modelling.py:203-208 (AnglesPredictor LN + dense2) fused with
//                     sampling.py:62-75 (posterior step) and :119-130 (per-column wrap)
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "philox.cuh"

namespace fd {

struct StepCoef {
  float c1, beta, s, sigma;
  int add_noise;  // t > 0
};
// Source of the step's normals z (sampling.py:73): the caller's draws (z != nullptr: slice of the step) or, in the
// throughput mode, element `offset + b * n_pad * F + n * F + f` of the library's Philox stream `seed` - the same
// element fd_randn would have written there, so the two paths agree bit for bit.
struct StepNoise {
  const float* z;
  unsigned long long seed, offset;
};
// Per-step arguments of a reverse step replayed as a CUDA graph (api.cu: run_steps): the graph's kernel arguments are
// frozen at capture time, so what changes from step to step - the time-embedding row, the posterior coefficients, the
// noise slice, the history slice - lives in one small device struct that is rewritten (stream-ordered) before every
// graph launch.  embed_kernel and tail_kernel read it when their `dyn` argument is non-null.
struct StepDyn {
  const float* temb;
  StepNoise noise;
  float* hist;
  StepCoef coef;
};

// ------------------------------------------------------------------------------------------------
// embed: h[r, :] = LN(x[b, n, :] @ W_in^T + b_in) * g + beta  +  temb[b, :]
// One warp per packed row at a time, lane owns columns lane + 32 * i; the warps are PERSISTENT (grid-stride over
// the rows) and keep everything that does not depend on the row in registers - the input projection (F <= 6 features:
// every shipped model has 6), its bias, the LayerNorm vectors and, in the sampling loop where every chain shares the
// step's time embedding, that row too.  (Round 1 launched one warp per row and re-read ~100 weights per row through
// L1: 90 us per reverse step for 137 MB of output, a quarter of the HBM rate.)
// ------------------------------------------------------------------------------------------------
constexpr int EMBED_FR = 6;  // features whose projection weights are held in registers (every shipped feature set has <= 6)

template <int VPL>
__global__ void __launch_bounds__(128, 3)
embed_kernel(const float* __restrict__ x, const int* __restrict__ row_src, int n_rows, int n_pad,
             int F, const float* __restrict__ w_in, const float* __restrict__ b_in,
             const float* __restrict__ g, const float* __restrict__ bta, float eps,
             const float* __restrict__ temb, int temb_stride, float* __restrict__ h_out,
             __half* __restrict__ o_hi, __half* __restrict__ o_lo, const StepDyn* __restrict__ dyn) {
  constexpr int H = VPL * 32;
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const bool wreg = F <= EMBED_FR;
  float w[VPL][EMBED_FR], bi[VPL], gg[VPL], bb[VPL];  // weights: independent of the chain of kernels
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int f = 0; f < EMBED_FR; ++f) w[i][f] = (wreg && f < F) ? w_in[c * F + f] : 0.0f;
    bi[i] = b_in[c]; gg[i] = g[c]; bb[i] = bta[c];
  }
  pdl_wait();  // x is the previous step's output; h / planes were read by the previous step's kernels
  const float* te0 = dyn ? dyn->temb : temb;
  float tev[VPL];
  if (temb_stride == 0) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) tev[i] = te0[lane + 32 * i];
  }
  for (int r = gw; r < n_rows; r += nw) {
    const int src = row_src[r];
    float xin[EMBED_FR];
#pragma unroll
    for (int f = 0; f < EMBED_FR; ++f) xin[f] = (f < F) ? x[(size_t)src * F + f] : 0.0f;
    float v[VPL];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float a = 0.0f;
      if (wreg) {  // zero weights beyond F leave the sum untouched: same fp32 operations, same order, as a loop over F
#pragma unroll
        for (int f = 0; f < EMBED_FR; ++f) a = fmaf(xin[f], w[i][f], a);
      } else {
        const int c = lane + 32 * i;
        for (int f = 0; f < F; ++f) a = fmaf(x[(size_t)src * F + f], w_in[c * F + f], a);
      }
      v[i] = a + bi[i];
      sum += v[i];
    }
    const float mean = warp_sum(sum) * (1.0f / H);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float d = v[i] - mean;
      sq = fmaf(d, d, sq);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / H) + eps);
    const float* te = te0 + (size_t)(src / n_pad) * temb_stride;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      const float y = ((v[i] - mean) * rstd) * gg[i] + bb[i] + (temb_stride == 0 ? tev[i] : te[c]);
      h_out[(size_t)r * H + c] = y;
      if (o_hi) {  // fp16 hi / lo operand planes for the tensor-core GEMM that consumes this row
        const __half hh = __float2half_rn(y);
        o_hi[(size_t)r * H + c] = hh;
        if (o_lo) o_lo[(size_t)r * H + c] = __float2half_rn(y - __half2float(hh));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layernorm: out[r, :] = LN(in[r, :] (+ resid[r, :])) * g + beta      (warp per row)
// The residual add of BertSelfOutput / BertOutput lives here in the tensor-core modes: this kernel
// streams its rows with full coalescing and deep memory-level parallelism, whereas the same 69 MB
// read inside the GEMM epilogue was latency-exposed (measured: +30% on those GEMMs).
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ in, const float* __restrict__ resid, int n_rows,
                 const float* __restrict__ g,
                 const float* __restrict__ bta, float eps, float* __restrict__ out,
                 __half* __restrict__ o_hi, __half* __restrict__ o_lo) {
  constexpr int H = VPL * 32;
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  float v[VPL];
  float sum = 0.0f;
#pragma unroll
  // `in` (the GEMM's fp32 output) and `resid` are dead after this kernel: streaming loads (evict-first)
  // keep them from displacing the operand planes written below, which the next GEMM reads right away.
  for (int i = 0; i < VPL; ++i) v[i] = __ldcs(in + (size_t)warp * H + lane + 32 * i);
  if (resid) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] += __ldcs(resid + (size_t)warp * H + lane + 32 * i);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) sum += v[i];
  const float mean = warp_sum(sum) * (1.0f / H);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float d = v[i] - mean;
    sq = fmaf(d, d, sq);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    const float y = ((v[i] - mean) * rstd) * g[c] + bta[c];
    __stcs(out + (size_t)warp * H + c, y);  // next read is two kernels away (the following LayerNorm's residual)
    if (o_hi) {
      const __half hh = __float2half_rn(y);
      o_hi[(size_t)warp * H + c] = hh;
      if (o_lo) o_lo[(size_t)warp * H + c] = __float2half_rn(y - __half2float(hh));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 SIMT GEMM:  C[M, N] = A[M, K] * W[N, K]^T + bias (+ residual) (+ GELU)
// 128 x 64 x 16 tiles, 256 threads, 8 x 4 register micro-tile, register-prefetch double buffer.
// M % 128 == 0, N % 64 == 0, K % 16 == 0 (the caller pads the row count).
// ------------------------------------------------------------------------------------------------
enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2,
       EPI_LNIN = 3, EPI_LNIN_GELU = 4, EPI_LNRES = 5 };  // tensor-core path only: LayerNorm folded into the projections (gemm_tc.cuh: TcLn)

template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ W,
                const float* __restrict__ bias, const float* __restrict__ resid,
                float* __restrict__ C, int M, int N, int K) {
  constexpr int BM = 128, BN = 64, BK = 16, PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader coordinates
  const float* a_ptr0 = A + (size_t)(m0 + lrow) * K + lk;
  const float* a_ptr1 = A + (size_t)(m0 + lrow + 64) * K + lk;
  const float* w_ptr = W + (size_t)(n0 + lrow) * K + lk;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  float4 ra0 = *reinterpret_cast<const float4*>(a_ptr0);
  float4 ra1 = *reinterpret_cast<const float4*>(a_ptr1);
  float4 rw = *reinterpret_cast<const float4*>(w_ptr);
  auto stash = [&](int buf) {
    As[buf][lk + 0][lrow] = ra0.x; As[buf][lk + 1][lrow] = ra0.y;
    As[buf][lk + 2][lrow] = ra0.z; As[buf][lk + 3][lrow] = ra0.w;
    As[buf][lk + 0][lrow + 64] = ra1.x; As[buf][lk + 1][lrow + 64] = ra1.y;
    As[buf][lk + 2][lrow + 64] = ra1.z; As[buf][lk + 3][lrow + 64] = ra1.w;
    Ws[buf][lk + 0][lrow] = rw.x; Ws[buf][lk + 1][lrow] = rw.y;
    Ws[buf][lk + 2][lrow] = rw.z; Ws[buf][lk + 3][lrow] = rw.w;
  };
  stash(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      ra0 = *reinterpret_cast<const float4*>(a_ptr0 + (size_t)(kt + 1) * BK);
      ra1 = *reinterpret_cast<const float4*>(a_ptr1 + (size_t)(kt + 1) * BK);
      rw = *reinterpret_cast<const float4*>(w_ptr + (size_t)(kt + 1) * BK);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }
  const float4 bv = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t off = (size_t)(m0 + ty * 8 + i) * N + n0 + tx * 4;
    float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
    if (EPI == EPI_BIAS_RESID) {
      const float4 r = *reinterpret_cast<const float4*>(resid + off);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (EPI == EPI_BIAS_GELU) {
      o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
    }
    *reinterpret_cast<float4*>(C + off) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// relative-key attention, fp32 CUDA cores.  grid = (heads, batch), 128 threads, thread = query row.
//   S[l, r] = (q_l . k_r + q_l . E[l - r + P - 1]) / sqrt(32) + key_bias[r];  ctx_l = softmax_r(S) V
// qkv is packed rows x 3H (q | k | v, head h at columns h*32).  Keys r >= n_keys are skipped, which
// equals the reference's additive -10000 (exp underflows to exactly 0 in fp32).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_simt_kernel(const float* __restrict__ qkv, const int* __restrict__ row_start,
                      const int* __restrict__ n_rows_arr, const int* __restrict__ n_keys_arr,
                      const float* __restrict__ key_bias, int n_pad,
                      const float* __restrict__ dist_emb, int max_pos, int H,
                      float* __restrict__ ctx) {
  constexpr int D = FD_HEAD_DIM, KB = 16, EP = D + 1;
  extern __shared__ __align__(16) float smem[];
  const int head = blockIdx.x, chain = blockIdx.y;
  const int r0 = row_start[chain];
  const int n_rows = n_rows_arr[chain], n_keys = n_keys_arr[chain];
  float* Ks = smem;                 // [n_keys][32]
  float* Vs = Ks + n_keys * D;      // [n_keys][32]
  float* Es = Vs + n_keys * D;      // [n_rows + n_keys - 1][33]
  float* Bs = Es + (n_rows + n_keys - 1) * EP;  // [n_keys]
  const int tid = threadIdx.x;
  const int ld = 3 * H;
  for (int i = tid; i < n_keys * (D / 4); i += blockDim.x) {
    const int r = i / (D / 4), c4 = (i % (D / 4)) * 4;
    const float* base = qkv + (size_t)(r0 + r) * ld + head * D + c4;
    *reinterpret_cast<float4*>(Ks + r * D + c4) = *reinterpret_cast<const float4*>(base + H);
    *reinterpret_cast<float4*>(Vs + r * D + c4) = *reinterpret_cast<const float4*>(base + 2 * H);
  }
  const int e_lo = max_pos - n_keys;  // first distance-embedding row needed
  const int n_e = n_rows + n_keys - 1;
  for (int i = tid; i < n_e * D; i += blockDim.x) {
    const int r = i / D, c = i % D;
    Es[r * EP + c] = dist_emb[(size_t)(e_lo + r) * D + c];
  }
  for (int i = tid; i < n_keys; i += blockDim.x)
    Bs[i] = key_bias ? key_bias[(size_t)chain * n_pad + i] : 0.0f;
  __syncthreads();
  const int l = tid;
  if (l >= n_rows) return;

  float q[D], acc[D];
  {
    const float* qp = qkv + (size_t)(r0 + l) * ld + head * D;
#pragma unroll
    for (int c = 0; c < D; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + c);
      q[c] = t.x; q[c + 1] = t.y; q[c + 2] = t.z; q[c + 3] = t.w;
    }
  }
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.0f;
  float m = -INFINITY, sum = 0.0f;
  const float sqrt_d = 5.65685424949238019521f;  // math.sqrt(32) rounded to fp32
  for (int k0 = 0; k0 < n_keys; k0 += KB) {
    float s[KB];
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const int r = k0 + j;
      if (r < n_keys) {
        const float* kr = Ks + r * D;
        const float* er = Es + (l - r + n_keys - 1) * EP;
        float dk = 0.0f, de = 0.0f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          dk = fmaf(q[c], kr[c], dk);
          de = fmaf(q[c], er[c], de);
        }
        s[j] = __fdiv_rn(dk + de, sqrt_d) + Bs[r];
      } else {
        s[j] = -INFINITY;
      }
      bm = fmaxf(bm, s[j]);
    }
    const float m_new = fmaxf(m, bm);
    const float scale = expf(m - m_new);  // exp(-inf) = 0 on the first block
    sum *= scale;
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] *= scale;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const int r = k0 + j;
      if (r < n_keys) {
        const float p = expf(s[j] - m_new);
        sum += p;
        const float* vr = Vs + r * D;
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fmaf(p, vr[c], acc[c]);
      }
    }
    m = m_new;
  }
  const float inv = 1.0f / sum;
  float* op = ctx + (size_t)(r0 + l) * H + head * D;
#pragma unroll
  for (int c = 0; c < D; c += 4)
    *reinterpret_cast<float4*>(op + c) =
        make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
}

// ------------------------------------------------------------------------------------------------
// tail: eps = LN(u) @ W2^T + b2   (u = gelu(dense1(h)), modelling.py:203-207), then either
//   MODE_EPS    : eps_out[b, n, :] = eps                                   (fd_forward)
//   MODE_SAMPLE : x = c1 * (x - beta * eps / s) [+ sigma * z]; wrap; history (sampling.py:62-75,119-131)
// one warp per packed row.  The posterior arithmetic uses explicitly rounded fp32 ops in the
// reference's order (multiply by beta, then divide by s) so it is bit-comparable given equal eps.
// ------------------------------------------------------------------------------------------------
template <int VPL, bool SAMPLE>
__global__ void __launch_bounds__(256)
tail_kernel(const float* __restrict__ u, const int* __restrict__ row_src, int n_rows, int F,
            const float* __restrict__ g, const float* __restrict__ bta, float eps_ln,
            const float* __restrict__ w2, const float* __restrict__ b2,
            float* __restrict__ eps_out,                         // MODE_EPS
            float* __restrict__ x, StepNoise noise,              // MODE_SAMPLE
            float* __restrict__ hist, StepCoef coef, uint32_t wrap_bits, const StepDyn* __restrict__ dyn) {
  constexpr int H = VPL * 32;
  pdl_trigger();
  extern __shared__ __align__(16) float w2s[];  // [F][H]
  for (int i = threadIdx.x; i < F * H; i += blockDim.x) w2s[i] = w2[i];  // weights: no dependence on the chain
  const int lane = threadIdx.x & 31;
  float gg[VPL], bb[VPL];  // head LayerNorm vectors: weights, fetched before the wait
#pragma unroll
  for (int i = 0; i < VPL; ++i) { gg[i] = g[lane + 32 * i]; bb[i] = bta[lane + 32 * i]; }
  __syncthreads();
  pdl_wait();
  if (SAMPLE && dyn) { noise = dyn->noise; hist = dyn->hist; coef = dyn->coef; }  // graph replay: this step's arguments
  // persistent warps: the 9 KB of dense2 weights are staged once per block, not once per 8 rows
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < n_rows; r += nw) {
    float v[VPL];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = u[(size_t)r * H + lane + 32 * i];
      sum += v[i];
    }
    const float mean = warp_sum(sum) * (1.0f / H);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float d = v[i] - mean;
      sq = fmaf(d, d, sq);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / H) + eps_ln);
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = ((v[i] - mean) * rstd) * gg[i] + bb[i];
    float mine = 0.0f;  // lane f keeps eps[f]
    for (int f = 0; f < F; ++f) {
      float p = 0.0f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) p = fmaf(v[i], w2s[f * H + lane + 32 * i], p);
      p = warp_sum(p);
      if (lane == f) mine = p + b2[f];
    }
    if (lane < F) {
      const size_t idx = (size_t)row_src[r] * F + lane;
      if (!SAMPLE) {
        eps_out[idx] = mine;
      } else {
        const float xv = x[idx];
        float y = __fmul_rn(coef.c1, __fsub_rn(xv, __fdiv_rn(__fmul_rn(coef.beta, mine), coef.s)));
        if (coef.add_noise) {
          const float zv = noise.z ? noise.z[idx] : philox_normal(noise.seed, noise.offset + idx);
          y = __fadd_rn(y, __fmul_rn(coef.sigma, zv));
        }
        if ((wrap_bits >> lane) & 1u) y = wrap_pi(y);
        x[idx] = y;
        if (hist) hist[idx] = y;
      }
    }
  }
}

}  // namespace fd
