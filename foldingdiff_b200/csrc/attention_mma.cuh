// Relative-key self-attention on tensor cores (warp-level mma.sync.m16n8k16, fp16 in / fp32 accumulate).
//
//   S[l, r] = (q_l . k_r + q_l . E[l - r + P - 1]) / sqrt(32) + bias_r ;  ctx_l = softmax_r(S) V
//   (transformers 4.11.3 BertSelfAttention, position_embedding_type = "relative_key";
//    call site /root/reference/foldingdiff/modelling.py:473)
//
// Building blocks of the mma.sync attention kernel (attention_pool.cuh, used by the 1-pass tc1x mode and as the A/B
// alternative of the tcgen05 kernel): one warp = 16 query rows of one (chain, head) against all keys of the chain,
// which sit in shared memory.  (The first kernel built on them - two fixed 8-warp groups per CTA - was superseded by
// the warp-pool kernel in round 1 and has been removed.)
//
//   * operands are the fp16 hi / lo planes the QKV GEMM epilogue wrote; every product runs as the
//     error-compensated triple  hi*hi + hi*lo + lo*hi  (same scheme as gemm_tc.cuh), fp32 accumulate
//   * Q.K^T:   A = Q rows (registers, straight from global), B = K tile via ldmatrix
//   * relative-key term: R = Q . E_window^T is a GEMM over the (n + 15) distance-embedding rows a
//     16-row block can see; the Toeplitz gather S[l, r] += R[l, l - r + c] is done through a per-warp
//     fp32 scratch in shared memory (each row is written and read back by the same warp)
//   * softmax in registers on the S accumulator fragments (quad shuffles for the row max / sum)
//   * P.V:     A = P (S fragments re-packed to fp16 hi / lo in registers), B = V via ldmatrix.trans
//   * output: ctx as fp16 hi / lo planes, ready to be the A operand of the attention-output GEMM
//
// Keys >= n_keys are masked to -inf, which equals the reference's additive -10000 (exp underflows to
// exactly 0 in fp32, sampling.py:56-58 + modelling.py:450-452).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace fd {

constexpr int ATT_WARPS = 16;
constexpr int ATT_PITCH = 32;     // halves per smem row: unpadded 64-byte rows, 16-byte chunks XOR-swizzled (att_sw)
constexpr int ATT_RP = 56;        // fp32 scratch pitch (== 24 mod 32: conflict-free float2 stores; >= 32 + 16 columns)
constexpr int ATT_E_TABLE = 256;  // rows of the padded per-layer table (255 real + 1 zero row)
constexpr int ATT_KV_HALVES = 128 * ATT_PITCH;  // one K or V plane of one work item

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Offset (in halves) of the 16-byte chunk `c` (0..3) of row `r` in a [rows][32 halves] tile.  Rows are 64 bytes,
// unpadded; the chunk index is XOR-ed with bits 1-2 of the row, so the eight rows of one ldmatrix 8x8 matrix
// (same logical chunk, consecutive rows) land in eight different 16-byte bank groups.
__device__ __forceinline__ int att_sw(int r, int c) { return r * ATT_PITCH + ((c ^ ((r >> 1) & 3)) << 3); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (x0, x1) -> packed fp16 hi and lo words
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <bool THREE>
__device__ __forceinline__ void att_load_q(uint32_t (&qa_hi)[2][4], uint32_t (&qa_lo)[2][4],
                                           const __half* __restrict__ qkv_hi, const __half* __restrict__ qkv_lo,
                                           int r0, int l0, int n_rows, int head, int H) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, ld = 3 * H;
  const int ra = min(l0 + g, n_rows - 1), rb = min(l0 + g + 8, n_rows - 1);  // clamp: rows >= n_rows are discarded
  const size_t oa = (size_t)(r0 + ra) * ld + head * FD_HEAD_DIM, ob = (size_t)(r0 + rb) * ld + head * FD_HEAD_DIM;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int c = ks * 16 + 2 * t;
    qa_hi[ks][0] = *reinterpret_cast<const uint32_t*>(qkv_hi + oa + c);
    qa_hi[ks][1] = *reinterpret_cast<const uint32_t*>(qkv_hi + ob + c);
    qa_hi[ks][2] = *reinterpret_cast<const uint32_t*>(qkv_hi + oa + c + 8);
    qa_hi[ks][3] = *reinterpret_cast<const uint32_t*>(qkv_hi + ob + c + 8);
    if (THREE) {
      qa_lo[ks][0] = *reinterpret_cast<const uint32_t*>(qkv_lo + oa + c);
      qa_lo[ks][1] = *reinterpret_cast<const uint32_t*>(qkv_lo + ob + c);
      qa_lo[ks][2] = *reinterpret_cast<const uint32_t*>(qkv_lo + oa + c + 8);
      qa_lo[ks][3] = *reinterpret_cast<const uint32_t*>(qkv_lo + ob + c + 8);
    }
  }
}

// One warp: 16 query rows [l0, l0 + 16) of one (chain, head) against all staged keys.
template <bool THREE>
__device__ __forceinline__ void att_rows(const __half* kv, const __half* Es_hi, const __half* Es_lo, float* Rw,
                                         const float* Bs, const uint32_t (&qa_hi)[2][4], const uint32_t (&qa_lo)[2][4],
                                         int r0, int l0, int n_rows, int n_keys,
                                         int head, int H, __half* __restrict__ ctx_hi, __half* __restrict__ ctx_lo) {
  const __half* Ks_hi = kv; const __half* Ks_lo = kv + ATT_KV_HALVES;
  const __half* Vs_hi = kv + 2 * ATT_KV_HALVES; const __half* Vs_lo = kv + 3 * ATT_KV_HALVES;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nk16 = (n_keys + 15) & ~15;

  // B-operand row address pattern of ldmatrix.x4 over a [8 rows][32 halves] tile: lane -> (row, 8-half column block)
  const int lm_row = lane & 7, lm_c = lane >> 3;

  // ---- S = Q K^T ---------------------------------------------------------------------------------
  float s[16][4];
  const int nnb = nk16 >> 3;  // n8 key blocks
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.0f;
    if (nb < nnb) {
      uint32_t kh[4], kl[4];
      ldsm_x4(kh, Ks_hi + att_sw(nb * 8 + lm_row, lm_c));
      mma_f16(s[nb], qa_hi[0], kh[0], kh[1]);
      mma_f16(s[nb], qa_hi[1], kh[2], kh[3]);
      if (THREE) {  // the correction products get their own accumulator: two independent chains per block
        float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        ldsm_x4(kl, Ks_lo + att_sw(nb * 8 + lm_row, lm_c));
        mma_f16(x, qa_hi[0], kl[0], kl[1]);
        mma_f16(x, qa_lo[0], kh[0], kh[1]);
        mma_f16(x, qa_hi[1], kl[2], kl[3]);
        mma_f16(x, qa_lo[1], kh[2], kh[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[nb][c] += x[c];
      }
    }
  }

  // ---- relative-key term, 32 keys at a time ------------------------------------------------------
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    const int kq = min(32, nk16 - 32 * qq);  // keys in this chunk (16 or 32); <= 0 -> nothing
    if (kq > 0) {
      // column j of the scratch <-> table row e0 + j, where  j = (l - l0) - (r - 32 qq) + (kq - 1)
      const int e0 = l0 - 32 * qq + 128 - kq;
      const int njb = (kq >> 3) + 2;
      for (int jb = 0; jb < njb; ++jb) {
        float r4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t eh[4], el[4];
        ldsm_x4(eh, Es_hi + att_sw(e0 + jb * 8 + lm_row, lm_c));
        mma_f16(r4, qa_hi[0], eh[0], eh[1]);
        mma_f16(r4, qa_hi[1], eh[2], eh[3]);
        if (THREE) {
          float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          ldsm_x4(el, Es_lo + att_sw(e0 + jb * 8 + lm_row, lm_c));
          mma_f16(x, qa_hi[0], el[0], el[1]);
          mma_f16(x, qa_lo[0], eh[0], eh[1]);
          mma_f16(x, qa_hi[1], el[2], el[3]);
          mma_f16(x, qa_lo[1], eh[2], eh[3]);
#pragma unroll
          for (int c = 0; c < 4; ++c) r4[c] += x[c];
        }
        *reinterpret_cast<float2*>(Rw + g * ATT_RP + jb * 8 + 2 * t) = make_float2(r4[0], r4[1]);
        *reinterpret_cast<float2*>(Rw + (g + 8) * ATT_RP + jb * 8 + 2 * t) = make_float2(r4[2], r4[3]);
      }
      __syncwarp();
#pragma unroll
      for (int nbl = 0; nbl < 4; ++nbl) {
        const int nb = qq * 4 + nbl;
        if (nbl * 8 < kq) {
          const int rr = nbl * 8 + 2 * t;  // key index inside the chunk (first of this thread's two)
          const int ca = g - rr + kq - 1, cb = g + 8 - rr + kq - 1;
          s[nb][0] += Rw[g * ATT_RP + ca];
          s[nb][1] += Rw[g * ATT_RP + ca - 1];
          s[nb][2] += Rw[(g + 8) * ATT_RP + cb];
          s[nb][3] += Rw[(g + 8) * ATT_RP + cb - 1];
        }
      }
      __syncwarp();
    }
  }

  // ---- scale, bias / mask, softmax (rows g and g + 8 of this warp's block) -----------------------
  // logits are kept in log2 units: t = s * (log2e / sqrt(32)) + bias * log2e, p = 2^(t - max t); the bias row
  // was staged pre-multiplied by log2e (masked keys: -inf).  t - max <= 0, so the bare ex2.approx is safe.
  const float c_scale = 0.17677669529663688110f * 1.44269504088896340736f;  // log2(e) / sqrt(32)
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    if (nb < nnb) {
      const float2 b = *reinterpret_cast<const float2*>(Bs + nb * 8 + 2 * t);
      s[nb][0] = fmaf(s[nb][0], c_scale, b.x); s[nb][1] = fmaf(s[nb][1], c_scale, b.y);
      s[nb][2] = fmaf(s[nb][2], c_scale, b.x); s[nb][3] = fmaf(s[nb][3], c_scale, b.y);
      m0 = fmaxf(m0, fmaxf(s[nb][0], s[nb][1]));
      m1 = fmaxf(m1, fmaxf(s[nb][2], s[nb][3]));
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float sum0 = 0.0f, sum1 = 0.0f, sum0b = 0.0f, sum1b = 0.0f;  // two partial sums per row: shorter add chains
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    if (nb < nnb) {
      s[nb][0] = ex2_approx(s[nb][0] - m0); s[nb][1] = ex2_approx(s[nb][1] - m0);
      s[nb][2] = ex2_approx(s[nb][2] - m1); s[nb][3] = ex2_approx(s[nb][3] - m1);
      sum0 += s[nb][0]; sum0b += s[nb][1];
      sum1 += s[nb][2]; sum1b += s[nb][3];
    }
  }
  sum0 += sum0b; sum1 += sum1b;
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  // ---- O = P V -----------------------------------------------------------------------------------
  float o[4][4];
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.0f;
  const int nkb = nk16 >> 4;
  // ldmatrix.x4.trans over a [16 keys][16 d] tile: lane -> (key row, d column block)
  const int vt_row = (lane & 7) + ((lane >> 3) & 1) * 8, vt_c = lane >> 4;
#pragma unroll
  for (int kb = 0; kb < 8; ++kb) {
    if (kb < nkb) {
      uint32_t p_hi[4], p_lo[4];
      split2(s[2 * kb][0], s[2 * kb][1], p_hi[0], p_lo[0]);
      split2(s[2 * kb][2], s[2 * kb][3], p_hi[1], p_lo[1]);
      split2(s[2 * kb + 1][0], s[2 * kb + 1][1], p_hi[2], p_lo[2]);
      split2(s[2 * kb + 1][2], s[2 * kb + 1][3], p_hi[3], p_lo[3]);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        uint32_t vh[4], vl[4];
        ldsm_x4_t(vh, Vs_hi + att_sw(kb * 16 + vt_row, dp * 2 + vt_c));
        mma_f16(o[2 * dp], p_hi, vh[0], vh[1]);
        mma_f16(o[2 * dp + 1], p_hi, vh[2], vh[3]);
        if (THREE) {
          ldsm_x4_t(vl, Vs_lo + att_sw(kb * 16 + vt_row, dp * 2 + vt_c));
          mma_f16(o[2 * dp], p_hi, vl[0], vl[1]);
          mma_f16(o[2 * dp + 1], p_hi, vl[2], vl[3]);
          mma_f16(o[2 * dp], p_lo, vh[0], vh[1]);
          mma_f16(o[2 * dp + 1], p_lo, vh[2], vh[3]);
        }
      }
    }
  }

  // ---- normalise and store ctx as hi / lo planes -------------------------------------------------
  const float i0 = 1.0f / sum0, i1 = 1.0f / sum1;
  const int ra = l0 + g, rb = l0 + g + 8;
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) {
    const int c = head * FD_HEAD_DIM + dn * 8 + 2 * t;
    uint32_t hi, lo;
    if (ra < n_rows) {
      split2(o[dn][0] * i0, o[dn][1] * i0, hi, lo);
      *reinterpret_cast<uint32_t*>(ctx_hi + (size_t)(r0 + ra) * H + c) = hi;
      if (THREE) *reinterpret_cast<uint32_t*>(ctx_lo + (size_t)(r0 + ra) * H + c) = lo;
    }
    if (rb < n_rows) {
      split2(o[dn][2] * i1, o[dn][3] * i1, hi, lo);
      *reinterpret_cast<uint32_t*>(ctx_hi + (size_t)(r0 + rb) * H + c) = hi;
      if (THREE) *reinterpret_cast<uint32_t*>(ctx_lo + (size_t)(r0 + rb) * H + c) = lo;
    }
  }
}

}  // namespace fd
