// Tensor-core projection GEMM for sm_100a:  C[M, N] = A[M, K] * W[N, K]^T (+bias)(+resid)(+GELU)
//
//   * tcgen05.mma (kind::f16, fp32 accumulate in TMEM), issued by one elected thread
//   * operands staged in shared memory by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, K-major)
//   * warp-specialised persistent CTAs: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc),
//     warps 2..9 = epilogue (tcgen05.ld -> registers -> bias / residual / GELU -> global); two warps
//     share each TMEM lane quadrant and split the tile's columns
//   * two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i + 1
//   * CTA pairs (cta_group::2, default): the two CTAs of a 2-cluster sit on the two SMs of a TPC and
//     compute ONE 256 x BN tile - each SM stages its own 128 rows of A and only HALF of the weight
//     tile, the MMA (issued by the leader CTA) reads both halves.  ncu on the 1-CTA kernel showed it
//     shared-memory-bandwidth bound (operand reads 107 B/clk + TMA writes 71 + epilogue staging 28
//     against 128 B/clk/SM, tensor pipe capped at ~60%); the pair halves the B-operand traffic per SM
//     and the smaller stage buys a third pipeline stage.
//   * (A/B fallback) independent CTAs in a 2-cluster sharing the weight tile by TMA multicast
//
// Precision.  The parity gate of this project is 1e-4 max-abs on fp32 angle tensors, which single
// pass fp16/bf16/tf32 operands do not meet.  FD_GEMM_TC_3X therefore runs
// an error-compensated split: every fp32 operand x is stored as two fp16 planes
//     hi = fp16(x * 2^s),  lo = fp16(x * 2^s - hi)            (s = 0 for activations; for a weight
//     matrix s is chosen so max|w| * 2^s is in [1024, 2048): keeps `lo` out of fp16 subnormals)
// and the product is accumulated as  hi*hi + hi*lo + lo*hi  in the fp32 TMEM accumulator (the dropped
// lo*lo term is ~2^-22 relative).  The epilogue multiplies by 2^-s (exact).  FD_GEMM_TC_1X issues
// only the hi*hi pass (throughput mode, ~2^-11 relative operands).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels_simt.cuh"  // EPI_* and gelu_erf

namespace fd {

constexpr int TC_BM = 128;         // rows per tile (one UMMA M)
constexpr int TC_BK = 64;          // fp16 elements per 128-byte swizzle row
constexpr int TC_UMMA_K = 16;      // K of one tcgen05.mma.kind::f16
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;  // TMA warp + MMA warp + epilogue warps
constexpr uint32_t TC_TMEM_COLS = 512;

struct TcPlane {         // fp16 hi / lo planes of an activation matrix [rows, k]
  __half* hi = nullptr;
  __half* lo = nullptr;
  int rows = 0, k = 0;
  CUtensorMap map_hi, map_lo;  // box {64, 128}
};
struct TcWeight {        // fp16 hi / lo planes of a weight matrix [n, k], scaled by 2^shift
  __half* hi = nullptr;
  __half* lo = nullptr;
  int n = 0, k = 0, bn = 0;
  float inv_scale = 1.0f;      // 2^-shift
  CUtensorMap map_hi, map_lo;    // box {64, bn}
  CUtensorMap half_hi, half_lo;  // box {64, bn / 2}: what one CTA of a 2-cluster fetches and multicasts
};
struct TcActs {
  TcPlane h, qkv, ctx, a, inter;
};

// LayerNorm folded into the projections (EPI_LNIN*, EPI_LNRES).  BertSelfOutput / BertOutput compute
//     y = LN(v) = (v - m) * r * gamma + beta,   v = dense(x) + residual,   m / r = the row's mean / 1 / sqrt(var + eps)
// and the next projection computes  y W^T + b.  Since  y W^T + b = r * (v W'^T - m * c) + d  with the STATIC
//     W' = W diag(gamma),   c[n] = sum_k gamma[k] W[n, k],   d[n] = sum_k beta[k] W[n, k] + b[n],
// the consuming GEMM can take the RAW rows v as its A operand (hi / lo planes written by the producing GEMM's
// epilogue) and apply the normalisation as a per-row affine map in its own epilogue.  The standalone LayerNorm
// launches (24 per reverse step, 16 bytes of HBM traffic per element each) disappear:
//   producer (EPI_LNRES)   v = acc * s + bias + residual;  writes v (fp32, the later residual source), its hi / lo
//                          planes, and per-row partial sums (sum v, sum v^2) of its column range - one slot per
//                          (column block, epilogue column group), summed by the readers in slot order (deterministic).
//                          The residual is either a plain fp32 tensor (layer 0: the embedding output) or the
//                          normalised value of the PREVIOUS site, recomputed on the fly from that site's v and sums.
//   consumer (EPI_LNIN*)   y = r * (acc * s - m * c[n]) + d[n]   (+ GELU)
struct TcLn {
  const float2* in_stats = nullptr;   // consumer: [rows][in_parts] partial (sum, sum of squares) of the A rows
  const float* in_c = nullptr;        //           c[n]  (d[n] travels as `bias`)
  int in_parts = 0;
  const float* res_v = nullptr;       // producer: residual source rows (fp32) ...
  const float2* res_stats = nullptr;  //           ... normalised with these sums (nullptr: used as they are)
  const float* res_g = nullptr;
  const float* res_b = nullptr;
  int res_parts = 0;
  float2* out_stats = nullptr;        // producer: [rows][2 N / BN]
  float inv_n = 0.0f, eps = 0.0f;     // 1 / hidden, LayerNorm epsilon
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline must end the kernel with an error flag, never hang the GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) return false;  // ~2 s
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// cta_group::2 load: issued by BOTH CTAs of the pair for their own smem; the byte count is reported to
// the LEADER's mbarrier (peer bit of the shared-window address cleared, as CUTLASS SM100_TMA_2SM_LOAD).
constexpr uint64_t TC_EVICT_FIRST = 0x12F0000000000000ull;  // streamed once (activation planes)
constexpr uint64_t TC_EVICT_LAST = 0x14F0000000000000ull;   // re-read by every CTA (weights)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                uint64_t policy) {
  const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// issue only; the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile whose rows are 128 bytes: 8-row groups are 1024 bytes
// apart (SBO), LBO is unused for swizzled K-major layouts (encoded 1), descriptor version 1
// (sm_100), layout type 2 = SWIZZLE_128B.  Fields are in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // version = 1
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = bn.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int bn, int m = TC_BM) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int NPASS, bool PAIR = false>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;                     // 16 KB per plane
  static constexpr int W_BYTES = (PAIR ? BN / 2 : BN) * TC_BK * 2;      // rows of W this CTA stages, 128 B each
  static constexpr int PLANES = NPASS == 1 ? 1 : 2;
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + W_BYTES);
  // epilogue staging: one [32][36] fp32 transpose buffer per epilogue warp
  static constexpr int EPI_PITCH = 36;
  static constexpr int EPI_BYTES = TC_EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int BUDGET = 227 * 1024 - EPI_BYTES - 2048;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
  static_assert(STAGES >= 2, "tile too large");
  static_assert(2 * BN <= (int)TC_TMEM_COLS, "two accumulators must fit TMEM");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
};

// four consecutive residual-source values
// (streaming load: the fp32 rows of a LayerNorm site are read exactly once, by the next site's epilogue, long after
// they have left L2 - they must not displace the hi / lo planes the NEXT kernel is about to read)
__device__ __forceinline__ float4 tc_load_res4(const TcLn& ln, const float* plain, size_t off) {
#ifdef TC_NO_STREAM_HINTS
  return *reinterpret_cast<const float4*>((plain ? plain : ln.res_v) + off);
#else
  return __ldcs(reinterpret_cast<const float4*>((plain ? plain : ln.res_v) + off));
#endif
}

template <int BN, int NPASS, int EPI, int CL, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const float* __restrict__ bias, const float* __restrict__ resid, float* __restrict__ C,
               __half* __restrict__ c_hi, __half* __restrict__ c_lo, int M, int N, int K,
               float out_scale, int* __restrict__ err_flag, unsigned long long a_policy, TcLn ln) {
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of 2");
  using Cfg = TcCfg<BN, NPASS, PAIR>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                          // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + Cfg::STAGES;           // [STAGES]   MMA -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;    // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;             // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // With CL == 2 a "tile" is a pair of vertically adjacent 128-row blocks, one per CTA of the cluster;
  // both CTAs walk the same tile sequence (they are coupled through the multicast barriers).
  const int n_blocks = N / BN, m_blocks = M / (TC_BM * CL);
  const int n_tiles = n_blocks * m_blocks, k_blocks = K / TC_BK;
  const int crank = CL > 1 ? (int)cluster_rank() : 0;
  const int tile0 = blockIdx.x / CL, tile_step = gridDim.x / CL;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  const bool leader = !PAIR || crank == 0;  // in a pair only the leader CTA issues MMAs
  if (threadIdx.x == 0) {
    // PAIR: `full` and `acc_empty` are only used in the leader (they collect both CTAs' bytes / arrivals);
    // `empty` and `acc_full` exist in both CTAs and are signalled by the leader's multicast commits.
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], PAIR ? 1 : CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], (PAIR ? 2 : 1) * 32 * TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) { if (PAIR) tmem_alloc_2sm(tmem_slot, TC_TMEM_COLS); else tmem_alloc(tmem_slot, TC_TMEM_COLS); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_w_hi);
    if (NPASS > 1) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_lo); }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peer barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; A planes / C are the chain's data

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      bool ok = true;
      for (int tile = tile0; tile < n_tiles && ok; tile += tile_step) {
        const int m0 = ((tile / n_blocks) * CL + crank) * TC_BM, n0 = (tile % n_blocks) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          // the stage is free once the MMA warps of ALL CTAs in the cluster have drained it
          if (!mbar_wait(&empty[stage], phase ^ 1)) { atomicExch(err_flag, 101); ok = false; break; }
          uint8_t* s = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* w_hi = s + Cfg::PLANES * Cfg::A_BYTES;
          uint8_t* w_lo = w_hi + Cfg::W_BYTES;
          if (PAIR) {
            // both CTAs stage their own 128 rows of A and their half of the weight tile; every byte is
            // accounted on the leader's barrier, which expects the pair's total
            if (leader) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_2d_2sm(s, &map_a_hi, &full[stage], kb * TC_BK, m0, a_policy);
            tma_load_2d_2sm(w_hi, &map_w_hi, &full[stage], kb * TC_BK, n0 + crank * (BN / 2), TC_EVICT_LAST);
            if (NPASS > 1) {
              tma_load_2d_2sm(s + Cfg::A_BYTES, &map_a_lo, &full[stage], kb * TC_BK, m0, a_policy);
              tma_load_2d_2sm(w_lo, &map_w_lo, &full[stage], kb * TC_BK, n0 + crank * (BN / 2), TC_EVICT_LAST);
            }
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(s, &map_a_hi, &full[stage], kb * TC_BK, m0);
          if (NPASS > 1) tma_load_2d(s + Cfg::A_BYTES, &map_a_lo, &full[stage], kb * TC_BK, m0);
          if (CL == 1) {
            tma_load_2d(w_hi, &map_w_hi, &full[stage], kb * TC_BK, n0);
            if (NPASS > 1) tma_load_2d(w_lo, &map_w_lo, &full[stage], kb * TC_BK, n0);
          } else {  // this CTA's half of the weight tile, delivered to both CTAs (map_w_* has a BN/2-row box)
            const int half_off = crank * (Cfg::W_BYTES / 2);
            tma_load_2d_mc(w_hi + half_off, &map_w_hi, &full[stage], kb * TC_BK, n0 + crank * (BN / 2), kMask);
            if (NPASS > 1)
              tma_load_2d_mc(w_lo + half_off, &map_w_lo, &full[stage], kb * TC_BK, n0 + crank * (BN / 2), kMask);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_f16(BN, PAIR ? 2 * TC_BM : TC_BM);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      bool ok = true;
      for (int tile = tile0; tile < n_tiles && ok; tile += tile_step) {
        if (!mbar_wait(&acc_empty[acc], acc_phase ^ 1)) { atomicExch(err_flag, 102); break; }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          if (!mbar_wait(&full[stage], phase)) { atomicExch(err_flag, 103); ok = false; break; }
          tc_fence_after();
          const uint32_t s = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t a_hi = s, a_lo = s + Cfg::A_BYTES;
          const uint32_t w_hi = s + Cfg::PLANES * Cfg::A_BYTES, w_lo = w_hi + Cfg::W_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            const uint32_t koff = k * TC_UMMA_K * 2;  // bytes inside the 128-byte swizzle row
            const uint64_t da_hi = umma_desc_sw128(a_hi + koff), dw_hi = umma_desc_sw128(w_hi + koff);
            if (PAIR) umma_f16_2sm(d_tmem, da_hi, dw_hi, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, da_hi, dw_hi, idesc, (kb | k) != 0 ? 1u : 0u);
            if (NPASS > 1) {
              const uint64_t da_lo = umma_desc_sw128(a_lo + koff), dw_lo = umma_desc_sw128(w_lo + koff);
              if (PAIR) { umma_f16_2sm(d_tmem, da_hi, dw_lo, idesc, 1u); umma_f16_2sm(d_tmem, da_lo, dw_hi, idesc, 1u); }
              else { umma_f16(d_tmem, da_hi, dw_lo, idesc, 1u); umma_f16(d_tmem, da_lo, dw_hi, idesc, 1u); }
            }
          }
          // frees this smem stage (in every CTA that stages into it) once the MMAs above retire
          if (PAIR) umma_commit_2sm_mc(&empty[stage], kMask);
          else if (CL == 1) umma_commit(&empty[stage]);
          else umma_commit_mc(&empty[stage], kMask);
          if (kb == k_blocks - 1) { if (PAIR) umma_commit_2sm_mc(&acc_full[acc], kMask); else umma_commit(&acc_full[acc]); }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // TMEM hands every thread one ROW of the accumulator (lane = row).  Writing global memory in
    // that shape touches 32 different cache lines per instruction (ncu: the kernel was bound by it,
    // long-scoreboard stalls behind scattered STG/LDG).  So each 32 x 32 chunk is transposed through
    // a per-warp shared-memory buffer: afterwards 8 consecutive lanes cover 128 contiguous bytes of
    // one row, and the bias / residual / GELU / fp16 hi-lo split run in that coalesced layout.
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
    constexpr int COLS_PER_WARP = BN / (TC_EPI_WARPS / 4);
    static_assert(COLS_PER_WARP % 32 == 0, "column split of the epilogue warps");
    constexpr int NCH = COLS_PER_WARP / 32;
    constexpr int EP = Cfg::EPI_PITCH;
    const int col_lo = ((warp - 2) >> 2) * COLS_PER_WARP;
    float* stg = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256) + (warp - 2) * 32 * EP;
    const int lr = lane >> 3, lc = (lane & 7) * 4;  // coalesced layout: row inside a 4-row group, first column
    constexpr bool LNIN = EPI == EPI_LNIN || EPI == EPI_LNIN_GELU;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      const int m0 = ((tile / n_blocks) * CL + crank) * TC_BM, n0 = (tile % n_blocks) * BN;
      const int row0 = m0 + quad * 32 + lr;  // this thread's rows: row0 + 4 i, i = 0..7
      // per-row LayerNorm statistics, fetched before the accumulator is waited for (independent of it)
      float mu[8], rs_[8];      // consumer: mean / rstd of the A rows;  producer: of the residual source rows
      float s1[8], s2[8];       // producer: partial sums of this thread's output columns
      if (LNIN || (EPI == EPI_LNRES && ln.res_stats)) {
        const float2* st = LNIN ? ln.in_stats : ln.res_stats;
        const int parts = LNIN ? ln.in_parts : ln.res_parts;
        // all loads of the 8 rows are issued before the first use (a loop that accumulates as it loads is a chain of
        // dependent L2 round trips: measured 13 us per tile); 2 or 4 partials per row = one or two 16-byte loads
        float a1[8], a2[8];
        if (parts == 4) {
          float4 u[8], w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4* q = reinterpret_cast<const float4*>(st + (size_t)(row0 + 4 * i) * 4);
            u[i] = q[0]; w[i] = q[1];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { a1[i] = ((u[i].x + u[i].z) + w[i].x) + w[i].z; a2[i] = ((u[i].y + u[i].w) + w[i].y) + w[i].w; }
        } else if (parts == 2) {
          float4 u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = *reinterpret_cast<const float4*>(st + (size_t)(row0 + 4 * i) * 2);
#pragma unroll
          for (int i = 0; i < 8; ++i) { a1[i] = u[i].x + u[i].z; a2[i] = u[i].y + u[i].w; }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            a1[i] = 0.0f; a2[i] = 0.0f;
            for (int p = 0; p < parts; ++p) {
              const float2 v2 = st[(size_t)(row0 + 4 * i) * parts + p];
              a1[i] += v2.x; a2[i] += v2.y;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float mean = a1[i] * ln.inv_n;
          mu[i] = mean;
          rs_[i] = 1.0f / sqrtf(fmaxf(fmaf(-mean, mean, a2[i] * ln.inv_n), 0.0f) + ln.eps);
        }
      }
      if (EPI == EPI_LNRES) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1[i] = 0.0f; s2[i] = 0.0f; }
      }
      // Two cheap pieces of software pipelining (measured on one box, profiles/r02_experiments.md: -1.1 % per pass):
      //   COLPREF   the 16-byte bias / c vectors of chunk k + 1 are loaded while chunk k is processed (LNIN variants only:
      //             the LNRES variant sits at the 168-register cap - 10 warps put 3 on one SM sub-partition - and spills)
      //   RESEARLY  the first chunk's residual rows go out before the accumulator is waited for (independent of it)
      // A full cross-tile pipeline of every epilogue load (residual rows, statistics, vectors) was built and measured
      // SLOWER for the LNIN variants (+7 %) and only 1 - 3 % faster for LNRES; it is not kept.
      constexpr bool COLPREF = LNIN;
      constexpr bool RESEARLY = true;
      float4 nb4 = make_float4(0.f, 0.f, 0.f, 0.f), nc4 = nb4;
      if (COLPREF) {
        nb4 = *reinterpret_cast<const float4*>(bias + n0 + col_lo + lc);
        nc4 = *reinterpret_cast<const float4*>(ln.in_c + n0 + col_lo + lc);
      }
      float4 rs[8];  // residual rows of the chunk being processed (issued one chunk ahead)
      if (RESEARLY && (EPI == EPI_BIAS_RESID || EPI == EPI_LNRES)) {  // independent of the accumulator: before the wait
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rs[i] = tc_load_res4(ln, EPI == EPI_LNRES ? nullptr : resid, (size_t)(row0 + 4 * i) * N + n0 + col_lo + lc);
      }
      if (!mbar_wait(&acc_full[acc], acc_phase)) { if (lane == 0) atomicExch(err_flag, 104); break; }
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      uint32_t v[32];
      tmem_ld32_issue(t_row + (uint32_t)col_lo, v);
      if (!RESEARLY && (EPI == EPI_BIAS_RESID || EPI == EPI_LNRES)) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rs[i] = tc_load_res4(ln, EPI == EPI_LNRES ? nullptr : resid, (size_t)(row0 + 4 * i) * N + n0 + col_lo + lc);
      }
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci) {
        const int c0 = col_lo + ci * 32;
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * EP + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (ci + 1 < NCH) {
          tmem_ld32_issue(t_row + (uint32_t)(c0 + 32), v);  // next chunk streams in under this chunk's math
        } else {
          tc_fence_before();  // every TMEM read of this accumulator has completed
          if (PAIR && !leader) mbar_arrive_remote(&acc_empty[acc], 0);  // the leader's MMA warp owns the barrier
          else mbar_arrive(&acc_empty[acc]);
        }
        float4 b4, c4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = c4, e4 = c4;
        if (COLPREF) {  // this chunk's vectors were loaded one chunk ago; the next chunk's go out now
          b4 = nb4; c4 = nc4;
          if (ci + 1 < NCH) {
            nb4 = *reinterpret_cast<const float4*>(bias + n0 + c0 + 32 + lc);
            nc4 = *reinterpret_cast<const float4*>(ln.in_c + n0 + c0 + 32 + lc);
          }
        } else {
          b4 = *reinterpret_cast<const float4*>(bias + n0 + c0 + lc);
          if (LNIN) c4 = *reinterpret_cast<const float4*>(ln.in_c + n0 + c0 + lc);
          if (EPI == EPI_LNRES && ln.res_stats) {
            g4 = *reinterpret_cast<const float4*>(ln.res_g + n0 + c0 + lc);
            e4 = *reinterpret_cast<const float4*>(ln.res_b + n0 + c0 + lc);
          }
        }
        float4 rcur[8];
        if (EPI == EPI_BIAS_RESID || EPI == EPI_LNRES) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rcur[i] = rs[i];
          if (ci + 1 < NCH) {  // the next chunk's residual rows travel under this chunk's math and stores
#pragma unroll
            for (int i = 0; i < 8; ++i)
              rs[i] = tc_load_res4(ln, EPI == EPI_LNRES ? nullptr : resid, (size_t)(row0 + 4 * i) * N + n0 + c0 + 32 + lc);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + lr;
          const float4 a = *reinterpret_cast<const float4*>(stg + r * EP + lc);
          const size_t off = (size_t)(m0 + quad * 32 + r) * N + n0 + c0 + lc;
          float o0, o1, o2, o3;
          if (LNIN) {  // y = rstd * (acc * s - mean * c[n]) + d[n]
            o0 = fmaf(rs_[i], fmaf(-mu[i], c4.x, a.x * out_scale), b4.x); o1 = fmaf(rs_[i], fmaf(-mu[i], c4.y, a.y * out_scale), b4.y);
            o2 = fmaf(rs_[i], fmaf(-mu[i], c4.z, a.z * out_scale), b4.z); o3 = fmaf(rs_[i], fmaf(-mu[i], c4.w, a.w * out_scale), b4.w);
          } else {
            o0 = fmaf(a.x, out_scale, b4.x); o1 = fmaf(a.y, out_scale, b4.y);
            o2 = fmaf(a.z, out_scale, b4.z); o3 = fmaf(a.w, out_scale, b4.w);
          }
          if (EPI == EPI_BIAS_RESID) { o0 += rcur[i].x; o1 += rcur[i].y; o2 += rcur[i].z; o3 += rcur[i].w; }
          if (EPI == EPI_LNRES) {
            float4 rr = rcur[i];
            if (ln.res_stats) {  // the previous site's LayerNorm output, recomputed from its raw rows
              rr.x = ((rr.x - mu[i]) * rs_[i]) * g4.x + e4.x; rr.y = ((rr.y - mu[i]) * rs_[i]) * g4.y + e4.y;
              rr.z = ((rr.z - mu[i]) * rs_[i]) * g4.z + e4.z; rr.w = ((rr.w - mu[i]) * rs_[i]) * g4.w + e4.w;
            }
            o0 += rr.x; o1 += rr.y; o2 += rr.z; o3 += rr.w;
            s1[i] += (o0 + o1) + (o2 + o3);
            s2[i] = fmaf(o0, o0, fmaf(o1, o1, fmaf(o2, o2, fmaf(o3, o3, s2[i]))));
          }
          if (EPI == EPI_BIAS_GELU || EPI == EPI_LNIN_GELU) { o0 = gelu_erf_fast(o0); o1 = gelu_erf_fast(o1); o2 = gelu_erf_fast(o2); o3 = gelu_erf_fast(o3); }
#ifdef TC_NO_STREAM_HINTS
          if (C) *reinterpret_cast<float4*>(C + off) = make_float4(o0, o1, o2, o3);
#else
          if (C) __stcs(reinterpret_cast<float4*>(C + off), make_float4(o0, o1, o2, o3));  // read once, three kernels later
#endif
          if (c_hi) {
            const __half2 h01 = __floats2half2_rn(o0, o1), h23 = __floats2half2_rn(o2, o3);
            uint2 ph;
            ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
            *reinterpret_cast<uint2*>(c_hi + off) = ph;
            if (c_lo) {
              const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
              const __half2 l01 = __floats2half2_rn(o0 - f01.x, o1 - f01.y), l23 = __floats2half2_rn(o2 - f23.x, o3 - f23.y);
              uint2 pl;
              pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
              *reinterpret_cast<uint2*>(c_lo + off) = pl;
            }
          }
        }
        __syncwarp();  // staging buffer is free for the next chunk
      }
      if (EPI == EPI_LNRES) {  // the 8 lanes that share a row combine; one slot per (column block, column group)
        const int parts = (N / BN) * (TC_EPI_WARPS / 4);
        const int slot = (n0 / BN) * (TC_EPI_WARPS / 4) + ((warp - 2) >> 2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
            s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
          }
          if ((lane & 7) == 0) ln.out_stats[(size_t)(row0 + 4 * i) * parts + slot] = make_float2(s1[i], s2[i]);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, TC_TMEM_COLS); else tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
__global__ void tc_split_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                __half* __restrict__ lo, size_t n4, float scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    const float x[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
    __half h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2half_rn(x[j]);
      l[j] = __float2half_rn(x[j] - __half2float(h[j]));
    }
    uint2 ph, pl;
    ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
    reinterpret_cast<uint2*>(hi)[i] = ph;
    if (lo) reinterpret_cast<uint2*>(lo)[i] = pl;
  }
}

__global__ void tc_absmax_kernel(const float* __restrict__ src, size_t n, unsigned int* out_bits,
                                 const float* __restrict__ col_gamma, int k) {
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(col_gamma ? src[i] * col_gamma[i % k] : src[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order as uints
}

// Accumulation de-bias (see tc_rz() below).  A weight matrix [n, k] is split into its fp16 hi / lo planes AFTER
// every 16-column chunk c of K has been scaled by  1 + alpha + beta * (k/16 - c): the tensor core's fp32
// accumulator truncates (round toward zero) at every tcgen05.mma, so the contribution of the chunk issued at
// accumulate step s of S shrinks by ~ eps * (S - s + 1); the shrink is a linear functional of the per-chunk
// products and the (static) weight operand can carry its inverse.  Rows < q_rows (the query rows of the fused QKV
// weight) additionally carry the same correction for the two 16-wide chunks of the attention kernel's K = 32
// products Q K^T and Q E^T (head dim d: chunk d / 16 of 2).  Create-time only: plain fp64 arithmetic.
__global__ void tc_split_weight_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                       int n, int k, double scale, double alpha, double beta, double beta_att, int q_rows,
                                       const float* __restrict__ col_gamma) {
  const size_t total = (size_t)n * k;
  const int nk = k / 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / k), col = (int)(i % k);
    double c = 1.0 + alpha + beta * (double)(nk - col / 16);
    if (row < q_rows) c *= 1.0 + alpha + beta_att * (double)(2 - (row % FD_HEAD_DIM) / 16);
    if (col_gamma) c *= (double)col_gamma[col];  // LayerNorm folded into the projection: W' = W diag(gamma) (TcLn)
    const double x = (double)src[i] * scale * c;
    const __half h = __double2half(x);
    hi[i] = h;
    lo[i] = __double2half(x - (double)__half2float(h));
  }
}
// Epilogue vectors of a projection whose input LayerNorm is folded in (TcLn):
//   c[n] = s_n * sum_k gamma[k] W[n, k],   d[n] = s_n * (sum_k beta[k] W[n, k] + b[n])
// (gamma == nullptr: no fold - c = 0, d = s_n * b).  s_n is the attention de-bias factor of the query rows (n < q_rows).
// One warp per output row, fp64 accumulation; create time only.
__global__ void tc_fold_vectors_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                                       const float* __restrict__ beta_ln, int n, int k, int q_rows, double alpha, double beta_att,
                                       float* __restrict__ c_out, float* __restrict__ d_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double sc = 0.0, sd = 0.0;
  if (gamma) {
    for (int j = lane; j < k; j += 32) {
      const double wv = (double)w[(size_t)row * k + j];
      sc += (double)gamma[j] * wv;
      sd += (double)beta_ln[j] * wv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sc += __shfl_xor_sync(0xffffffffu, sc, o);
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
  }
  if (lane == 0) {
    const double s = row < q_rows ? 1.0 + alpha + beta_att * (double)(2 - (row % FD_HEAD_DIM) / 16) : 1.0;
    c_out[row] = (float)(s * sc);
    d_out[row] = (float)(s * (sd + (double)b[row]));
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);

inline tc_encode_fn tc_encoder() {
  static tc_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tc_encode_fn>(p);
  }
  return fn;
}

// [rows, k] fp16 row-major -> TMA map with box {64 (k), box_rows}, 128-byte swizzle.
inline int tc_make_map(CUtensorMap* map, const __half* base, int rows, int k, int box_rows) {
  tc_encode_fn enc = tc_encoder();
  if (!enc) return 1;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

inline int tc_pick_bn(int n) {
  if (n % 192 == 0) return 192;
  if (n % 128 == 0) return 128;
  return 64;
}

inline int tc_alloc_plane(TcPlane* p, int rows, int k) {
  p->rows = rows; p->k = k;
  if (cudaMalloc(&p->hi, sizeof(__half) * (size_t)rows * k) != cudaSuccess) return 1;
  if (cudaMalloc(&p->lo, sizeof(__half) * (size_t)rows * k) != cudaSuccess) return 1;
  cudaMemset(p->hi, 0, sizeof(__half) * (size_t)rows * k);
  cudaMemset(p->lo, 0, sizeof(__half) * (size_t)rows * k);
  if (tc_make_map(&p->map_hi, p->hi, rows, k, TC_BM)) return 2;
  if (tc_make_map(&p->map_lo, p->lo, rows, k, TC_BM)) return 2;
  return 0;
}
inline void tc_free_plane(TcPlane* p) {
  cudaFree(p->hi); cudaFree(p->lo);
  p->hi = p->lo = nullptr;
}
inline int tc_alloc_acts(TcActs* a, int rows_pad, int hidden, int inter) {
  if (tc_alloc_plane(&a->h, rows_pad, hidden)) return 1;
  if (tc_alloc_plane(&a->qkv, rows_pad, 3 * hidden)) return 1;
  if (tc_alloc_plane(&a->ctx, rows_pad, hidden)) return 1;
  if (tc_alloc_plane(&a->a, rows_pad, hidden)) return 1;
  if (tc_alloc_plane(&a->inter, rows_pad, inter)) return 1;
  return 0;
}
inline void tc_free_acts(TcActs* a) {
  tc_free_plane(&a->h); tc_free_plane(&a->qkv); tc_free_plane(&a->ctx); tc_free_plane(&a->a); tc_free_plane(&a->inter);
}

// Tensor-core accumulation de-bias constants (tc_split_weight_kernel; attention P split).
// tools/rz_calib.py (B200, profiles/r02_rz_calibration.md): with same-sign operands a 16-wide K chunk of the 3-pass
// product loses 2.8e-7 of the running sum per chunk still to come (alignment truncation of every product plus the
// truncation of the sum); with zero-mean products - the network's case - only the part that is a linear functional
// of the partial sums survives on average, ~1e-7 per chunk, and that is what a static pre-scale can cancel.
// tools/rz_sweep.py then picked beta on the real network: forward error against the fp64 oracle (production shape)
// rms 1.03e-6 at beta = 0 -> 2.41e-7 for beta in [0.95e-7, 1.05e-7], the fp32 CUDA-core path's own 2.37e-7.
// FOLDINGDIFF_B200_RZ="alpha,beta[,beta_attention]" overrides ("0,0" = off).
struct TcRz { double alpha, beta, beta_att; };
inline TcRz tc_rz() {
  static TcRz rz = [] {
    TcRz r{0.0, 1.0e-7, 1.0e-7};
    const char* e = getenv("FOLDINGDIFF_B200_RZ");
    if (e && e[0]) {
      char* end = nullptr;
      const double a = strtod(e, &end);
      if (end && *end == ',') {
        r.alpha = a;
        r.beta = r.beta_att = strtod(end + 1, &end);
        if (end && *end == ',') r.beta_att = strtod(end + 1, nullptr);
      }
    }
    return r;
  }();
  return rz;
}

// fp32 [n, k] device weight -> scaled fp16 hi / lo planes + TMA maps.  Synchronous (create time).
// q_rows: leading rows that are attention queries (fused QKV weight), see tc_split_weight_kernel.
// col_gamma: LayerNorm weight folded into the K columns (TcLn), or nullptr.
inline int tc_pack_weight(const float* w_dev, int n, int k, TcWeight* out, int q_rows = 0, const float* col_gamma = nullptr) {
  out->n = n; out->k = k; out->bn = tc_pick_bn(n);
  const size_t cnt = (size_t)n * k;
  unsigned int* bits = nullptr;
  if (cudaMalloc(&bits, sizeof(unsigned int)) != cudaSuccess) return 1;
  cudaMemset(bits, 0, sizeof(unsigned int));
  tc_absmax_kernel<<<64, 256>>>(w_dev, cnt, bits, col_gamma, k);
  unsigned int hb = 0;
  if (cudaMemcpy(&hb, bits, sizeof(hb), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(bits); return 1; }
  cudaFree(bits);
  float amax;
  memcpy(&amax, &hb, sizeof(amax));
  int shift = 0;
  if (amax > 0.0f && amax < 3.0e38f) {
    int e;
    frexpf(amax, &e);       // amax = f * 2^e, f in [0.5, 1)
    shift = 11 - e;         // amax * 2^shift in [1024, 2048)
    if (shift > 24) shift = 24;
    if (shift < -4) shift = -4;
  }
  out->inv_scale = ldexpf(1.0f, -shift);
  if (cudaMalloc(&out->hi, sizeof(__half) * cnt) != cudaSuccess) return 1;
  if (cudaMalloc(&out->lo, sizeof(__half) * cnt) != cudaSuccess) return 1;
  const TcRz rz = tc_rz();
  tc_split_weight_kernel<<<256, 256>>>(w_dev, out->hi, out->lo, n, k, ldexp(1.0, shift), rz.alpha, rz.beta, rz.beta_att, q_rows, col_gamma);
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (tc_make_map(&out->map_hi, out->hi, n, k, out->bn)) return 2;
  if (tc_make_map(&out->map_lo, out->lo, n, k, out->bn)) return 2;
  if (tc_make_map(&out->half_hi, out->hi, n, k, out->bn / 2)) return 2;
  if (tc_make_map(&out->half_lo, out->lo, n, k, out->bn / 2)) return 2;
  return 0;
}
inline void tc_free_weight(TcWeight* w) {
  cudaFree(w->hi); cudaFree(w->lo);
  w->hi = w->lo = nullptr;
}

// fp32 activations -> fp16 hi (/ lo) planes.
inline void tc_split(const float* src, TcPlane* dst, int rows, int cols, int mode, cudaStream_t st) {
  const size_t n4 = (size_t)rows * cols / 4;
  const int blocks = (int)((n4 + 255) / 256 > 148 * 8 ? 148 * 8 : (n4 + 255) / 256);
  tc_split_kernel<<<blocks, 256, 0, st>>>(src, dst->hi, mode == 1 /*FD_GEMM_TC_3X*/ ? dst->lo : nullptr, n4, 1.0f);
}

inline std::mutex& tc_cfg_mutex() {  // guards the per-device one-time state below (handles may live on several threads)
  static std::mutex m;
  return m;
}
inline int* tc_err_flag() {  // one flag per device (a process may own several handles on several GPUs)
  static int* flags[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(tc_cfg_mutex());
  if (!flags[dev]) {
    if (cudaMalloc(&flags[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flags[dev], 0, sizeof(int));
  }
  return flags[dev];
}
// max-dynamic-smem is a per-device function attribute: remember on which devices it has been set
inline bool tc_need_configure(unsigned long long* seen_mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(tc_cfg_mutex());
  const unsigned long long bit = 1ull << (dev & 63);
  if (*seen_mask & bit) return false;
  *seen_mask |= bit;
  return true;
}
// Reads (and clears) the device-side pipeline error flag; synchronises the device.
inline int tc_check_error() {
  int* f = tc_err_flag();
  int v = 0;
  if (!f) return -1;
  if (cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  if (v) cudaMemset(f, 0, sizeof(int));
  return v;
}

// FOLDINGDIFF_B200_TC_MODE: "pair" (default: cta_group::2 CTA pairs), "mcast" (independent CTAs of a
// 2-cluster sharing W by TMA multicast), "single" (no cluster).  A/B switch for profiling.
inline int tc_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FOLDINGDIFF_B200_TC_MODE");
    mode = 2;
    if (e && e[0] == 'm') mode = 1;
    if (e && e[0] == 's') mode = 0;
  }
  return mode;
}

template <int BN, int NPASS, int EPI, int CL, bool PAIR>
int tc_launch_cl(const TcPlane* a, const TcWeight* w, const float* bias, const float* resid, float* C,
                 TcPlane* c_tc, int M, int N, int K, int sm_count, cudaStream_t st, const TcLn& ln) {
  using Cfg = TcCfg<BN, NPASS, PAIR>;
  static unsigned long long configured = 0;
  auto kern = tc_gemm_kernel<BN, NPASS, EPI, CL, PAIR>;
  if (tc_need_configure(&configured) &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) return 10;
  int* err = tc_err_flag();
  if (!err) return 11;
  const int tiles = (M / (TC_BM * CL)) * (N / BN);           // cluster-level tiles
  int clusters = sm_count / CL;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // common.cuh: PDL contract
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  __half* c_hi = c_tc ? c_tc->hi : nullptr;
  __half* c_lo = (c_tc && NPASS > 1) ? c_tc->lo : nullptr;
  const CUtensorMap& wh = CL > 1 ? w->half_hi : w->map_hi;
  const CUtensorMap& wl = CL > 1 ? w->half_lo : w->map_lo;
  // the A row block is streamed: evict-first keeps it from displacing the output in L2 - but only when few
  // column tiles re-read it (measured: hurts the 6-tile QKV projection, helps the 2-tile ones)
  const unsigned long long a_policy = (N / BN <= 2) ? TC_EVICT_FIRST : 0x1000000000000000ull;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a->map_hi, a->map_lo, wh, wl, bias, resid, C, c_hi, c_lo, M, N, K,
                                     w->inv_scale, err, a_policy, ln);
  return e == cudaSuccess ? 0 : 12;
}

template <int BN, int NPASS, int EPI>
int tc_launch(const TcPlane* a, const TcWeight* w, const float* bias, const float* resid, float* C,
              TcPlane* c_tc, int M, int N, int K, int sm_count, cudaStream_t st, const TcLn& ln) {
  if (M % (2 * TC_BM) == 0 && tc_mode() == 2)
    return tc_launch_cl<BN, NPASS, EPI, 2, true>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
  if (M % (2 * TC_BM) == 0 && tc_mode() == 1)
    return tc_launch_cl<BN, NPASS, EPI, 2, false>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
  return tc_launch_cl<BN, NPASS, EPI, 1, false>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
}

template <int BN, int NPASS>
int tc_dispatch_epi(int epi, const TcPlane* a, const TcWeight* w, const float* bias, const float* resid,
                    float* C, TcPlane* c_tc, int M, int N, int K, int sm_count, cudaStream_t st, const TcLn& ln) {
  switch (epi) {
    case EPI_BIAS: return tc_launch<BN, NPASS, EPI_BIAS>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
    case EPI_BIAS_GELU: return tc_launch<BN, NPASS, EPI_BIAS_GELU>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
    case EPI_LNIN: return tc_launch<BN, NPASS, EPI_LNIN>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
    case EPI_LNIN_GELU: return tc_launch<BN, NPASS, EPI_LNIN_GELU>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
    case EPI_LNRES: return tc_launch<BN, NPASS, EPI_LNRES>(a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
  }
  return 22;  // EPI_BIAS_RESID belongs to the fp32 CUDA-core GEMM only
}

// mode: 1 = FD_GEMM_TC_3X, 2 = FD_GEMM_TC_1X
inline int tc_gemm(int mode, int epi, const TcPlane* a, const TcWeight* w, const float* bias,
                   const float* resid, float* C, TcPlane* c_tc, int M, int N, int K, int sm_count,
                   cudaStream_t st, const TcLn& ln = TcLn()) {
  if (!a || !w || M % TC_BM || K % TC_BK || N % w->bn || a->k != K || w->k != K || w->n != N) return 20;
  const bool three = mode == 1;
  switch (w->bn) {
    case 192:
      return three ? tc_dispatch_epi<192, 3>(epi, a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln)
                   : tc_dispatch_epi<192, 1>(epi, a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
    case 128:
      return three ? tc_dispatch_epi<128, 3>(epi, a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln)
                   : tc_dispatch_epi<128, 1>(epi, a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
    case 64:
      return three ? tc_dispatch_epi<64, 3>(epi, a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln)
                   : tc_dispatch_epi<64, 1>(epi, a, w, bias, resid, C, c_tc, M, N, K, sm_count, st, ln);
  }
  return 21;
}

}  // namespace fd
