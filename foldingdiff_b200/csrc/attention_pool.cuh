// Relative-key attention, warp-pool scheduling (the default tensor-core attention kernel).
//
// Same math as attention_mma.cuh (att_rows: one warp = 16 query rows of one (chain, head) against all keys), but
// the 16 warps of the persistent CTA are no longer tied to an 8-warp group per work item.  A chain of n residues
// has ceil(n / 16) row blocks; with lengths 50..127 a fixed 8-warp group keeps on average 6 of its 8 warps busy
// (ncu on the grouped kernel: a quarter of the warp slots idle).  Here:
//
//   * K / V of an item (chain, head) arrive by TMA (one 128-row x 64-byte box per plane, SWIZZLE_64B == the
//     att_sw layout ldmatrix wants) into a ring of 4 shared-memory slots, completion on an mbarrier per slot;
//   * row blocks are handed out one at a time from a CTA-wide ticket word (slot sequence number, block index),
//     so every warp always has work while any block of any resident item is left;
//   * the warp that finishes the LAST block of a slot refills that slot with the item 4 places further down
//     the CTA's list (bias row + descriptor by the warp, four TMA loads by its lane 0) - no producer warp, no
//     CTA- or group-wide barrier in the steady state.
//
// Rows past n_keys in a slot hold whatever follows the chain in the packed qkv plane (finite values, or zeros
// past the end of the tensor): their logits are masked to -inf by the bias row and their P is exactly 0.
#pragma once
#include "attention_mma.cuh"
#include "gemm_tc.cuh"

namespace fd {

constexpr int ATTP_NB = 4;                           // ring slots
constexpr int ATTP_SLOT_HALVES = 4 * ATT_KV_HALVES;  // {K hi, K lo, V hi, V lo}, 8 KB each

struct AttpSlot {  // descriptor of the item resident (or arriving) in a ring slot; read with one 8-byte LDS
  int r0;        // first packed row of the chain
  uint32_t dims; // n_rows | n_keys << 8 | row blocks << 16 | head << 24
};
struct AttpCtl {
  unsigned long long full[ATTP_NB];  // mbarriers: TMA bytes of the slot have landed
  AttpSlot slot[ATTP_NB];
  int seq[ATTP_NB];                  // sequence number of the item each slot holds (-1: none yet)
  int done[ATTP_NB];                 // row blocks of that item finished so far
  unsigned int ticket;               // (slot sequence number << 8) | next row block
  int pad[3];
};

constexpr size_t attp_smem_bytes() {
  return (size_t)2 * ATT_E_TABLE * ATT_PITCH * 2     // E hi / lo
         + (size_t)ATTP_NB * ATTP_SLOT_HALVES * 2    // K / V ring
         + (size_t)ATT_WARPS * 16 * ATT_RP * 4       // R scratch
         + (size_t)ATTP_NB * 128 * 4                 // bias rows
         + sizeof(AttpCtl) + 1024;                   // control block + slack for the 1024-byte alignment
}

// Shared-state reads go through explicit ld.volatile.shared (LDS): a volatile load through a generic pointer
// compiles to LD.E.STRONG.SYS, which ncu showed as the hottest stall of the first version of this kernel.
__device__ __forceinline__ uint32_t lds_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_u64(const void* p) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(void* p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// Bring item number `s` of this CTA's list into ring slot s % ATTP_NB.  Called by ONE whole warp.
template <bool THREE>
__device__ __forceinline__ void attp_issue(AttpCtl* ctl, __half* ring, float* bias_rows, const CUtensorMap* map_hi,
                                           const CUtensorMap* map_lo, const int* __restrict__ row_start,
                                           const int* __restrict__ n_rows_arr, const int* __restrict__ n_keys_arr,
                                           const float* __restrict__ key_bias, int n_pad, int H, int heads,
                                           int n_items, int s) {
  const int lane = threadIdx.x & 31, b = s % ATTP_NB;
  const int item = n_items - 1 - ((int)blockIdx.x + s * (int)gridDim.x);  // batches are length-sorted: long first
  const int chain = item / heads, head = item % heads;
  const int r0 = row_start[chain], nr = n_rows_arr[chain], nk = n_keys_arr[chain];
  float* Bs = bias_rows + b * 128;
  for (int i = lane; i < 128; i += 32)  // log2 units, like the logits; keys >= n_keys are masked
    Bs[i] = (i < nk) ? (key_bias ? key_bias[(size_t)chain * n_pad + i] * 1.44269504088896340736f : 0.0f) : -INFINITY;
  if (lane == 0) {
    ctl->slot[b].r0 = r0;
    ctl->slot[b].dims = (uint32_t)nr | (uint32_t)nk << 8 | (uint32_t)((nr + 15) >> 4) << 16 | (uint32_t)head << 24;
    ctl->done[b] = 0;
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the slot vs the TMA writes
    uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[b]);
    __half* dst = ring + (size_t)b * ATTP_SLOT_HALVES;
    mbar_expect_tx(bar, (THREE ? 4u : 2u) * ATT_KV_HALVES * 2u);
    tma_load_2d(dst, map_hi, bar, H + head * FD_HEAD_DIM, r0);
    tma_load_2d(dst + 2 * ATT_KV_HALVES, map_hi, bar, 2 * H + head * FD_HEAD_DIM, r0);
    if (THREE) {
      tma_load_2d(dst + ATT_KV_HALVES, map_lo, bar, H + head * FD_HEAD_DIM, r0);
      tma_load_2d(dst + 3 * ATT_KV_HALVES, map_lo, bar, 2 * H + head * FD_HEAD_DIM, r0);
    }
    __threadfence_block();
    sts_u32(&ctl->seq[b], (uint32_t)s);  // publish: descriptor, bias row and barrier phase are set
  }
  __syncwarp();
}

template <bool THREE>
__global__ void __launch_bounds__(ATT_WARPS * 32, 1)
attention_pool_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                      const __half* __restrict__ qkv_hi, const __half* __restrict__ qkv_lo,
                      const int* __restrict__ row_start, const int* __restrict__ n_rows_arr,
                      const int* __restrict__ n_keys_arr, const float* __restrict__ key_bias, int n_pad,
                      const __half* __restrict__ e_hi, const __half* __restrict__ e_lo, int H, int heads,
                      int n_items, __half* __restrict__ ctx_hi, __half* __restrict__ ctx_lo,
                      int* __restrict__ err_flag) {
  extern __shared__ uint8_t attp_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attp_raw) + 1023) & ~(uintptr_t)1023);
  __half* ring = reinterpret_cast<__half*>(base);                    // 1024-aligned: TMA swizzle atoms
  __half* Es_hi = ring + (size_t)ATTP_NB * ATTP_SLOT_HALVES;
  __half* Es_lo = Es_hi + ATT_E_TABLE * ATT_PITCH;
  float* Rs = reinterpret_cast<float*>(Es_lo + ATT_E_TABLE * ATT_PITCH);
  float* bias_rows = Rs + ATT_WARPS * 16 * ATT_RP;
  AttpCtl* ctl = reinterpret_cast<AttpCtl*>(bias_rows + ATTP_NB * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_slots = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA
  if (tid == 0) {
    for (int b = 0; b < ATTP_NB; ++b) {
      mbar_init(reinterpret_cast<uint64_t*>(&ctl->full[b]), 1);
      ctl->seq[b] = -1; ctl->done[b] = 0;
    }
    ctl->ticket = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&map_hi);
    if (THREE) tma_prefetch_desc(&map_lo);
  }
  for (int i = tid; i < ATT_E_TABLE * 4; i += ATT_WARPS * 32) {
    const int r = i >> 2;
    cp_async16(Es_hi + att_sw(r, i & 3), e_hi + (size_t)r * FD_HEAD_DIM + (i & 3) * 8);
    if (THREE) cp_async16(Es_lo + att_sw(r, i & 3), e_lo + (size_t)r * FD_HEAD_DIM + (i & 3) * 8);
  }
  cp_async_wait_all();
  __syncthreads();
  if (warp < ATTP_NB && warp < n_slots)
    attp_issue<THREE>(ctl, ring, bias_rows, &map_hi, &map_lo, row_start, n_rows_arr, n_keys_arr, key_bias, n_pad, H,
                      heads, n_items, warp);

  float* Rw = Rs + warp * 16 * ATT_RP;
  for (;;) {
    // ---- take the next row block of the oldest item that still has one -------------------------------
    int s = -1, rb = 0;
    if (lane == 0) {
      int spins = 0;
      for (;;) {
        const uint32_t st = lds_u32(&ctl->ticket);
        const int ts = (int)(st >> 8), trb = (int)(st & 255u);
        if (ts >= n_slots) break;  // nothing left for this CTA
        if ((int)lds_u32(&ctl->seq[ts % ATTP_NB]) != ts) {  // its slot is still being drained by the item before
          if (++spins > (1 << 24)) { atomicExch(err_flag, 3); s = -2; break; }  // bounded: never hang the GPU
          __nanosleep(64);
          continue;
        }
        const int nb = (int)((lds_u32(&ctl->slot[ts % ATTP_NB].dims) >> 16) & 255u);
        const uint32_t nxt = (trb + 1 < nb) ? st + 1u : (uint32_t)(ts + 1) << 8;
        if (atomicCAS(&ctl->ticket, st, nxt) == st) { s = ts; rb = trb; break; }
      }
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    rb = __shfl_sync(0xffffffffu, rb, 0);
    if (s < 0) break;
    const int b = s % ATTP_NB;
    const uint2 d = lds_u64(&ctl->slot[b]);
    const int r0 = (int)d.x, n_rows = (int)(d.y & 255u), n_keys = (int)((d.y >> 8) & 255u);
    const int nb = (int)((d.y >> 16) & 255u), head = (int)(d.y >> 24);
    const int l0 = rb * 16;
    uint32_t qa_hi[2][4], qa_lo[2][4];
    att_load_q<THREE>(qa_hi, qa_lo, qkv_hi, qkv_lo, r0, l0, n_rows, head, H);
    if (!mbar_wait(reinterpret_cast<uint64_t*>(&ctl->full[b]), (uint32_t)((s / ATTP_NB) & 1))) {
      if (lane == 0) atomicExch(err_flag, 4);
      break;
    }
    __threadfence_block();  // the issuer's bias row (generic stores) is ordered before its seq publication
    att_rows<THREE>(ring + (size_t)b * ATTP_SLOT_HALVES, Es_hi, Es_lo, Rw, bias_rows + b * 128, qa_hi, qa_lo, r0, l0,
                    n_rows, n_keys, head, H, ctx_hi, ctx_lo);
    // ---- the warp that finishes an item's last block refills the slot --------------------------------
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence_block();
      last = (atomicAdd(&ctl->done[b], 1) + 1 == nb);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last && s + ATTP_NB < n_slots)
      attp_issue<THREE>(ctl, ring, bias_rows, &map_hi, &map_lo, row_start, n_rows_arr, n_keys_arr, key_bias, n_pad, H,
                        heads, n_items, s + ATTP_NB);
  }
}

// [rows, ld] fp16 plane -> TMA map with box {32 halves (one head), 128 rows}, 64-byte swizzle (== att_sw).
inline int attp_make_map(CUtensorMap* map, const __half* base, int rows, int ld) {
  tc_encode_fn enc = tc_encoder();
  if (!enc) return 1;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)FD_HEAD_DIM, 128u};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <bool THREE>
inline int attp_launch(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const __half* qkv_hi, const __half* qkv_lo,
                       const int* row_start, const int* n_rows, const int* n_keys, const float* key_bias, int n_pad,
                       const __half* e_hi, const __half* e_lo, int H, int heads, int n_items, __half* ctx_hi,
                       __half* ctx_lo, int sm_count, cudaStream_t st) {
  static unsigned long long configured = 0;
  auto kern = attention_pool_kernel<THREE>;
  if (tc_need_configure(&configured) &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attp_smem_bytes()) != cudaSuccess)
    return 10;
  int* err = tc_err_flag();
  if (!err) return 11;
  const int grid = n_items < sm_count ? n_items : sm_count;
  kern<<<grid, ATT_WARPS * 32, attp_smem_bytes(), st>>>(map_hi, map_lo, qkv_hi, qkv_lo, row_start, n_rows, n_keys,
                                                        key_bias, n_pad, e_hi, e_lo, H, heads, n_items, ctx_hi, ctx_lo,
                                                        err);
  return cudaGetLastError() == cudaSuccess ? 0 : 12;
}

}  // namespace fd
