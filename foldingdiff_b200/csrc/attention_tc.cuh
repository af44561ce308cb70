// Relative-key attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   S[l, r] = (q_l . k_r + q_l . E[l - r + 127]) / sqrt(32) + bias_r ;  ctx_l = softmax_r(S) V
//   (transformers 4.11.3 BertSelfAttention, position_embedding_type = "relative_key";
//    call site /root/reference/foldingdiff/modelling.py:473)
//
// Work item = one (chain, head): the whole chain is ONE 128-row MMA tile (n <= 128 residues).  Per item
//
//   S = Q K^T          tcgen05.mma  M = 128, N = nk32, K = 32      (nk32 = keys rounded up to 32)
//   R = Q Ewin^T       tcgen05.mma  M = 128, N = nk32 + nr32       Ewin = rows [128 - nk32, 128 + nr32) of the
//                                                                   layer's distance table (resident in smem)
//   S[l, r] += R[l, l - r + nk32 - 1]      the Toeplitz gather ("skew"): thread l owns TMEM lane l, so it only
//                                          ever needs ITS OWN row of R at a lane-dependent column shift; the
//                                          shift goes through a thread-private shared-memory row (64-column
//                                          window per 32 keys: STS.128 in, 32 scalar LDS out, conflict-free)
//   P = exp2(S' - max) in registers (one thread = one query row), written back to TMEM as fp16 hi / lo
//   O = P V            tcgen05.mma  M = 128, N = 32, A = P from TMEM, B = V straight from the TMA tile
//                      (MN-major descriptor: V is [key][dim], no transpose pass)
//
// every product as the error-compensated triple hi*hi + hi*lo + lo*hi in one fp32 TMEM accumulator (the
// scheme of gemm_tc.cuh).  Q / K / V tiles (fp16 hi / lo planes written by the QKV GEMM epilogue) arrive by
// TMA, 64-byte swizzle, into a 3-slot ring.
//
// Warp roles (384 threads, one persistent CTA per SM):  warps 0..3 and 4..7 = two softmax warpgroups that take
// alternate items, warp 8 = TMA producer, warp 9 = S / R MMA issuer + TMEM owner, warps 10 / 11 = P V MMA issuers of
// warpgroup 0's / warpgroup 1's items.
//
// TMEM (512 columns), round-2 layout: every warpgroup owns a 128-column region, the relative-key product is shared -
//     warpgroup w:  [128 w, 128 w + 128)   S [0, nk32)  ->  P round 0 (keys 0..63) hi [0, 32) lo [32, 64) | O [64, 96) |
//                                          P round 1 hi / lo: [96, 112) / [112, 128) if nk32 <= 96, else [0, 32) / [32, 64)
//     R:            [256, 512)
// S is read into registers before anything else is written to the region, so P and O reuse its columns.  Round 1 only
// has to wait for round 0's P V when the chain has more than 96 keys, and by then its exponentials have been computed
// under that very product.  (Round-1 layout: ONE 64-column P buffer and one O for both warpgroups; ncu showed 39 % of the
// softmax warps' samples in the three waits of that hand-off - profiles/r01_attention_tc_analysis.md.)  The price: S of
// item i + 2 cannot be issued before O of item i has been read (R of the next item still can: it is issued first).
//
// Every mbarrier wait is bounded (mbar_wait): a broken pipeline sets the error flag and ends the kernel.
#pragma once
#include "attention_pool.cuh"

namespace fd {

constexpr int ATC_SLOTS = 3;
constexpr int ATC_PLANE_BYTES = 128 * 64;            // 128 rows x 32 halves, 64-byte rows
constexpr int ATC_SLOT_BYTES = 6 * ATC_PLANE_BYTES;  // Q hi, Q lo, K hi, K lo, V hi, V lo
constexpr int ATC_SCR_PITCH = 68;                    // words per thread-private skew row (== 4 mod 32)
constexpr int ATC_THREADS = 384;  // warpgroups: 0 and 1 = softmax, 2 = producers (warp 8 TMA, warp 9 MMA)
constexpr uint32_t ATC_COL_WG = 128, ATC_COL_R = 256, ATC_COL_O = 64, ATC_COL_P1 = 96;  // see the TMEM layout above
constexpr int ATC_DBG_ROW = 128 + 128 + 32 + 2;      // floats per row of the debug dump

constexpr size_t atc_smem_bytes() {
  return (size_t)ATC_SLOTS * ATC_SLOT_BYTES + 2 * ATT_E_TABLE * 64   // ring + E hi / lo
         + (size_t)128 * ATC_SCR_PITCH * 4                           // skew scratch (one warpgroup at a time)
         + 512 + 1024;                                               // barriers + item descriptor ring + alignment slack
}

// Operand descriptors, high 32 bits (SBO | version 1 | layout type) - host-computed so a debug run can override
// them from the environment; the low word is (address >> 4) | LBO << 16.
struct AtcDesc {
  uint32_t k_hi32;    // K-major, 64-byte swizzle: 8-row groups 512 bytes apart
  uint32_t v_hi32;    // MN-major, 64-byte swizzle: 8-key groups 512 bytes apart
  uint32_t k_lbo, v_lbo;
  uint32_t pv_idesc;  // kind::f16 instruction descriptor of P V (N = 32, B MN-major)
};
inline AtcDesc atc_default_desc() {
  AtcDesc d;
  d.k_hi32 = (512u >> 4) | (1u << 14) | (4u << 29);
  d.v_hi32 = (512u >> 4) | (1u << 14) | (4u << 29);
  d.k_lbo = 1; d.v_lbo = 1;
  d.pv_idesc = umma_idesc_f16(32) | (1u << 16);
  auto env = [](const char* name, uint32_t& v) {
    const char* e = getenv(name);
    if (e && e[0]) v = (uint32_t)strtoul(e, nullptr, 0);
  };
  env("FOLDINGDIFF_B200_ATC_KHI", d.k_hi32); env("FOLDINGDIFF_B200_ATC_VHI", d.v_hi32);
  env("FOLDINGDIFF_B200_ATC_KLBO", d.k_lbo); env("FOLDINGDIFF_B200_ATC_VLBO", d.v_lbo);
  env("FOLDINGDIFF_B200_ATC_PVIDESC", d.pv_idesc);
  return d;
}

__device__ __forceinline__ uint64_t atc_desc(uint32_t smem_addr, uint32_t lbo, uint32_t hi32) {
  return (uint64_t)(((smem_addr & 0x3FFFFu) >> 4) | (lbo << 16)) | ((uint64_t)hi32 << 32);
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// explicit shared-space accesses for the skew scratch: through the generic pointer (the 1024-byte alignment
// arithmetic hides the address space from the compiler) they compile to LD.E / ST.E
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// split2 (attention_mma.cuh) of (x0, x1) * (1 + c), c ~ 1e-6, with the scale folded into the lo words:
// hi = fp16(x), lo = fp16(x - hi * t), t = 1 - c  (= (x - hi) + hi * c; below 1 the fp32 grid is twice as fine as above)
__device__ __forceinline__ void split2s(float x0, float x1, float t, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(fmaf(-hf.x, t, x0), fmaf(-hf.y, t, x1));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier helpers on shared-space addresses (computed once per thread).  The wait passes a suspend-time hint:
// ncu on the first version showed 37% of all issued instructions were try_wait spin iterations of parked warps.
__device__ __forceinline__ bool atc_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(100000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool atc_wait(uint32_t bar, uint32_t parity) {  // bounded like mbar_wait (~2 s)
  if (atc_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!atc_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ void atc_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void atc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t lds_u32a(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
constexpr int ATC_META_RING = 8;  // item descriptors staged in shared memory by the TMA warp, 32 bytes each

struct AtcItem {  // raw loads only: the derived sizes are computed at the point of use, so a descriptor fetched one
  int chain, head, r0, n_rows, n_keys;  // iteration ahead never stalls the iteration that issued its loads
  __device__ __forceinline__ int nk32() const { return (n_keys + 31) & ~31; }
  __device__ __forceinline__ int nr32() const { return (n_rows + 31) & ~31; }
};
__device__ __forceinline__ AtcItem atc_item(int it, int n_items, int heads, const int* __restrict__ row_start,
                                            const int* __restrict__ n_rows_arr, const int* __restrict__ n_keys_arr) {
  AtcItem a;
  const int item = n_items - 1 - ((int)blockIdx.x + it * (int)gridDim.x);  // batches are length-sorted: long first
  a.chain = item / heads; a.head = item % heads;
  // asm volatile: issued HERE.  As plain __restrict__ loads the compiler sank them to their first use in the next
  // iteration (ncu: 17% of the softmax warps' time was the L2 latency of these three words).
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(a.r0) : "l"(row_start + a.chain));
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(a.n_rows) : "l"(n_rows_arr + a.chain));
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(a.n_keys) : "l"(n_keys_arr + a.chain));
  return a;
}

template <bool DBG>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const int* __restrict__ row_start, const int* __restrict__ n_rows_arr,
                    const int* __restrict__ n_keys_arr, const float* __restrict__ key_bias, int n_pad,
                    const __half* __restrict__ e_hi, const __half* __restrict__ e_lo, int H, int heads, int n_items,
                    __half* __restrict__ ctx_hi, __half* __restrict__ ctx_lo, AtcDesc dsc, float* __restrict__ dbg,
                    int* __restrict__ err_flag, float rz_alpha, float rz_beta) {
  pdl_trigger();
  extern __shared__ uint8_t atc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  __half* Es_hi = reinterpret_cast<__half*>(smem + ATC_SLOTS * ATC_SLOT_BYTES);
  __half* Es_lo = Es_hi + ATT_E_TABLE * ATT_PITCH;
  float* scr = reinterpret_cast<float*>(Es_lo + ATT_E_TABLE * ATT_PITCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scr + 128 * ATC_SCR_PITCH);
  uint64_t* kv_full = bars;                  // [3] TMA -> MMA
  uint64_t* kv_empty = bars + ATC_SLOTS;     // [3] MMA (commit) -> TMA
  // A barrier a warpgroup WAITS on must show it every phase: a waiter that only looks at every other phase cannot
  // tell "two phases behind" from "done" by parity (warpgroup 1 would sail through its very first wait).  So the
  // MMA -> softmax barriers exist once per warpgroup; the softmax -> MMA ones are seen in item order by the MMA warp.
  uint64_t* sr_full = bars + 2 * ATC_SLOTS;  // [2] MMA -> softmax warpgroup (it & 1): its S and the shared R are complete
  uint64_t* o_full = sr_full + 2;            // [2] MMA -> softmax warpgroup: O complete
  uint64_t* p0_done = sr_full + 4;           // [2] MMA -> softmax warpgroup: round-0 P V retired (its P columns may be rewritten)
  uint64_t* reg_empty = sr_full + 6;         // [2] softmax warpgroup -> MMA: O has been read, the region is free for the next S
  uint64_t* p_full = sr_full + 8;            // [2][2] softmax warpgroup -> MMA, per 64-key round: index 2 w + r
  uint64_t* r_empty = sr_full + 12;          // softmax -> MMA: the skew has consumed R
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sr_full + 13);
  // Item descriptors {r0, n_rows, n_keys, head, chain}: the TMA warp, three items ahead of everyone, is the only role
  // that reads them from global memory; it parks them in an 8-entry shared ring BEFORE arming the item's kv_full
  // barrier, so every later role reads them with one LDS after a barrier it waits on anyway.  (ncu: with each role
  // prefetching its own copy the compiler spilled the in-flight registers and stalled at the spill - ISETP / IMAD /
  // BRA / STL long-scoreboard stalls were a quarter of the softmax warps' busy samples.)  An entry is rewritten 8
  // items later, after kv_empty of item it + 5: both warpgroups have left item `it` by then.
  int* meta = reinterpret_cast<int*>(bars) + 48;  // byte 192 of the 512-byte barrier block (barriers + TMEM slot end at 156)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_it = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA

  if (tid == 0) {
    for (int i = 0; i < ATC_SLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }
    mbar_init(r_empty, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sr_full[i], 1); mbar_init(&o_full[i], 1); mbar_init(&p0_done[i], 1); mbar_init(&reg_empty[i], 4);
      mbar_init(&p_full[2 * i], 4); mbar_init(&p_full[2 * i + 1], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&map_hi); tma_prefetch_desc(&map_lo);
    meta[8 * ATC_META_RING] = n_it;  // loop bound of the softmax warps, read back with LDS (one register less to spill)
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  // the layer's distance table: 256 rows x 64 bytes per plane, 64-byte swizzle (same layout the TMA tiles have)
  for (int i = tid; i < ATT_E_TABLE * 4; i += ATC_THREADS) {
    const int r = i >> 2;
    cp_async16(Es_hi + att_sw(r, i & 3), e_hi + (size_t)r * FD_HEAD_DIM + (i & 3) * 8);
    cp_async16(Es_lo + att_sw(r, i & 3), e_lo + (size_t)r * FD_HEAD_DIM + (i & 3) * 8);
  }
  cp_async_wait_all();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes above -> tensor-core (async proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // the prologue (barriers, TMEM, the layer's distance table: weights) ran under the QKV GEMM's tail
  // shared-space addresses of the barriers (8 bytes each, same order as above)
  const uint32_t b_kv_full = smem_u32(kv_full), b_kv_empty = smem_u32(kv_empty), b_sr_full = smem_u32(sr_full);
  const uint32_t b_o_full = smem_u32(o_full), b_p0_done = smem_u32(p0_done), b_reg_empty = smem_u32(reg_empty);
  const uint32_t b_p_full = smem_u32(p_full), b_r_empty = smem_u32(r_empty);
  const uint32_t meta_s = smem_u32(meta);
  auto meta_item = [&](int it) {  // valid once a barrier downstream of kv_full(it) has been waited for
    const uint32_t m = meta_s + 32u * (uint32_t)(it & (ATC_META_RING - 1));
    AtcItem a;
    a.r0 = (int)lds_u32a(m); a.n_rows = (int)lds_u32a(m + 4); a.n_keys = (int)lds_u32a(m + 8);
    a.head = (int)lds_u32a(m + 12); a.chain = (int)lds_u32a(m + 16);
    return a;
  };

  // register reallocation between the warpgroups (pool = 384 x 168): producers keep 40, softmax threads get 232 -
  // one query row is 128 fp32 logits plus the fp16 hi / lo staging, which does not fit the static 168
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      AtcItem nxt = atc_item(0, n_items, heads, row_start, n_rows_arr, n_keys_arr);
      for (int it = 0; it < n_it; ++it) {
        const int slot = it % ATC_SLOTS, use = it / ATC_SLOTS;
        const AtcItem a = nxt;
        if (it + 1 < n_it) nxt = atc_item(it + 1, n_items, heads, row_start, n_rows_arr, n_keys_arr);
        if (!atc_wait(b_kv_empty + 8 * slot, (uint32_t)((use & 1) ^ 1))) { atomicExch(err_flag, 301); break; }
        {
          int* m = meta + 8 * (it & (ATC_META_RING - 1));
          m[0] = a.r0; m[1] = a.n_rows; m[2] = a.n_keys; m[3] = a.head; m[4] = a.chain;
          __threadfence_block();  // ordered before the (releasing) expect_tx arrive below
        }
        uint8_t* s = ring + (size_t)slot * ATC_SLOT_BYTES;
        const int cq = a.head * FD_HEAD_DIM;
        mbar_expect_tx(&kv_full[slot], ATC_SLOT_BYTES);
        // Q / K / V planes are read exactly once (by this kernel): evict-first keeps the 206 MB stream from pushing the
        // ctx planes written below - the next GEMM's A operand - out of the 126 MB L2
#ifdef ATC_NO_EVICT_HINT
#define ATC_TMA(dst, map, col) tma_load_2d(dst, map, &kv_full[slot], col, a.r0)
#else
#define ATC_TMA(dst, map, col) tma_load_2d_hint(dst, map, &kv_full[slot], col, a.r0, TC_EVICT_FIRST)
#endif
        ATC_TMA(s, &map_hi, cq);
        ATC_TMA(s + 2 * ATC_PLANE_BYTES, &map_hi, H + cq);
        ATC_TMA(s + 1 * ATC_PLANE_BYTES, &map_lo, cq);
        ATC_TMA(s + 3 * ATC_PLANE_BYTES, &map_lo, H + cq);
        ATC_TMA(s + 4 * ATC_PLANE_BYTES, &map_hi, 2 * H + cq);
        ATC_TMA(s + 5 * ATC_PLANE_BYTES, &map_lo, 2 * H + cq);
#undef ATC_TMA
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t e_hi_s = smem_u32(Es_hi), e_lo_s = smem_u32(Es_lo);
      // item descriptors one iteration ahead: their (dependent) global loads stay off the issue path
      for (int it = 0; it < n_it; ++it) {
        {  // ---- R (shared region) and S (the warpgroup's region) of item `it`
          const int slot = it % ATC_SLOTS, w = it & 1, j = it >> 1;
          if (!atc_wait(b_kv_full + 8 * slot, (uint32_t)((it / ATC_SLOTS) & 1))) { atomicExch(err_flag, 302); break; }
          const AtcItem a = meta_item(it);
          const int nk32 = a.nk32(), nr32 = a.nr32();
          const uint32_t s0 = smem_u32(ring + (size_t)slot * ATC_SLOT_BYTES);
          const uint32_t q_hi = s0, q_lo = s0 + ATC_PLANE_BYTES, k_hi = s0 + 2 * ATC_PLANE_BYTES, k_lo = s0 + 3 * ATC_PLANE_BYTES;
          const uint32_t id_s = umma_idesc_f16(nk32), id_r = umma_idesc_f16(nk32 + nr32);
          const uint32_t e_off = (uint32_t)(128 - nk32) * 64u;
          // R first: its region frees early (after the previous item's skew), the S region only after that warpgroup's
          // previous item has read its O
          if (!atc_wait(b_r_empty, (uint32_t)((it & 1) ^ 1))) { atomicExch(err_flag, 303); break; }
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ko = ks * 32;  // bytes inside the 64-byte swizzle row
            const uint64_t dq_hi = atc_desc(q_hi + ko, dsc.k_lbo, dsc.k_hi32), dq_lo = atc_desc(q_lo + ko, dsc.k_lbo, dsc.k_hi32);
            umma_f16(tmem + ATC_COL_R, dq_hi, atc_desc(e_hi_s + e_off + ko, dsc.k_lbo, dsc.k_hi32), id_r, ks);
            umma_f16(tmem + ATC_COL_R, dq_hi, atc_desc(e_lo_s + e_off + ko, dsc.k_lbo, dsc.k_hi32), id_r, 1u);
            umma_f16(tmem + ATC_COL_R, dq_lo, atc_desc(e_hi_s + e_off + ko, dsc.k_lbo, dsc.k_hi32), id_r, 1u);
          }
          if (!atc_wait(b_reg_empty + 8 * w, (uint32_t)((j & 1) ^ 1))) { atomicExch(err_flag, 309); break; }
          tc_fence_after();
          const uint32_t t_s = tmem + ATC_COL_WG * (uint32_t)w;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ko = ks * 32;
            const uint64_t dq_hi = atc_desc(q_hi + ko, dsc.k_lbo, dsc.k_hi32), dq_lo = atc_desc(q_lo + ko, dsc.k_lbo, dsc.k_hi32);
            umma_f16(t_s, dq_hi, atc_desc(k_hi + ko, dsc.k_lbo, dsc.k_hi32), id_s, ks);
            umma_f16(t_s, dq_hi, atc_desc(k_lo + ko, dsc.k_lbo, dsc.k_hi32), id_s, 1u);
            umma_f16(t_s, dq_lo, atc_desc(k_hi + ko, dsc.k_lbo, dsc.k_hi32), id_s, 1u);
          }
          atc_commit(b_sr_full + 8 * w);
          atc_commit(b_kv_empty + 8 * slot);  // this thread's half of the slot release (Q / K reads retired)
        }
      }
    }
  } else {
    // ===================== MMA issuers 2 and 3 (warps 10, 11): O = P V of warpgroup 0's / warpgroup 1's items ==========
    // Threads of their own: with one issuer walking S/R(it), PV(it-1), S/R(it+1), ... in program order the S / R
    // products of the next item queued behind the P hand-off of the previous one - on the softmax warpgroups'
    // critical path (ncu, round 1: they spent most of their wait time on sr_full).  One P V issuer PER warpgroup since
    // round 2: a single one walks the items in order and a warpgroup whose P is ready waits behind the other
    // warpgroup's unfinished round (ncu: 13 % of all samples in the o_full wait).
    if (lane == 0) {
      bool ok = true;
      for (int j = warp - 10; j < n_it && ok; j += 2) {
        {  // ---- O = P V of item j, 64 keys per round, inside the item's warpgroup region
          const int slot = j % ATC_SLOTS, w = j & 1;
          const uint32_t wpar = (uint32_t)((j >> 1) & 1);
          if (!atc_wait(b_kv_full + 8 * slot, (uint32_t)((j / ATC_SLOTS) & 1))) { atomicExch(err_flag, 310); break; }
          const int nk32 = meta_item(j).nk32();
          const uint32_t s0 = smem_u32(ring + (size_t)slot * ATC_SLOT_BYTES);
          const uint32_t v_hi = s0 + 4 * ATC_PLANE_BYTES, v_lo = s0 + 5 * ATC_PLANE_BYTES;
          const uint32_t t_w = tmem + ATC_COL_WG * (uint32_t)w, t_o = t_w + ATC_COL_O;
          for (int r = 0; r < 2 && ok; ++r) {
            if (!atc_wait(b_p_full + 8 * (2 * w + r), wpar)) { atomicExch(err_flag, 305); ok = false; break; }
            tc_fence_after();
            const int nks = min(4, (nk32 - 64 * r) >> 4);
            // round 1 of a chain of at most 96 keys has its own 32 columns; otherwise the rounds share [0, 64)
            const bool own = r == 1 && nk32 <= 96;
            const uint32_t p_hi0 = t_w + (own ? ATC_COL_P1 : 0u), lo_off = own ? 16u : 32u;
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t vb = (uint32_t)(64 * r + 16 * ks) * 64u;
              const uint32_t p_hi = p_hi0 + 8 * ks, p_lo = p_hi + lo_off;
              umma_f16_ts(t_o, p_hi, atc_desc(v_hi + vb, dsc.v_lbo, dsc.v_hi32), dsc.pv_idesc, (r | ks) != 0 ? 1u : 0u);
              umma_f16_ts(t_o, p_lo, atc_desc(v_hi + vb, dsc.v_lbo, dsc.v_hi32), dsc.pv_idesc, 1u);
              umma_f16_ts(t_o, p_hi, atc_desc(v_lo + vb, dsc.v_lbo, dsc.v_hi32), dsc.pv_idesc, 1u);
            }
            if (r == 0) atc_commit(b_p0_done + 8 * w);
          }
          atc_commit(b_o_full + 8 * w);
          atc_commit(b_kv_empty + 8 * slot);  // the other half of the slot release (V reads retired)
        }
      }
    }
  }
  } else {
    // ===================== softmax warpgroups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wg = warp >> 2, quad = warp & 3;  // TMEM lanes [32 quad, 32 quad + 32) are this warp's
    const int row = quad * 32 + lane;                 // query row == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_reg = t_lane + ATC_COL_WG * (uint32_t)wg;  // this warpgroup's S / P / O region
    float* srow = scr + (size_t)row * ATC_SCR_PITCH;
    const float c_scale = 0.17677669529663688110f * 1.44269504088896340736f;  // log2(e) / sqrt(32)
    // descriptor fields are re-read from the shared ring wherever they are used (an LDS each) instead of being kept
    // in registers across the item: held live they were spilled, and the reloads missed L1 under the store traffic
    auto fld = [&](int it, int k) { return (int)lds_u32a(meta_s + 32u * (uint32_t)(it & (ATC_META_RING - 1)) + 4u * (uint32_t)k); };
#define ATC_ACTIVE(it) (quad * 32 < fld(it, 1))
    for (int it = wg; it < (int)lds_u32a(meta_s + 32u * ATC_META_RING); it += 2) {  // n_it, parked behind the ring
      const uint32_t wpar = (uint32_t)((it >> 1) & 1);  // per-warpgroup phase parity
      uint32_t su[128];
      if (!atc_wait(b_sr_full + 8 * wg, wpar)) { if (lane == 0) atomicExch(err_flag, 306); break; }
      tc_fence_after();
      const int nk32 = (fld(it, 2) + 31) & ~31;
      float* drow = (DBG && dbg) ? dbg + ((size_t)(fld(it, 4) * heads + fld(it, 3)) * 128 + row) * ATC_DBG_ROW : nullptr;
      // All per-key work below is organised in 32-key chunks guarded by ONE warp-uniform test each; inside a chunk
      // the code is straight-line (ncu on the first version: a branch per key cost a third of the kernel).
      float m = -INFINITY, sum = 0.0f;
      if (ATC_ACTIVE(it)) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c * 32 < nk32) tmem_ld32_issue(t_reg + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&su[32 * c]));
        tmem_ld_wait();
      }
      if (ATC_ACTIVE(it)) {
        if (DBG && drow) {
#pragma unroll
          for (int k = 0; k < 128; ++k) if (k < nk32) drow[k] = __uint_as_float(su[k]);
        }
        // ---- skew: S[l, 32 c + i] += R[l, 32 (quad - c) + nk32 - 32 + (lane - i + 31)]
        const uint32_t srow_s = smem_u32(srow), win_s = srow_s + 4u * (uint32_t)(lane + 31);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < nk32) {
            const uint32_t cb = (uint32_t)(32 * (quad - c) + nk32 - 32);
            // two 32-column pieces through ONE register array: with both in flight (64 registers next to the 128
            // logits) loop invariants spilled, and their reloads miss L1 (ncu: 14 % local hit rate under the store
            // traffic) - a tcgen05.ld + wait is 58 cycles (profiles/r01_tmem_latency.txt), an L2 round trip ten times that
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld32(t_lane + ATC_COL_R + cb + 32 * h, v);
#pragma unroll
              for (int q = 0; q < 8; ++q) sts_v4(srow_s + 128 * h + 16 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
              su[32 * c + i] = __float_as_uint(__uint_as_float(su[32 * c + i]) + lds_f32(win_s - 4u * (uint32_t)i));
          }
        }
        if (DBG && drow) {
#pragma unroll
          for (int k = 0; k < 128; ++k) if (k < nk32) drow[128 + k] = __uint_as_float(su[k]);
        }
      }
      tc_fence_before();  // every TMEM read of S / R by this warp has completed (tcgen05.wait::ld above)
      __syncwarp();
      if (lane == 0) atc_arrive(b_r_empty);  // R may be overwritten by the next item's Q E^T (S stays this warpgroup's)

      // ---- mask, row max on the raw logits, p = 2^((s - max) * log2e / sqrt(32)), row sum
      if (ATC_ACTIVE(it)) {
        const int n_keys = fld(it, 2);
        if (key_bias) {  // additive attention mask of the forward API, folded into the raw logits (x sqrt(32))
          const float* kb = key_bias + (size_t)fld(it, 4) * n_pad;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 32 < nk32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int k = 32 * c + i;
                su[k] = __float_as_uint(fmaf(kb[min(k, n_keys - 1)], 5.65685424949238019521f, __uint_as_float(su[k])));
              }
            }
          }
        }
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < nk32) {
            if (c * 32 + 32 > n_keys) {  // the chunk that holds the end of the chain: keys >= n_keys -> -inf
#pragma unroll
              for (int i = 0; i < 32; ++i)
                su[32 * c + i] = (32 * c + i < n_keys) ? su[32 * c + i] : 0xff800000u;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(su[32 * c + i])); m1 = fmaxf(m1, __uint_as_float(su[32 * c + i + 1]));
              m2 = fmaxf(m2, __uint_as_float(su[32 * c + i + 2])); m3 = fmaxf(m3, __uint_as_float(su[32 * c + i + 3]));
            }
          }
        }
        m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      // ---- per 64-key round: p = 2^((s - max) * log2e / sqrt(32)), row sum, P as fp16 hi / lo planes into TMEM.
      // The exponentials of round 1 are computed AFTER round 0 has been handed to the tensor core, so they run under
      // its P V products (ncu, round 1: the wait for p_empty[0] behind a finished softmax was 11 % of all samples).
      const float neg = -m * c_scale;
      float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
      bool ok = true;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool live = ATC_ACTIVE(it) && 64 * r < nk32;
        if (live) {
#pragma unroll
          for (int c = 2 * r; c < 2 * r + 2; ++c) {
            if (c * 32 < nk32) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const int k = 32 * c + i;
                const float p0 = ex2_approx(fmaf(__uint_as_float(su[k]), c_scale, neg));
                const float p1 = ex2_approx(fmaf(__uint_as_float(su[k + 1]), c_scale, neg));
                const float p2 = ex2_approx(fmaf(__uint_as_float(su[k + 2]), c_scale, neg));
                const float p3 = ex2_approx(fmaf(__uint_as_float(su[k + 3]), c_scale, neg));
                su[k] = __float_as_uint(p0); su[k + 1] = __float_as_uint(p1);
                su[k + 2] = __float_as_uint(p2); su[k + 3] = __float_as_uint(p3);
                s0 += p0; s1 += p1; s2 += p2; s3 += p3;
              }
            }
          }
        }
        // round 0 overwrites S (already in registers); round 1 has its own columns unless the chain has more than 96
        // keys - then it reuses round 0's once that P V has retired (its exponentials were computed under it)
        const bool own = r == 1 && nk32 <= 96;
        if (r == 1 && nk32 > 96) {
          if (!atc_wait(b_p0_done + 8 * wg, wpar)) { if (lane == 0) atomicExch(err_flag, 307); ok = false; break; }
          tc_fence_after();
        }
        if (live) {
          const uint32_t t_p = t_reg + (own ? ATC_COL_P1 : 0u), lo_off = own ? 16u : 32u;
#pragma unroll
          for (int g = 0; g < 2; ++g) {  // 32 keys -> 16 packed columns per plane
            if (64 * r + 32 * g < nk32) {
              uint32_t ph[16], pl[16];
              // accumulation de-bias of O = P V (gemm_tc.cuh: tc_rz): the 16-key chunk j of nk32 / 16 still has
              // 3 (nk32 / 16 - j) truncating accumulates ahead of it; its inverse rides in the lo plane for free
              // (lo = p - hi * (1 - c) is the FSUB of the plain split turned into an FMA)
              const float sc0 = 1.0f - (rz_alpha + rz_beta * (float)((nk32 >> 4) - (4 * r + 2 * g)));
              const float sc1 = 1.0f - (rz_alpha + rz_beta * (float)((nk32 >> 4) - (4 * r + 2 * g + 1)));
#pragma unroll
              for (int q = 0; q < 16; ++q)
                split2s(__uint_as_float(su[64 * r + 32 * g + 2 * q]), __uint_as_float(su[64 * r + 32 * g + 2 * q + 1]),
                        q < 8 ? sc0 : sc1, ph[q], pl[q]);
              tmem_st16(t_p + 16 * g, ph);
              tmem_st16(t_p + lo_off + 16 * g, pl);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) atc_arrive(b_p_full + 8 * (2 * wg + r));
      }
      sum = (s0 + s1) + (s2 + s3);
      m *= c_scale;  // (debug dump: row max in log2 units)
      if (!ok) break;
      // ---- O: normalise and store ctx as hi / lo planes
      if (!atc_wait(b_o_full + 8 * wg, wpar)) { if (lane == 0) atomicExch(err_flag, 308); break; }
      tc_fence_after();
      uint32_t o[32];
      const bool active = ATC_ACTIVE(it);
      if (active) tmem_ld32(t_reg + ATC_COL_O, o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) atc_arrive(b_reg_empty + 8 * wg);  // the region is free: the S product of this warpgroup's next item may land
      if (active && row < fld(it, 1)) {
        const float inv = 1.0f / sum;
        if (DBG && drow) {
#pragma unroll
          for (int d = 0; d < 32; ++d) drow[256 + d] = __uint_as_float(o[d]);
          drow[288] = m; drow[289] = sum;
        }
        uint32_t oh[16], ol[16];
#pragma unroll
        for (int q = 0; q < 16; ++q)
          split2(__uint_as_float(o[2 * q]) * inv, __uint_as_float(o[2 * q + 1]) * inv, oh[q], ol[q]);
        const size_t off = (size_t)(fld(it, 0) + row) * H + fld(it, 3) * FD_HEAD_DIM;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(ctx_hi + off + 8 * q) = make_uint4(oh[4 * q], oh[4 * q + 1], oh[4 * q + 2], oh[4 * q + 3]);
          *reinterpret_cast<uint4*>(ctx_lo + off + 8 * q) = make_uint4(ol[4 * q], ol[4 * q + 1], ol[4 * q + 2], ol[4 * q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

#undef ATC_ACTIVE

// FOLDINGDIFF_B200_ATT: unset / "tc" = this kernel, "pool" = the mma.sync kernel of attention_pool.cuh.  (A second
// tcgen05 version with two threads per query row was measured slower in round 1 - profiles/r01_attention_tc_analysis.md -
// and has been removed.)
inline bool atc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FOLDINGDIFF_B200_ATT");
    v = (!e || !e[0] || e[0] == 't') ? 1 : 0;
  }
  return v != 0;
}
inline float*& atc_debug_dump() {  // test hook: device buffer [items][128][ATC_DBG_ROW] or nullptr
  static float* p = nullptr;
  return p;
}

inline int atc_launch(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const int* row_start, const int* n_rows,
                      const int* n_keys, const float* key_bias, int n_pad, const __half* e_hi, const __half* e_lo, int H,
                      int heads, int n_items, __half* ctx_hi, __half* ctx_lo, int sm_count, cudaStream_t st) {
  static unsigned long long configured = 0;
  static const AtcDesc dsc = atc_default_desc();
  if (tc_need_configure(&configured) &&
      (cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atc_smem_bytes()) != cudaSuccess ||
       cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atc_smem_bytes()) != cudaSuccess))
    return 10;
  int* err = tc_err_flag();
  if (!err) return 11;
  const int grid = n_items < sm_count ? n_items : sm_count;
  const TcRz rz = tc_rz();
  cudaError_t e;
  if (atc_debug_dump())
    e = launch_pdl(attention_tc_kernel<true>, dim3(grid), dim3(ATC_THREADS), atc_smem_bytes(), st, map_hi, map_lo, row_start,
                   n_rows, n_keys, key_bias, n_pad, e_hi, e_lo, H, heads, n_items, ctx_hi, ctx_lo, dsc, atc_debug_dump(), err,
                   (float)rz.alpha, (float)rz.beta_att);
  else
    e = launch_pdl(attention_tc_kernel<false>, dim3(grid), dim3(ATC_THREADS), atc_smem_bytes(), st, map_hi, map_lo, row_start,
                   n_rows, n_keys, key_bias, n_pad, e_hi, e_lo, H, heads, n_items, ctx_hi, ctx_lo, dsc, (float*)nullptr, err,
                   (float)rz.alpha, (float)rz.beta_att);
  return e == cudaSuccess ? 0 : 12;
}

}  // namespace fd
