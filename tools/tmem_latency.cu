// Micro-benchmark of the TMEM / mbarrier round trips the tcgen05 attention kernel is built from (round-2 planning aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I foldingdiff_b200/csrc tools/tmem_latency.cu -o tools/tmem_latency
// Prints cycles (clock64, one SM) for: tcgen05.ld x16/x32/x64(2 x x32) + wait::ld, tcgen05.st x16 + wait::st, ld.shared,
// a batch of 12 small tcgen05.mma (M=128, N=32, K=16) + commit -> mbarrier, and the arrive -> try_wait hand-off between warps.
#include <cstdio>

#include "attention_tc.cuh"

using namespace fd;

__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out) {
  __shared__ __align__(1024) uint8_t tile[2 * 128 * 64];  // A / B operand tiles (zeros): 128 rows x 64 bytes each
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float buf[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (int)sizeof(tile) / 4; i += 128) reinterpret_cast<uint32_t*>(tile)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot, t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr int REP = 64;
  uint32_t v[32], w[32];
  long long t0, t1;
  uint32_t acc = 0;

  // 1. tcgen05.ld x16 + wait
  t0 = clock64();
  for (int r = 0; r < REP; ++r) { uint32_t u[16]; tmem_ld16(t_lane + (acc & 1), u); acc += u[0] & 1; }
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / REP;
  // 2. tcgen05.ld x32 + wait
  t0 = clock64();
  for (int r = 0; r < REP; ++r) { tmem_ld32(t_lane + (acc & 1), v); acc += v[0] & 1; }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = (t1 - t0) / REP;
  // 3. two x32 loads, one wait
  t0 = clock64();
  for (int r = 0; r < REP; ++r) { tmem_ld32_issue(t_lane + (acc & 1), v); tmem_ld32_issue(t_lane + 32, w); tmem_ld_wait(); acc += (v[0] ^ w[0]) & 1; }
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = (t1 - t0) / REP;
  // 4. four x32 loads (128 columns), one wait - issued back to back into the same registers (latency only)
  t0 = clock64();
  for (int r = 0; r < REP; ++r) {
    tmem_ld32_issue(t_lane + (acc & 1), v); tmem_ld32_issue(t_lane + 32, w); tmem_ld32_issue(t_lane + 64, v); tmem_ld32_issue(t_lane + 96, w);
    tmem_ld_wait(); acc += (v[0] ^ w[0]) & 1;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = (t1 - t0) / REP;
  // 5. tcgen05.st x16 + wait::st
  t0 = clock64();
  for (int r = 0; r < REP; ++r) { uint32_t u[16]; for (int i = 0; i < 16; ++i) u[i] = acc + i; tmem_st16(t_lane + 256, u); tmem_st_wait(); }
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = (t1 - t0) / REP;
  // 6. dependent ld.shared chain
  buf[threadIdx.x] = 0.0f;
  __syncthreads();
  t0 = clock64();
  float f = 0.0f;
  for (int r = 0; r < REP; ++r) f += lds_f32(smem_u32(&buf[(threadIdx.x + (int)f) & 127]));
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = (t1 - t0) / REP;
  acc += (uint32_t)f;
  __syncthreads();
  // 7. 12 small MMAs (N = 32) + commit -> own wait; 8. 12 larger MMAs (6 x N=96 + 6 x N=192) + commit
  if (warp == 1 && lane == 0) {
    const AtcDesc d = {(512u >> 4) | (1u << 14) | (4u << 29), (512u >> 4) | (1u << 14) | (4u << 29), 1, 1, umma_idesc_f16(32) | (1u << 16)};
    const uint32_t a_s = smem_u32(tile), b_s = a_s + 128 * 64;
    uint32_t ph = 0;
    long long sum_small = 0, sum_big = 0, sum_ts = 0;
    for (int r = 0; r < 16; ++r) {
      t0 = clock64();
      for (int k = 0; k < 12; ++k) umma_f16(tmem + 448, atc_desc(a_s, 1, d.k_hi32), atc_desc(b_s, 1, d.k_hi32), umma_idesc_f16(32), k);
      umma_commit(&bars[0]);
      mbar_wait(&bars[0], ph);
      t1 = clock64(); sum_small += t1 - t0;
      t0 = clock64();
      for (int k = 0; k < 12; ++k) umma_f16_ts(tmem + 448, tmem + 384 + 8 * (k & 3), atc_desc(b_s, d.v_lbo, d.v_hi32), d.pv_idesc, k);
      umma_commit(&bars[0]);
      mbar_wait(&bars[0], ph ^ 1);
      t1 = clock64(); sum_ts += t1 - t0;
      t0 = clock64();
      for (int k = 0; k < 6; ++k) umma_f16(tmem, atc_desc(a_s, 1, d.k_hi32), atc_desc(b_s, 1, d.k_hi32), umma_idesc_f16(96), k);
      for (int k = 0; k < 6; ++k) umma_f16(tmem + 128, atc_desc(a_s, 1, d.k_hi32), atc_desc(b_s, 1, d.k_hi32), umma_idesc_f16(128), k);
      umma_commit(&bars[0]);
      mbar_wait(&bars[0], ph);
      t1 = clock64(); sum_big += t1 - t0;
      ph ^= 1;
    }
    out[6] = sum_small / 16; out[7] = sum_ts / 16; out[8] = sum_big / 16;
  }
  __syncthreads();
  // 9. hand-off: warp 2 arrives on bars[1] and stamps the clock, warp 3 spins on try_wait and stamps when it sees it
  __shared__ long long stamp[2];
  long long hand = 0;
  for (int r = 0; r < 32; ++r) {
    __syncthreads();
    if (warp == 2 && lane == 0) { for (int i = 0; i < 200; ++i) acc += clock64() & 1; stamp[0] = clock64(); mbar_arrive(&bars[1]); }
    if (warp == 3 && lane == 0) { mbar_wait(&bars[1], (uint32_t)(r & 1)); stamp[1] = clock64(); }
    __syncthreads();
    hand += stamp[1] - stamp[0];
  }
  if (threadIdx.x == 0) out[9] = hand / 32;
  if (acc == 0x7fffffffu) out[15] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaMemset(d, 0, 16 * sizeof(long long));
  lat_kernel<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"tcgen05.ld x16 + wait::ld", "tcgen05.ld x32 + wait::ld", "2 x tcgen05.ld x32 + one wait", "4 x tcgen05.ld x32 + one wait",
                         "tcgen05.st x16 + wait::st", "ld.shared (dependent)", "12 x mma SS N=32 + commit -> wait",
                         "12 x mma TS N=32 (A in TMEM, B MN-major) + commit -> wait", "6 x mma N=96 + 6 x mma N=128 + commit -> wait",
                         "mbarrier arrive -> try_wait seen by another warp"};
  printf("status: %s\n", cudaGetErrorString(e));
  for (int i = 0; i < 10; ++i) printf("%-62s %6lld cycles\n", names[i], h[i]);
  return e == cudaSuccess ? 0 : 1;
}
