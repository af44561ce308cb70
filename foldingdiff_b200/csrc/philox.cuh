// Counter-based normal generator for callers that do not bring their own noise
// (fd_p_sample_steps_philox, fd_sample_host(noise_host == NULL), fd_randn: the throughput mode).  Philox4x32-10 + Box-Muller; element i of
// a stream is a pure function of (seed, offset + i), so results do not depend on chunking or on the
// number of GPUs a batch is sharded over.  This is NOT torch's stream: parity runs pass the
// reference's own draws (sampling.py:73) through `noise_dev` instead.
#pragma once
#include "common.cuh"

namespace fd {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float u01(uint32_t v) {  // (0, 1]
  return ((float)(v >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// element g of stream `seed`: counter = g / 4 (one Philox block = two Box-Muller pairs), lane = g % 4
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t g) {
  const uint64_t quad = g >> 2;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const int lane = (int)(g & 3);
  const uint32_t a = (lane < 2) ? r.x : r.z, b = (lane < 2) ? r.y : r.w;
  const float rad = sqrtf(-2.0f * logf(u01(a)));
  const float ang = 6.28318530717958647692f * u01(b);
  return (lane & 1) ? rad * sinf(ang) : rad * cosf(ang);
}

__global__ void philox_randn_kernel(float* __restrict__ dst, long long n, uint64_t seed,
                                    uint64_t offset) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // quad index in this call
  const long long i0 = q * 4;
  if (i0 >= n) return;
  // offsets need not be multiples of 4
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (i0 + j >= n) break;
    dst[i0 + j] = philox_normal(seed, offset + (uint64_t)(i0 + j));
  }
}

inline void launch_philox_randn(float* dst, long long n, uint64_t seed, uint64_t offset,
                                cudaStream_t st) {
  const long long quads = (n + 3) / 4;
  const int threads = 256;
  const long long blocks = (quads + threads - 1) / threads;
  philox_randn_kernel<<<(unsigned)blocks, threads, 0, st>>>(dst, n, seed, offset);
}

}  // namespace fd
