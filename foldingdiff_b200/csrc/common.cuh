// Shared helpers for the foldingdiff_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FD_HEAD_DIM 32
#define FD_MAX_FEATURES 16
#define FD_ROW_TILE 256  // packed-row count is padded to this (two MMA M tiles: one per CTA of a 2-cluster)

namespace fd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// HF "gelu" == torch.nn.functional.gelu (exact erf form), modelling.py:196 / BertIntermediate.
__device__ __forceinline__ float gelu_erf(float x) {
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}

// utils.modulo_with_wrapped_range(v, -pi, pi) on an fp32 tensor (utils.py:99-106):
// ((v - lo) % (hi - lo)) + lo with the Python-float bounds rounded to fp32 and torch's
// floor-mod (fmod, then shift negatives up by the divisor).  Bit-exact with torch CPU.
__device__ __forceinline__ float wrap_pi(float v) {
  const float lo = -3.14159265358979323846f;   // float(-pi)
  const float span = 6.28318530717958647692f;  // float(2*pi)
  float s = __fsub_rn(v, lo);
  float m = fmodf(s, span);
  if (m != 0.0f && m < 0.0f) m = __fadd_rn(m, span);
  return __fadd_rn(m, lo);
}

}  // namespace fd
