// Shared helpers for the foldingdiff_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#define FD_HEAD_DIM 32
#define FD_MAX_FEATURES 16
#define FD_ROW_TILE 256  // packed-row count is padded to this (two MMA M tiles: one per CTA of a 2-cluster)

namespace fd {

// Programmatic dependent launch (PDL).  Every kernel of the reverse step is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (launch_pdl below): kernel N + 1 may be scheduled onto SMs that
// kernel N has already left and run its prologue (barrier init, TMEM allocation, tensor-map prefetch, the attention
// kernel's distance-table load) under N's tail.  Contract kept by EVERY kernel on the step path:
//   pdl_trigger()  first statement: the dependent grid may launch as soon as all CTAs of this one have started;
//   pdl_wait()     before the first access (read OR write) to memory another kernel of the chain touches: returns
//                  when the preceding grid has completed and its writes are visible.  Each kernel waits for its
//                  predecessor, so completion is transitive along the chain (WAR hazards included).
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// Same function, 15 instructions instead of ~27 (the exact-erf GELU made the FFN1 GEMM epilogue issue-bound:
// ncu tensor pipe 49% vs 77-79% for its siblings).  With h = x/2 and e = erfc(|x|/sqrt2):
//     gelu(x) = (h + |h|) - |h| * e,         e = exp2(a * q(a)),  a = min(|x|, 4 sqrt2)
// q is a degree-8 minimax fit of -log2(erfc(a/sqrt2))/a (tools/fit_gelu.py).  Measured against fp64 on
// 1.8M points in [-9, 9]: max abs error 2.4e-7 (torch's own fp32 F.gelu: 1.2e-6), i.e. within 1 ulp of the
// result everywhere - and unlike 0.5x(1+erf) it keeps full relative precision on the negative tail.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float a = fminf(fabsf(x), 5.656854153f);
  float q = 5.128660518e-07f;
  q = fmaf(q, a, -9.560286344e-06f);
  q = fmaf(q, a, 7.497410843e-05f);
  q = fmaf(q, a, -2.843483817e-04f);
  q = fmaf(q, a, 1.499190421e-05f);
  q = fmaf(q, a, 6.931118667e-03f);
  q = fmaf(q, a, -5.243476480e-02f);
  q = fmaf(q, a, -4.592214525e-01f);
  q = fmaf(q, a, -1.151104212e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * a));
  const float h = 0.5f * x, ah = fabsf(h);
  return fmaf(-ah, e, h + ah);
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

// Host side: launch `kern` with the PDL attribute (FOLDINGDIFF_B200_PDL=0 turns the overlap off for A/B runs).
#ifdef __CUDACC__
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("FOLDINGDIFF_B200_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

}  // namespace fd
