/*
 * Plain-C caller of the sampler's C ABI (include/foldingdiff_b200.h): the whole reverse-diffusion loop with host
 * buffers in and out - what /root/reference/foldingdiff/sampling.py:79-132 (p_sample_loop) does for a Python caller.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -I include examples/sample_host.c \
 *       -L foldingdiff_b200/csrc -lfoldingdiff_b200 -Wl,-rpath,$PWD/foldingdiff_b200/csrc -lm -o sample_host
 *   ./sample_host [chains [timesteps]]
 *
 * Weights are synthetic (a seeded LCG; LayerNorm gains 1) because a checkpoint reader is not part of the C boundary:
 * a real caller fills `weights[]` in the order documented above FD_W_HEAD from its own state-dict reader.  The
 * schedule tables are the caller's job as well, exactly as for the Python host (fd_create comment): the linear beta
 * schedule of beta_schedules.py:20-27 and the posterior coefficients of beta_schedules.py:45-62 are evaluated here.
 * On a machine without an sm_100 device fd_create fails with FD_ERR_CUDA and the program says so: there is no CPU
 * fallback to fall into.  tests/test_c_example.py builds this file as strict C99 on the CPU box and checks exactly that;
 * past fd_create it has not been executed on a GPU yet (the entry points it calls are covered by tests/test_gpu_*.py
 * through ctypes).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "foldingdiff_b200.h"

static uint32_t lcg_state = 7344u;

static float lcg_uniform(void) { /* (-1, 1) */
  lcg_state = lcg_state * 1664525u + 1013904223u;
  return (float)((lcg_state >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

static float* tensor(size_t n, float scale, float offset) {
  float* p = (float*)malloc(n * sizeof(float));
  size_t i;
  if (!p) { fprintf(stderr, "out of memory\n"); exit(2); }
  for (i = 0; i < n; ++i) p[i] = offset + scale * lcg_uniform();
  return p;
}

int main(int argc, char** argv) {
  const int32_t chains = argc > 1 ? atoi(argv[1]) : 8;
  const int32_t T = argc > 2 ? atoi(argv[2]) : 100;
  /* the architecture of the shipped CATH model: hidden 384, 12 layers, 12 heads, intermediate 768, 128 positions */
  fd_dims d;
  const int32_t H = 384, L = 12, I = 768, P = 128, F = 6, N = 64;
  int32_t n_w, i, l, rc;
  const float** weights;
  float *time_table, *coef, *x0, *out;
  int32_t* lengths;
  uint8_t wrap[6] = {1, 1, 1, 1, 1, 1};
  fd_handle* h = NULL;
  double alphabar = 1.0, alphabar_prev = 1.0;

  if (chains < 1 || T < 2) { fprintf(stderr, "usage: %s [chains >= 1 [timesteps >= 2]]\n", argv[0]); return 2; }
  printf("library: ABI %d, %s\n", (int)fd_abi_version(), fd_build_info());
  if (fd_abi_version() != FD_ABI_VERSION) { fprintf(stderr, "header / library ABI mismatch\n"); return 2; }

  memset(&d, 0, sizeof d);
  d.hidden = H; d.layers = L; d.heads = H / 32; d.intermediate = I; d.max_pos = P; d.n_features = F;
  d.timesteps = T; d.ln_eps = 1e-12f; d.head_ln_eps = 1e-12f;

  n_w = fd_num_weights(L);
  if (n_w != FD_W_HEAD + L * FD_W_PER_LAYER + FD_W_TAIL) { fprintf(stderr, "unexpected weight count %d\n", (int)n_w); return 2; }
  weights = (const float**)calloc((size_t)n_w, sizeof(float*));
  if (!weights) return 2;
  i = 0;
  weights[i++] = tensor((size_t)H * F, 0.05f, 0.0f);      /* inputs_to_hidden_dim.weight */
  weights[i++] = tensor((size_t)H, 0.02f, 0.0f);          /* .bias */
  weights[i++] = tensor((size_t)H, 0.05f, 1.0f);          /* embeddings.LayerNorm.weight */
  weights[i++] = tensor((size_t)H, 0.02f, 0.0f);          /* .bias */
  for (l = 0; l < L; ++l) {
    int q;
    for (q = 0; q < 3; ++q) {                              /* query / key / value */
      weights[i++] = tensor((size_t)H * H, 0.035f, 0.0f);
      weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
    }
    weights[i++] = tensor((size_t)(2 * P - 1) * 32, 0.1f, 0.0f); /* distance_embedding */
    weights[i++] = tensor((size_t)H * H, 0.035f, 0.0f);    /* attention.output.dense */
    weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
    weights[i++] = tensor((size_t)H, 0.05f, 1.0f);         /* attention.output.LayerNorm */
    weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
    weights[i++] = tensor((size_t)I * H, 0.035f, 0.0f);    /* intermediate.dense */
    weights[i++] = tensor((size_t)I, 0.02f, 0.0f);
    weights[i++] = tensor((size_t)H * I, 0.035f, 0.0f);    /* output.dense */
    weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
    weights[i++] = tensor((size_t)H, 0.05f, 1.0f);         /* output.LayerNorm */
    weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
  }
  weights[i++] = tensor((size_t)H * H, 0.035f, 0.0f);     /* token_decoder.dense1 */
  weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
  weights[i++] = tensor((size_t)H, 0.05f, 1.0f);          /* token_decoder.layer_norm */
  weights[i++] = tensor((size_t)H, 0.02f, 0.0f);
  weights[i++] = tensor((size_t)F * H, 0.035f, 0.0f);     /* token_decoder.dense2 */
  weights[i++] = tensor((size_t)F, 0.02f, 0.0f);
  if (i != n_w) { fprintf(stderr, "weight list has %d entries, library expects %d\n", (int)i, (int)n_w); return 2; }

  /* time embedding rows [sin(2 pi t w), cos(2 pi t w)] with random frequencies (modelling.py:59-71) */
  time_table = (float*)malloc((size_t)T * H * sizeof(float));
  coef = (float*)malloc((size_t)T * 4 * sizeof(float));
  if (!time_table || !coef) return 2;
  {
    float* w = tensor((size_t)H / 2, 1.0f, 0.0f);
    int32_t t, k;
    for (t = 0; t < T; ++t)
      for (k = 0; k < H / 2; ++k) {
        const float a = (float)t * w[k] * 2.0f * 3.14159265358979f;
        time_table[(size_t)t * H + k] = sinf(a);
        time_table[(size_t)t * H + H / 2 + k] = cosf(a);
      }
    free(w);
    /* linear schedule 1e-4 .. 0.02 and {1/sqrt(alpha_t), beta_t, sqrt(1 - alphabar_t), sqrt(posterior variance_t)} */
    for (t = 0; t < T; ++t) {
      const double beta = 1e-4 + (0.02 - 1e-4) * (double)t / (double)(T - 1);
      const double alpha = 1.0 - beta;
      alphabar_prev = t == 0 ? 1.0 : alphabar;
      alphabar = alphabar_prev * alpha;
      coef[4 * t + 0] = (float)(1.0 / sqrt(alpha));
      coef[4 * t + 1] = (float)beta;
      coef[4 * t + 2] = (float)sqrt(1.0 - alphabar);
      coef[4 * t + 3] = (float)sqrt(beta * (1.0 - alphabar_prev) / (1.0 - alphabar));
    }
  }

  rc = fd_create(&d, weights, n_w, time_table, coef, /*device*/ 0, FD_GEMM_TC_3X, &h);
  if (rc != FD_OK) {
    fprintf(stderr, "fd_create failed (%d): %s\n", (int)rc, fd_last_error());
    return 1;
  }

  lengths = (int32_t*)malloc((size_t)chains * sizeof(int32_t));
  x0 = tensor((size_t)chains * N * F, 3.14159f, 0.0f);     /* any start in [-pi, pi) will do for a demonstration */
  out = (float*)malloc((size_t)chains * N * F * sizeof(float));
  if (!lengths || !out) return 2;
  for (i = 0; i < chains; ++i) lengths[i] = 40 + (i * 7) % (N - 40 + 1);

  /* noise_host == NULL: the library draws the per-step normals from its Philox stream `seed` */
  rc = fd_sample_host(h, chains, N, lengths, x0, T, NULL, 7344u, wrap, /*full_history*/ 0, out);
  if (rc != FD_OK) {
    fprintf(stderr, "fd_sample_host failed (%d): %s\n", (int)rc, fd_last_error());
    fd_destroy(h);
    return 1;
  }
  {
    double lo = 1e30, hi = -1e30;
    int32_t b, n, f;
    for (b = 0; b < chains; ++b)
      for (n = 0; n < lengths[b]; ++n)
        for (f = 0; f < F; ++f) {
          const double v = out[((size_t)b * N + n) * F + f];
          if (!(v == v)) { fprintf(stderr, "NaN in chain %d\n", (int)b); fd_destroy(h); return 1; }
          if (v < lo) lo = v;
          if (v > hi) hi = v;
        }
    printf("%d chains x %d reverse steps: angles in [%.6f, %.6f] (wrapped to [-pi, pi)), %lld kernel launches\n",
           (int)chains, (int)T, lo, hi, (long long)fd_launch_count(h));
  }
  fd_destroy(h);
  return 0;
}
