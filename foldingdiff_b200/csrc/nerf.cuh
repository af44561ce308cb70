// Batched NeRF: internal angles -> backbone N / CA / C coordinates.
//
// Restates on the device what /root/reference/foldingdiff/nerf.py does per chain in Python
// (NERFBuilder.cartesian_coords :79-122 -> 3 (L-1) calls of place_dihedral :145-204, then the optional centring
// :124-128), as driven by angles_and_coords.create_new_chain_nerf (angles_and_coords.py:112-184): residue i+1's
// N, CA, C are placed from (psi_i, omega_i, phi_{i+1}) with bond angles (CA:C:1N)_i, (C:1N:1CA)_i, tau_i and the
// fixed bond lengths 1.34 / 1.46 / 1.54 A, starting from the first residue of 1CRN.
//
// Placement k depends on the three atoms before it, so a chain is a serial recurrence of 3 (L-1) steps; chains are
// independent.  One thread per chain in fp64 (the reference's coordinate arithmetic is float64): 512 chains x 381
// placements is ~0.1 GFLOP - microseconds of work, against ~20 ms per chain for the Python loop.
#pragma once
#include "common.cuh"

namespace fd {

struct NerfCols {
  int phi, psi, omega, tau, ca_c_n, c_n_ca;  // column of each angle in the feature axis, or -1 for the default
};

__device__ __forceinline__ void nerf_place(const double (&a)[3], const double (&b)[3], const double (&c)[3],
                                           double bond_angle, double bond_length, double torsion, double (&d)[3]) {
  const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  double bc[3] = {c[0] - b[0], c[1] - b[1], c[2] - b[2]};
  const double ibc = 1.0 / sqrt(bc[0] * bc[0] + bc[1] * bc[1] + bc[2] * bc[2]);
  bc[0] *= ibc; bc[1] *= ibc; bc[2] *= ibc;
  double n[3] = {ab[1] * bc[2] - ab[2] * bc[1], ab[2] * bc[0] - ab[0] * bc[2], ab[0] * bc[1] - ab[1] * bc[0]};
  const double in = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  n[0] *= in; n[1] *= in; n[2] *= in;
  const double nbc[3] = {n[1] * bc[2] - n[2] * bc[1], n[2] * bc[0] - n[0] * bc[2], n[0] * bc[1] - n[1] * bc[0]};
  double sa, ca, st, ct;
  sincos(bond_angle, &sa, &ca);
  sincos(torsion, &st, &ct);
  const double d0 = -bond_length * ca, d1 = bond_length * ct * sa, d2 = bond_length * st * sa;
  d[0] = bc[0] * d0 + nbc[0] * d1 + n[0] * d2 + c[0];
  d[1] = bc[1] * d0 + nbc[1] * d1 + n[1] * d2 + c[1];
  d[2] = bc[2] * d0 + nbc[2] * d1 + n[2] * d2 + c[2];
}

// angles (B, n_pad, F) fp32; lengths (B); coords (B, 3 * n_pad, 3) fp32, rows >= 3 * len are zero.
__global__ void nerf_kernel(const float* __restrict__ angles, const int* __restrict__ lengths, int batch, int n_pad,
                            int F, NerfCols cols, int center, float* __restrict__ coords) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int L = lengths[b];
  const float* ang = angles + (size_t)b * n_pad * F;
  float* out = coords + (size_t)b * n_pad * 9;
  double p0[3] = {17.047, 14.099, 3.625}, p1[3] = {16.967, 12.784, 4.338}, p2[3] = {15.685, 12.755, 5.133};
  double sum[3] = {p0[0] + p1[0] + p2[0], p0[1] + p1[1] + p2[1], p0[2] + p1[2] + p2[2]};
  for (int k = 0; k < 3; ++k) { out[k] = (float)p0[k]; out[3 + k] = (float)p1[k]; out[6 + k] = (float)p2[k]; }
  const double kTau = 109.0 / 180.0 * 3.14159265358979323846, kCaCN = 115.0 / 180.0 * 3.14159265358979323846,
               kCNCa = 121.0 / 180.0 * 3.14159265358979323846;
  for (int i = 0; i + 1 < L; ++i) {
    const float* r = ang + (size_t)i * F;
    const double psi = r[cols.psi], omega = r[cols.omega], phi = r[F + cols.phi];
    const double a_cn = cols.ca_c_n >= 0 ? (double)r[cols.ca_c_n] : kCaCN;
    const double a_nca = cols.c_n_ca >= 0 ? (double)r[cols.c_n_ca] : kCNCa;
    const double a_tau = cols.tau >= 0 ? (double)r[cols.tau] : kTau;
    double nn[3], ca[3], cc[3];
    nerf_place(p0, p1, p2, a_cn, 1.34, psi, nn);     // next N  from (N, CA, C)
    nerf_place(p1, p2, nn, a_nca, 1.46, omega, ca);  // next CA from (CA, C, N')
    nerf_place(p2, nn, ca, a_tau, 1.54, phi, cc);    // next C  from (C, N', CA')
    float* o = out + (size_t)(i + 1) * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[k] = (float)nn[k]; o[3 + k] = (float)ca[k]; o[6 + k] = (float)cc[k];
      sum[k] += nn[k] + ca[k] + cc[k];
      p0[k] = nn[k]; p1[k] = ca[k]; p2[k] = cc[k];
    }
  }
  if (center) {
    const double inv = 1.0 / (3.0 * L);
    const float m[3] = {(float)(sum[0] * inv), (float)(sum[1] * inv), (float)(sum[2] * inv)};
    for (int i = 0; i < 3 * L; ++i) { out[3 * i] -= m[0]; out[3 * i + 1] -= m[1]; out[3 * i + 2] -= m[2]; }
  }
  for (int i = 9 * L; i < 9 * n_pad; ++i) out[i] = 0.0f;
}

}  // namespace fd
