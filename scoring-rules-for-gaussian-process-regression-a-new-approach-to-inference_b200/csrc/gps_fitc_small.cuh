// One-warp dense helpers for the replicated M x M algebra of the FITC paths (rows / columns of the matrices
// live in registers; MP is a compile-time constant so every register array is statically indexed) and the
// m8n8k4 FP64 tensor instruction.  Shared by gps_fitc.cu and gps_fitc_fused.cu.
#pragma once
#include "gps_common.cuh"

namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// layout of the replicated small-matrix buffer (doubles), MP-strided
struct SmallLayout {
  int us, la, lai, kuu, lc, lci, beta, bbar, cbar, vyb, lainv, lcinv, fh, fhv, fh2, fgv, total;
  __host__ __device__ SmallLayout(int MP, int D) {
    int o = 0;
    us = o;   o += MP * D;
    la = o;   o += MP * MP;
    lai = o;  o += MP;
    kuu = o;  o += MP * MP;
    lc = o;   o += MP * MP;
    lci = o;  o += MP;
    beta = o; o += MP;
    bbar = o; o += MP;
    cbar = o; o += MP * MP;
    vyb = o;  o += MP;
    lainv = o; o += MP * MP;
    lcinv = o; o += MP * MP;
    fh = o;    o += 4 * MP * MP;   // per-fold Hhat_f (block objectives)
    fhv = o;   o += 4 * MP;        // per-fold h_f
    fh2 = o;   o += 4 * MP * MP;   // per-fold H_bar_f (kc)
    fgv = o;   o += 4 * MP;        // per-fold g_bar_f (kc)
    total = o;
  }
};

// ---- small dense helpers: ONE WARP, rows/columns of the M x M matrices live in registers ---------
// (MP is a compile-time constant, so every register array is statically indexed; lanes >= MP idle
// but take part in the shuffles).  All serial chains are one shuffle / multiply / fma long per step.

// In-place lower Cholesky of A[MP][MP] (shared, row-major).  Lane i owns row i.  Right-looking:
// pivot and column broadcast by shuffle.  Returns the first bad pivot (1-based) or 0.
template <int MP>
__device__ int warp_chol(double* A, int lane) {
  double r[MP];
#pragma unroll
  for (int j = 0; j < MP; ++j) r[j] = (lane < MP) ? A[lane * MP + j] : 0.0;
  // The pivot chain runs on a private copy of the lane's own diagonal entry: dg -= l_ij^2 needs only this lane's
  // l_ij, so the next pivot can be broadcast (and its rsqrt started) without waiting for the shuffle-driven update
  // of the other columns.  Serial chain per step: shfl -> rsqrt -> mul -> fma instead of shfl -> rsqrt -> mul ->
  // shfl -> fma behind 2 (MP - j) queued shuffles (4.0 -> 1.8 us for MP = 24, measured with the phase stamps).
  double dg = (lane < MP) ? A[lane * MP + lane] : 1.0;
  int bad = 0;
#pragma unroll
  for (int j = 0; j < MP; ++j) {
    double piv = __shfl_sync(0xffffffffu, dg, j);
    if (!(piv > 0.0)) {
      if (!bad) bad = j + 1;
      piv = 1.0;
    }
    const double rs = rsqrt(piv);
    const double lj = (lane == j) ? piv * rs : r[j] * rs;   // lanes > j: L[i][j]; lane j: sqrt(piv)
    dg = (lane > j) ? fma(-lj, lj, dg) : dg;
#pragma unroll
    for (int c = j + 1; c < MP; ++c) {
      const double lc = __shfl_sync(0xffffffffu, lj, c);
      r[c] = fma(-lj, lc, r[c]);                  // meaningful for lanes >= c (lower triangle)
    }
    r[j] = lj;
  }
  if (lane < MP) {
#pragma unroll
    for (int j = 0; j < MP; ++j) A[lane * MP + j] = (j <= lane) ? r[j] : 0.0;
  }
  __syncwarp();
  return bad;
}

// Solve L' x = b for the column held by this lane (b[] in registers); L lower in shared memory
// (warp-uniform addresses: broadcast reads), Li = 1 / diag(L).
template <int MP>
__device__ __forceinline__ void col_solve_LT(const double* __restrict__ L, const double* __restrict__ Li,
                                             double (&b)[MP]) {
  // volatile: keep the broadcast loads next to their fma instead of hoisting the whole triangle
  // into registers (the fully unrolled body otherwise spills for MP >= 24)
  const volatile double* Lv = L;
#pragma unroll
  for (int i = MP - 1; i >= 0; --i) {
    b[i] *= Li[i];
#pragma unroll
    for (int k = 0; k < i; ++k) b[k] = fma(-Lv[i * MP + k], b[i], b[k]);
  }
}

// X = L^-1 for lower-triangular L[MP][MP] in shared memory: lane c solves L x = e_c by
// column-oriented forward substitution and writes column c of X (row-major, zeros above the diagonal).
template <int MP>
__device__ __forceinline__ void warp_tri_inverse(const double* __restrict__ L, const double* __restrict__ Li,
                                                 double* __restrict__ Xout, int lane) {
  double x[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
  const volatile double* Lv = L;
#pragma unroll
  for (int i = 0; i < MP; ++i) {
    x[i] *= Li[i];
#pragma unroll
    for (int k = i + 1; k < MP; ++k) x[k] = fma(-Lv[k * MP + i], x[i], x[k]);
  }
  if (lane < MP) {
#pragma unroll
    for (int i = 0; i < MP; ++i) Xout[i * MP + lane] = (i >= lane) ? x[i] : 0.0;
  }
}

// Solve L x = v (forward) or L' x = v (backward) for ONE vector spread over the lanes (lane m holds
// v_m); returns x_m.
template <int MP>
__device__ __forceinline__ double warp_fwd_vec(const double* __restrict__ L, const double* __restrict__ Li,
                                               double v, int lane) {
#pragma unroll
  for (int j = 0; j < MP; ++j) {
    const double xj = __shfl_sync(0xffffffffu, v, j) * Li[j];
    if (lane == j) v = xj;
    else if (lane > j && lane < MP) v = fma(-L[lane * MP + j], xj, v);
  }
  return v;
}
template <int MP>
__device__ __forceinline__ double warp_bwd_vecT(const double* __restrict__ L, const double* __restrict__ Li,
                                                double v, int lane) {
#pragma unroll
  for (int j = MP - 1; j >= 0; --j) {
    const double xj = __shfl_sync(0xffffffffu, v, j) * Li[j];
    if (lane == j) v = xj;
    else if (lane < j) v = fma(-L[j * MP + lane], xj, v);
  }
  return v;
}

// Adjoint of a Cholesky factorisation: lb[] = column `lane` of Lbar (lower triangular) in, column
// `lane` of Abar = 0.5 L^-T (P + P') L^-1, P = Phi(L' Lbar), out.  T[MP][MP] shared scratch.
template <int MP>
__device__ __forceinline__ void warp_chol_adjoint(const double* __restrict__ L, const double* __restrict__ Li,
                                                  double* __restrict__ T, double (&lb)[MP], int lane) {
  double p[MP];
  const volatile double* Lv = L;
#pragma unroll
  for (int r = 0; r < MP; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = r; k < MP; ++k) s = fma(Lv[k * MP + r], lb[k], s);   // (L' Lbar)[r][lane]
    p[r] = (r > lane) ? s : ((r == lane) ? 0.5 * s : 0.0);
  }
  if (lane < MP) {
#pragma unroll
    for (int r = 0; r < MP; ++r) T[r * MP + lane] = p[r];
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < MP; ++r) p[r] += (lane < MP) ? T[lane * MP + r] : 0.0;   // Z = P + P'
  __syncwarp();
  col_solve_LT<MP>(L, Li, p);                                                    // Y = L^-T Z
  if (lane < MP) {
#pragma unroll
    for (int r = 0; r < MP; ++r) T[r * MP + lane] = p[r];
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < MP; ++r) p[r] = (lane < MP) ? T[lane * MP + r] : 0.0;     // column of Y'
  __syncwarp();
  col_solve_LT<MP>(L, Li, p);                                                    // (Y L^-1)' = symmetric
#pragma unroll
  for (int r = 0; r < MP; ++r) lb[r] = 0.5 * p[r];
}


}  // namespace
