// FITC objective + gradient in O(N M^2): Woodbury restatement of the dense big_Q path
// (K20:222-236, K20:329-344, K20:434-452) with the three-pass analytic adjoint of SURVEY.md
// App. A.2, including the gradient wrt the inducing inputs (K20:247).
//
//   begin   replicated: Us = U / l, K_uu, L_A = chol(K_uu + jitter I)
//   pass 1  rows: k_i -> V_i = L_A^-1 k_i, lambda_i          reduce  C - I = sum V V'/lambda, v_y
//   pass 2  replicated: L_C, beta;  rows: W_i, d_i, alpha_i, score + seeds
//                                                             reduce  R = sum rbar W W', beta_bar, obj
//   pass 3  replicated: C_bar, vy_bar;  rows: lambda_bar, V_bar, Kuf_bar
//                                                             reduce  S = sum V_bar V', S0, P, g_b, sum lambda_bar
//   finish  replicated: L_A_bar -> A_bar, kernel-parameter and inducing-input gradients
//
// Rows are independent: a rank owns a contiguous block of rows and the three packed
// accumulators are what gets all-reduced.  One thread owns one row (its M-vectors
// live in registers, the M x M factors are broadcast from shared memory); the M x M outer-product
// reductions over the block's 128 rows run on the DMMA tensor path from transposed shared tiles.
// M is padded to MP = 8, 16, 24 or 32 (pad inducing points have k = 0 and an identity factor).
#include "gps_common.cuh"
#include "gps_fitc_small.cuh"

namespace {

constexpr int RB = 128;        // rows per block iteration = threads per block
constexpr int LDT = RB + 4;    // transposed tile stride: conflict-free for row-owner writes and DMMA loads
constexpr double INV_SQRT_PI = 0.56418958354775628695;
constexpr double INV_SQRT_2PI = 0.39894228040143267794;
constexpr double INV_SQRT2 = 0.70710678118654752440;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

// acc[e] = sum_b part[b][e] by the whole block, in block order (deterministic).  The loop is
// unrolled so that 16 independent loads are in flight per thread: the sum is latency-bound.
__device__ __forceinline__ void block_reduce_partials(const double* __restrict__ part, int nblocks, int len,
                                                      double* __restrict__ acc_out) {
  for (int e = threadIdx.x; e < len; e += blockDim.x) {
    double s = 0.0;
#pragma unroll 16
    for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * len + e];
    acc_out[e] = s;
  }
}

// ---- begin: replicated K_uu and its factor --------------------------------------------------------
template <int MP>
__global__ void __launch_bounds__(128)
fitc_small0_kernel(const double* __restrict__ U, const double* __restrict__ par, double* __restrict__ small,
                   int M, int D, double jitter, int* __restrict__ info) {
  __shared__ double A[MP * MP];
  __shared__ double us[MP * 16];
  const SmallLayout lo(MP, D);
  const int tid = threadIdx.x, lane = tid & 31;
  const double ea = par[0];
  for (int e = tid; e < MP * D; e += 128) {
    const int m = e / D, d = e - m * D;
    const double v = (m < M) ? U[m * D + d] * par[2 + d] : 0.0;
    us[e] = v;
    small[lo.us + e] = v;
  }
  __syncthreads();
  for (int e = tid; e < MP * MP; e += 128) {
    const int i = e / MP, j = e - i * MP;
    double v = 0.0;
    if (i < M && j < M) {
      double r2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const double df = us[i * D + d] - us[j * D + d];
        r2 = fma(df, df, r2);
      }
      v = ea * exp(-0.5 * r2);
    }
    small[lo.kuu + e] = v;
    A[e] = v + ((i == j) ? ((i < M) ? jitter : 1.0) : 0.0);
  }
  __syncthreads();
  if (tid < 32) {
    const int bad = warp_chol<MP>(A, lane);
    if (bad && lane == 0) atomicCAS(info, 0, bad);
  }
  __syncthreads();
  __shared__ double Ai[MP];
  for (int e = tid; e < MP * MP; e += 128) small[lo.la + e] = A[e];
  if (tid < MP) {
    Ai[tid] = 1.0 / A[tid * MP + tid];
    small[lo.lai + tid] = Ai[tid];
  }
  __syncthreads();
  if (tid < 32) warp_tri_inverse<MP>(A, Ai, small + lo.lainv, lane);   // L_A^-1 for the tile kernels
}

// per-row kernel vector k_m = ea exp(-0.5 |us_m - xs|^2); xs read from the transposed tile column `tid`
template <int MP>
__device__ __forceinline__ void kernel_vec(const double* __restrict__ Us, const double* __restrict__ XsT, int tid,
                                           int M, int D, double ea, double* k) {
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) {
      const double df = Us[m * D + d] - XsT[d * LDT + tid];
      r2 = fma(df, df, r2);
    }
    k[m] = (m < M) ? ea * exp(-0.5 * r2) : 0.0;
  }
}

// per-thread triangular solves, column oriented: after x_m is final it is eliminated from all later
// entries with independent fmas, so the serial chain is one multiply + one fma per step.
template <int MP>
__device__ __forceinline__ void fwd_subst(const double* __restrict__ L, const double* __restrict__ Li, double* v) {
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    v[m] *= Li[m];
#pragma unroll
    for (int j = m + 1; j < MP; ++j) v[j] = fma(-L[j * MP + m], v[m], v[j]);
  }
}

template <int MP>
__device__ __forceinline__ void bwd_subst_T(const double* __restrict__ L, const double* __restrict__ Li, double* v) {
#pragma unroll
  for (int m = MP - 1; m >= 0; --m) {
    v[m] *= Li[m];
#pragma unroll
    for (int j = 0; j < m; ++j) v[j] = fma(-L[m * MP + j], v[m], v[j]);
  }
}

// warp-level accumulation over this warp's 32 rows of the transposed tiles:
//   C[mf][nf] += sum_r A_T[m][r] * (B_T[n][r] * bscale[r])
template <int MFR, int NFR, bool SCALE>
__device__ __forceinline__ void tile_outer(const double* __restrict__ AT, const double* __restrict__ BT,
                                           const double* __restrict__ bscale, int warp, int lane,
                                           double (&acc)[MFR][NFR][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int r = warp * 32 + kk * 4 + t;
    double a[MFR], b[NFR];
#pragma unroll
    for (int i = 0; i < MFR; ++i) a[i] = AT[(i * 8 + g) * LDT + r];
    const double sc = SCALE ? bscale[r] : 1.0;
#pragma unroll
    for (int j = 0; j < NFR; ++j) b[j] = BT[(j * 8 + g) * LDT + r] * sc;
#pragma unroll
    for (int i = 0; i < MFR; ++i)
#pragma unroll
      for (int j = 0; j < NFR; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// column accumulation: c[mf] += sum_r A_T[m][r] * col[r]   (result sits in fragment column 0)
template <int MFR>
__device__ __forceinline__ void tile_col(const double* __restrict__ AT, const double* __restrict__ col, int warp,
                                         int lane, double (&acc)[MFR][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int r = warp * 32 + kk * 4 + t;
    const double b = (g == 0) ? col[r] : 0.0;
#pragma unroll
    for (int i = 0; i < MFR; ++i) dmma(acc[i][0], acc[i][1], AT[(i * 8 + g) * LDT + r], b);
  }
}

// sum the four warps' fragments into shared Cs[MROWS][ncols] (zeroed by the caller), serially per warp
template <int MFR, int NFR>
__device__ __forceinline__ void frags_to_smem(double (&acc)[MFR][NFR][2], double* Cs, int ncols, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int w = 0; w < 4; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < MFR; ++i)
#pragma unroll
        for (int j = 0; j < NFR; ++j) {
          Cs[(i * 8 + g) * ncols + j * 8 + 2 * t] += acc[i][j][0];
          Cs[(i * 8 + g) * ncols + j * 8 + 2 * t + 1] += acc[i][j][1];
        }
    }
    __syncthreads();
  }
}

template <int MFR>
__device__ __forceinline__ void colfrag_to_smem(double (&acc)[MFR][2], double* vs, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int w = 0; w < 4; ++w) {
    if (warp == w && t == 0) {
#pragma unroll
      for (int i = 0; i < MFR; ++i) vs[i * 8 + g] += acc[i][0];
    }
    __syncthreads();
  }
}

// ---- pass 1 ---------------------------------------------------------------------------------------
// part1[block] = [ sum V V'/lambda (MP*MP) | sum V y/lambda (MP) ]
template <int MP>
__global__ void __launch_bounds__(RB)
fitc_row1_kernel(const double* __restrict__ X, const double* __restrict__ y, int64_t N, int D, int M,
                 const double* __restrict__ par, const double* __restrict__ small, double* __restrict__ Vg,
                 double* __restrict__ lamg, double* __restrict__ part) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8;
  const SmallLayout lo(MP, D);
  double* Us = sh;                       // [MP][D]
  double* LA = Us + MP * D;              // [MP][MP]
  double* LAi = LA + MP * MP;            // [MP]
  double* XsT = LAi + MP;                // [D][LDT]
  double* WT = XsT + (size_t)D * LDT;    // [MP][LDT]   v / sqrt(lambda)
  double* Ys = WT + (size_t)MP * LDT;    // [RB]        y / sqrt(lambda)
  double* Cs = Ys + RB;                  // [MP][MP] + [MP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < MP * D; e += RB) Us[e] = small[lo.us + e];
  for (int e = tid; e < MP * MP; e += RB) LA[e] = small[lo.la + e];
  if (tid < MP) LAi[tid] = small[lo.lai + tid];
  for (int e = tid; e < MP * MP + MP; e += RB) Cs[e] = 0.0;
  const double ea = par[0], sn2 = par[1];
  double cacc[MF][MF][2], vacc[MF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    vacc[i][0] = vacc[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < MF; ++j) cacc[i][j][0] = cacc[i][j][1] = 0.0;
  }
  __syncthreads();
  for (int64_t base = (int64_t)blockIdx.x * RB; base < N; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < N;
    for (int d = 0; d < D; ++d) XsT[d * LDT + tid] = live ? X[i * D + d] * par[2 + d] : 0.0;
    double v[MP];
    kernel_vec<MP>(Us, XsT, tid, M, D, ea, v);
    fwd_subst<MP>(LA, LAi, v);
    double q = 0.0;
#pragma unroll
    for (int m = 0; m < MP; ++m) q = fma(v[m], v[m], q);
    const double lam = ea - q + sn2;
    const double rs = live ? rsqrt(lam) : 0.0;
    if (live) {
      lamg[i] = lam;
#pragma unroll
      for (int m = 0; m < MP; ++m) Vg[i * MP + m] = v[m];
    }
    // rsqrt() is approximate in fp64: refine once (Newton) so that rs^2 = 1/lambda to rounding
    const double rs2 = live ? rs * (1.5 - 0.5 * lam * rs * rs) : 0.0;
#pragma unroll
    for (int m = 0; m < MP; ++m) WT[m * LDT + tid] = v[m] * rs2;
    Ys[tid] = live ? y[i] * rs2 : 0.0;
    __syncwarp();
    tile_outer<MF, MF, false>(WT, WT, nullptr, warp, lane, cacc);
    tile_col<MF>(WT, Ys, warp, lane, vacc);
    __syncwarp();
  }
  frags_to_smem<MF, MF>(cacc, Cs, MP, warp, lane);
  colfrag_to_smem<MF>(vacc, Cs + MP * MP, warp, lane);
  for (int e = tid; e < MP * MP + MP; e += RB) part[(int64_t)blockIdx.x * (MP * MP + MP) + e] = Cs[e];
}

// acc = sum of the per-block partials (multi-GPU path: the result is handed to the all-reduce)
__global__ void __launch_bounds__(256)
fitc_reduce_kernel(const double* __restrict__ part, int nblocks, int len, double* __restrict__ acc) {
  block_reduce_partials(part, nblocks, len, acc);
}

// first stage of the reduction for large grids: blockIdx.y sums its slice of the partial blocks,
// blockIdx.x its slice of the entries; out[g][e].  Fixed slices: deterministic.
__global__ void __launch_bounds__(128)
fitc_reduce_stage_kernel(const double* __restrict__ part, int nblocks, int len, int per_group,
                         double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  const int b0 = blockIdx.y * per_group;
  const int b1 = min(nblocks, b0 + per_group);
  double s = 0.0;
#pragma unroll 16
  for (int b = b0; b < b1; ++b) s += part[(int64_t)b * len + e];
  out[(int64_t)blockIdx.y * len + e] = s;
}

// ---- replicated step before pass 2: C = I + acc1, L_C, beta -----------------------------------------
// If `part` is non-null the per-block partials of pass 1 are summed into acc1 first (single GPU:
// no separate reduce launch).
template <int MP>
__global__ void __launch_bounds__(128)
fitc_small1_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ acc1,
                   double* __restrict__ small, int D, int* __restrict__ info) {
  __shared__ double C[MP * MP];
  __shared__ double Li[MP];
  const SmallLayout lo(MP, D);
  const int tid = threadIdx.x, lane = tid & 31;
  if (part) {
    block_reduce_partials(part, nblocks, MP * MP + MP, acc1);
    __syncthreads();
  }
  for (int e = tid; e < MP * MP; e += 128) {
    const int i = e / MP, j = e - i * MP;
    C[e] = acc1[e] + (i == j ? 1.0 : 0.0);
  }
  __syncthreads();
  if (tid < 32) {
    const int bad = warp_chol<MP>(C, lane);
    if (bad && lane == 0) atomicCAS(info, 0, 1000000 + bad);
    if (lane < MP) Li[lane] = 1.0 / C[lane * MP + lane];
    __syncwarp();
    const double x = warp_fwd_vec<MP>(C, Li, lane < MP ? acc1[MP * MP + lane] : 0.0, lane);   // beta = L_C^-1 v_y
    if (lane < MP) {
      small[lo.beta + lane] = x;
      small[lo.lci + lane] = Li[lane];
    }
    warp_tri_inverse<MP>(C, Li, small + lo.lcinv, lane);   // L_C^-1 for the tile kernels
  }
  __syncthreads();
  for (int e = tid; e < MP * MP; e += 128) small[lo.lc + e] = C[e];
}

// ---- pass 2 ---------------------------------------------------------------------------------------
// part2[block] = [ sum rbar W W' (MP*MP) | sum tbar W (MP) | obj partial ]
template <int MP>
__global__ void __launch_bounds__(RB)
fitc_row2_kernel(const double* __restrict__ y, int64_t N, int D, int score, double invN,
                 const double* __restrict__ small, const double* __restrict__ Vg, double* __restrict__ Wg,
                 double* __restrict__ rowv, double* __restrict__ part) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8;
  const SmallLayout lo(MP, D);
  double* LC = sh;                  // [MP][MP]
  double* LCi = LC + MP * MP;       // [MP]
  double* beta = LCi + MP;          // [MP]
  double* WT = beta + MP;           // [MP][LDT]
  double* rbs = WT + (size_t)MP * LDT;  // [RB]
  double* tbs = rbs + RB;           // [RB]
  double* Cs = tbs + RB;            // [MP][MP] + [MP]
  double* red = Cs + MP * MP + MP;  // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < MP * MP; e += RB) LC[e] = small[lo.lc + e];
  if (tid < MP) {
    LCi[tid] = small[lo.lci + tid];
    beta[tid] = small[lo.beta + tid];
  }
  for (int e = tid; e < MP * MP + MP; e += RB) Cs[e] = 0.0;
  double racc[MF][MF][2], bacc[MF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    bacc[i][0] = bacc[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < MF; ++j) racc[i][j][0] = racc[i][j][1] = 0.0;
  }
  double obj = 0.0;
  __syncthreads();
  double* lamg = rowv;
  double* lb0g = rowv + N;
  double* rbg = rowv + 2 * N;
  double* tbg = rowv + 3 * N;
  double* alg = rowv + 4 * N;
  double* dg = rowv + 5 * N;
  for (int64_t base = (int64_t)blockIdx.x * RB; base < N; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < N;
    double w[MP];
#pragma unroll
    for (int m = 0; m < MP; ++m) w[m] = live ? Vg[i * MP + m] : 0.0;
    fwd_subst<MP>(LC, LCi, w);
    double rbar = 0.0, tbar = 0.0;
    if (live) {
      double r = 0.0, wb = 0.0;
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        r = fma(w[m], w[m], r);
        wb = fma(w[m], beta[m], wb);
      }
      const double lam = lamg[i], yi = y[i];
      const double il = 1.0 / lam;
      const double d = il - r * il * il;
      const double alpha = (yi - wb) * il;
      double abar, dbar, lb0 = 0.0;
      if (score == GPS_CRPS) {
        const double s2 = 1.0 / d, s = sqrt(s2), z = alpha * s;
        const double tpm1 = erf(z * INV_SQRT2);
        const double g = z * tpm1 + 2.0 * INV_SQRT_2PI * exp(-0.5 * z * z) - INV_SQRT_PI;
        obj += s * g * invN;
        abar = tpm1 * s2 * invN;
        dbar = -(0.5 * s2 * s * g + 0.5 * tpm1 * alpha * s2 * s2) * invN;
      } else if (score == GPS_LOGS) {
        const double s2 = 1.0 / d;
        obj += (0.5 * alpha * alpha * s2 - 0.5 * log(d) + HALF_LOG_2PI) * invN;
        abar = alpha * s2 * invN;
        dbar = -(0.5 * alpha * alpha * s2 * s2 + 0.5 * s2) * invN;
      } else {  // NLML: 0.5 log lambda + 0.5 y alpha per row (K20:337-340 through Woodbury)
        obj += 0.5 * log(lam) + 0.5 * yi * alpha;
        abar = 0.5 * yi;
        dbar = 0.0;
        lb0 = 0.5 * il;
      }
      lb0 += dbar * (-il * il + 2.0 * r * il * il * il) - abar * alpha * il;
      rbar = -dbar * il * il;
      tbar = -abar * il;
      lb0g[i] = lb0;
      rbg[i] = rbar;
      tbg[i] = tbar;
      alg[i] = alpha;
      dg[i] = d;
#pragma unroll
      for (int m = 0; m < MP; ++m) Wg[i * MP + m] = w[m];
    }
#pragma unroll
    for (int m = 0; m < MP; ++m) WT[m * LDT + tid] = w[m];
    rbs[tid] = rbar;
    tbs[tid] = tbar;
    __syncwarp();
    tile_outer<MF, MF, true>(WT, WT, rbs, warp, lane, racc);
    tile_col<MF>(WT, tbs, warp, lane, bacc);
    __syncwarp();
  }
  frags_to_smem<MF, MF>(racc, Cs, MP, warp, lane);
  colfrag_to_smem<MF>(bacc, Cs + MP * MP, warp, lane);
  const double o = block_sum(obj, red);
  const int len = MP * MP + MP + 1;
  for (int e = tid; e < MP * MP + MP; e += RB) part[(int64_t)blockIdx.x * len + e] = Cs[e];
  if (tid == 0) part[(int64_t)blockIdx.x * len + MP * MP + MP] = o;
}

// ---- replicated step before pass 3: C_bar, vy_bar ----------------------------------------------------
template <int MP>
__global__ void __launch_bounds__(128)
fitc_small2_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ acc2,
                   double* __restrict__ small, int M, int D, int score) {
  __shared__ double LC[MP * MP];
  __shared__ double T[MP * MP];
  __shared__ double Li[MP], beta[MP], bb[MP];
  const SmallLayout lo(MP, D);
  const int tid = threadIdx.x, lane = tid & 31;
  if (part) {
    block_reduce_partials(part, nblocks, MP * MP + MP + 1, acc2);
    __syncthreads();
  }
  for (int e = tid; e < MP * MP; e += 128) LC[e] = small[lo.lc + e];
  if (tid < MP) {
    Li[tid] = small[lo.lci + tid];
    beta[tid] = small[lo.beta + tid];
    bb[tid] = acc2[MP * MP + tid];
  }
  __syncthreads();
  if (tid < 32) {
    const int c = lane < MP ? lane : 0;
    double sw[MP];
#pragma unroll
    for (int r = 0; r < MP; ++r)   // column c of S_W = beta beta_bar' + 2 R + beta_bar beta'
      sw[r] = (lane < MP) ? beta[r] * bb[c] + 2.0 * acc2[r * MP + c] + bb[r] * beta[c] : 0.0;
    col_solve_LT<MP>(LC, Li, sw);                                     // L_C^-T S_W
#pragma unroll
    for (int r = 0; r < MP; ++r) {
      double v = (r >= lane && lane < MP) ? -sw[r] : 0.0;
      if (score == GPS_NLML && r == lane && lane < M) v += Li[r];    // d/dL_C of sum log diag(L_C)
      sw[r] = v;
    }
    warp_chol_adjoint<MP>(LC, Li, T, sw, lane);
    if (lane < MP) {
#pragma unroll
      for (int r = 0; r < MP; ++r) small[lo.cbar + r * MP + lane] = sw[r];
    }
    const double vb = warp_bwd_vecT<MP>(LC, Li, lane < MP ? bb[lane] : 0.0, lane);   // vy_bar = L_C^-T beta_bar
    if (lane < MP) {
      small[lo.vyb + lane] = vb;
      small[lo.bbar + lane] = bb[lane];
    }
  }
}

// ---- pass 3 ---------------------------------------------------------------------------------------
// part3[block] = [ S = sum Vbar V' (MP*MP) | P = sum G [xs | 1] (MP * 8*NF) | g_b rows (D) | sum lambda_bar ]
template <int MP, int NF>
__global__ void __launch_bounds__(RB)
fitc_row3_kernel(const double* __restrict__ X, const double* __restrict__ y, int64_t N, int D, int M,
                 const double* __restrict__ par, const double* __restrict__ small,
                 const double* __restrict__ Vg, const double* __restrict__ Wg, const double* __restrict__ rowv,
                 double* __restrict__ part) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8;
  constexpr int PC = 8 * NF;         // columns of P: xs_0..xs_{D-1}, 1, zero padding
  const SmallLayout lo(MP, D);
  double* Us = sh;                          // [MP][D]
  double* LA = Us + MP * D;                 // [MP][MP]
  double* LAi = LA + MP * MP;               // [MP]
  double* LC = LAi + MP;                    // [MP][MP]
  double* LCi = LC + MP * MP;               // [MP]
  double* Cb = LCi + MP;                    // [MP][MP]
  double* beta = Cb + MP * MP;              // [MP]
  double* bbar = beta + MP;                 // [MP]
  double* vyb = bbar + MP;                  // [MP]
  double* XsT = vyb + MP;                   // [PC][LDT]
  double* VT = XsT + (size_t)PC * LDT;      // [MP][LDT]   V
  double* VbT = VT + (size_t)MP * LDT;      // [MP][LDT]   V_bar
  double* GT = VbT + (size_t)MP * LDT;      // [MP][LDT]   G = Kuf_bar o Kuf
  double* gbs = GT + (size_t)MP * LDT;      // [D][RB]     per-thread g_b partials
  double* Cs = gbs + (size_t)D * RB;        // [MP][MP] + [MP][PC]
  double* red = Cs + MP * MP + MP * PC;     // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < MP * D; e += RB) Us[e] = small[lo.us + e];
  for (int e = tid; e < MP * MP; e += RB) {
    LA[e] = small[lo.la + e];
    LC[e] = small[lo.lc + e];
    Cb[e] = small[lo.cbar + e];
  }
  if (tid < MP) {
    LAi[tid] = small[lo.lai + tid];
    LCi[tid] = small[lo.lci + tid];
    beta[tid] = small[lo.beta + tid];
    bbar[tid] = small[lo.bbar + tid];
    vyb[tid] = small[lo.vyb + tid];
  }
  for (int e = tid; e < MP * MP + MP * PC; e += RB) Cs[e] = 0.0;
  for (int d = 0; d < D; ++d) gbs[d * RB + tid] = 0.0;
  for (int c = D + 1; c < PC; ++c) XsT[c * LDT + tid] = 0.0;
  double sacc[MF][MF][2], pacc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
#pragma unroll
    for (int j = 0; j < MF; ++j) sacc[i][j][0] = sacc[i][j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NF; ++j) pacc[i][j][0] = pacc[i][j][1] = 0.0;
  }
  double sum_lb = 0.0;
  const double ea = par[0];
  const double* lamg = rowv;
  const double* lb0g = rowv + N;
  const double* rbg = rowv + 2 * N;
  const double* tbg = rowv + 3 * N;
  __syncthreads();
  for (int64_t base = (int64_t)blockIdx.x * RB; base < N; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < N;
    for (int d = 0; d < D; ++d) XsT[d * LDT + tid] = live ? X[i * D + d] * par[2 + d] : 0.0;
    XsT[D * LDT + tid] = live ? 1.0 : 0.0;
    double v[MP], cv[MP], w[MP];
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      v[m] = live ? Vg[i * MP + m] : 0.0;
      w[m] = live ? Wg[i * MP + m] : 0.0;
      VT[m * LDT + tid] = v[m];
    }
    const double lam = live ? lamg[i] : 1.0, yi = live ? y[i] : 0.0;
    const double il = 1.0 / lam;
    const double rbar = live ? rbg[i] : 0.0, tbar = live ? tbg[i] : 0.0;
    // cv = C_bar v ; s1 = v' C_bar v ; bw = beta_bar' w
    double s1 = 0.0, bw = 0.0;
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < MP; ++j) s = fma(Cb[m * MP + j], v[j], s);
      cv[m] = s;
      s1 = fma(v[m], s, s1);
      bw = fma(bbar[m], w[m], bw);
    }
    const double lb = live ? (lb0g[i] - bw * yi * il * il - s1 * il * il) : 0.0;
    sum_lb += lb;
    // W_bar = tbar beta + 2 rbar w  (in place), then V_bar = L_C^-T W_bar + vy_bar y/lam + 2 cv/lam - 2 lb v
#pragma unroll
    for (int m = 0; m < MP; ++m) w[m] = fma(tbar, beta[m], 2.0 * rbar * w[m]);
    bwd_subst_T<MP>(LC, LCi, w);
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      w[m] += vyb[m] * yi * il + 2.0 * cv[m] * il - 2.0 * lb * v[m];
      VbT[m * LDT + tid] = w[m];
    }
    // Kuf_bar = L_A^-T V_bar (in place), G = Kuf_bar o Kuf (kept in w[]), g_b rows
    bwd_subst_T<MP>(LA, LAi, w);
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      double r2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const double df = Us[m * D + d] - XsT[d * LDT + tid];
        r2 = fma(df, df, r2);
      }
      w[m] = (m < M && live) ? w[m] * ea * exp(-0.5 * r2) : 0.0;
      GT[m * LDT + tid] = w[m];
    }
    for (int d = 0; d < D; ++d) {
      const double xd = XsT[d * LDT + tid];
      double sgb = 0.0;
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        const double df = Us[m * D + d] - xd;
        sgb = fma(w[m], df * df, sgb);
      }
      gbs[d * RB + tid] += sgb;
    }
    __syncwarp();
    tile_outer<MF, MF, false>(VbT, VT, nullptr, warp, lane, sacc);
    tile_outer<MF, NF, false>(GT, XsT, nullptr, warp, lane, pacc);
    __syncwarp();
  }
  frags_to_smem<MF, MF>(sacc, Cs, MP, warp, lane);
  frags_to_smem<MF, NF>(pacc, Cs + MP * MP, PC, warp, lane);
  const int len = MP * MP + MP * PC + D + 1;
  double* out = part + (int64_t)blockIdx.x * len;
  for (int e = tid; e < MP * MP + MP * PC; e += RB) out[e] = Cs[e];
  for (int d = 0; d < D; ++d) {
    const double s = block_sum(gbs[d * RB + tid], red);
    if (tid == 0) out[MP * MP + MP * PC + d] = s;
  }
  const double s = block_sum(sum_lb, red);
  if (tid == 0) out[MP * MP + MP * PC + D] = s;
}

// ---- finish: L_A_bar -> A_bar, all gradients --------------------------------------------------------
// out = [obj | g_theta (D+2) | g_U (M*D)]
template <int MP>
__global__ void __launch_bounds__(128)
fitc_small3_kernel(const double* __restrict__ part, int nblocks, const double* __restrict__ acc2,
                   double* __restrict__ acc3, const double* __restrict__ small, const double* __restrict__ par,
                   int M, int D, int PC, int score, double world_n, double* __restrict__ out) {
  __shared__ double LA[MP * MP];
  __shared__ double T[MP * MP];
  __shared__ double G2[MP * MP];
  __shared__ double Li[MP];
  __shared__ double us[MP * 16];
  __shared__ double red[32];
  const SmallLayout lo(MP, D);
  const int tid = threadIdx.x, lane = tid & 31;
  if (part) {
    block_reduce_partials(part, nblocks, MP * MP + MP * PC + D + 1, acc3);
    __syncthreads();
  }
  for (int e = tid; e < MP * MP; e += 128) LA[e] = small[lo.la + e];
  for (int e = tid; e < MP * D; e += 128) us[e] = small[lo.us + e];
  if (tid < MP) Li[tid] = small[lo.lai + tid];
  __syncthreads();
  if (tid < 32) {
    const int c = lane < MP ? lane : 0;
    double sc[MP];
#pragma unroll
    for (int r = 0; r < MP; ++r) sc[r] = (lane < MP) ? acc3[r * MP + c] : 0.0;   // column c of S
    col_solve_LT<MP>(LA, Li, sc);                                                  // L_A^-T S
#pragma unroll
    for (int r = 0; r < MP; ++r) sc[r] = (r >= lane && lane < MP) ? -sc[r] : 0.0;  // L_A_bar
    warp_chol_adjoint<MP>(LA, Li, T, sc, lane);                                    // A_bar
    if (lane < MP) {
#pragma unroll
      for (int r = 0; r < MP; ++r) G2[r * MP + lane] = sc[r] * small[lo.kuu + r * MP + lane];   // A_bar o K_uu
    }
  }
  __syncthreads();
  const double* P = acc3 + MP * MP;          // [MP][PC]: columns 0..D-1 = sum G xs_d, column D = sum G
  const double* gbrow = P + MP * PC;         // [D]
  const double sum_lb = gbrow[D];
  const double ea = par[0], sn2 = par[1];
  double sa = 0.0;
  for (int e = tid; e < MP * MP; e += 128) sa += G2[e];
  for (int m = tid; m < MP; m += 128) sa += P[m * PC + D];
  sa = block_sum(sa, red);
  if (tid == 0) {
    double obj = acc2[MP * MP + MP];
    if (score == GPS_NLML) {
      obj += 0.5 * world_n * 1.83787706640934548356;               // N/2 log 2 pi
      for (int m = 0; m < M; ++m) obj -= log(small[lo.lci + m]);   // + sum log diag(L_C)
    }
    out[0] = obj;
    out[1] = ea * sum_lb + sa;
    out[1 + D + 1] = sn2 * sum_lb;
  }
  for (int d = 0; d < D; ++d) {
    double sb = 0.0;
    for (int e = tid; e < MP * MP; e += 128) {
      const int i = e / MP, j = e - i * MP;
      const double df = us[i * D + d] - us[j * D + d];
      sb = fma(G2[e], df * df, sb);
    }
    sb = block_sum(sb, red);
    if (tid == 0) out[2 + d] = gbrow[d] + sb;
  }
  // g_U[m][d] = -invl_d ( us_md S0_m - P_md ) - 2 invl_d sum_m' G2_mm' (us_md - us_m'd)
  for (int e = tid; e < M * D; e += 128) {
    const int m = e / D, d = e - m * D;
    double t = 0.0;
    for (int j = 0; j < MP; ++j) t = fma(G2[m * MP + j], us[m * D + d] - us[j * D + d], t);
    out[1 + D + 2 + e] = -par[2 + d] * (us[m * D + d] * P[m * PC + D] - P[m * PC + d]) - 2.0 * par[2 + d] * t;
  }
}

// =====================================================================================================
// Tile formulation of the three row passes (default).  Same math, same accumulators, but every
// M x M operation on a block of 128 rows is a small GEMM on the DMMA tensor path:
//     V = K L_A^-T,   W = V L_C^-T,   CV = V C_bar,   V_bar = W_bar L_C^-1 + ...,   Kuf_bar = V_bar L_A^-1
// Row tiles live transposed in shared memory ([MP][132]: a row owner walks a column conflict-free,
// the m8n8k4 A-fragment loads are conflict-free); the M x M factors are staged once per CTA as
// B operands ([MP][MP + 4]).  Per-row scalars (lambda, d, alpha, seeds) stay thread-per-row.  V and W
// are kept in HBM transposed ([MP][N]) so tile loads and stores are coalesced.  Registers per thread
// drop from 255 to well under 128 and nothing serial is longer than one fragment product.
// =====================================================================================================
template <int MP>
struct TileCfg {
  static constexpr int MF = 4;            // 32 rows per warp
  static constexpr int NF = MP / 8;
  static constexpr int KS = MP / 4;
  static constexpr int LDM = MP + 4;      // B-operand stride: conflict-free fragment loads
};

// out[r][n] = sum_k in[r][k] * Mat[k][n] for this warp's 32 rows; tiles transposed ([MP][LDT]).
// in == out is allowed (all reads complete before the first write).
template <int MP>
__device__ __forceinline__ void warp_tile_mm(const double* inT, const double* __restrict__ Mat, double* outT,
                                             int warp, int lane) {
  using C = TileCfg<MP>;
  const int g = lane >> 2, t = lane & 3, r0 = warp * 32;
  double acc[C::MF][C::NF][2];
#pragma unroll
  for (int i = 0; i < C::MF; ++i)
#pragma unroll
    for (int j = 0; j < C::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < C::KS; ++kk) {
    double a[C::MF], b[C::NF];
#pragma unroll
    for (int i = 0; i < C::MF; ++i) a[i] = inT[(kk * 4 + t) * LDT + r0 + i * 8 + g];
#pragma unroll
    for (int j = 0; j < C::NF; ++j) b[j] = Mat[(kk * 4 + t) * C::LDM + j * 8 + g];
#pragma unroll
    for (int i = 0; i < C::MF; ++i)
#pragma unroll
      for (int j = 0; j < C::NF; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < C::MF; ++i)
#pragma unroll
    for (int j = 0; j < C::NF; ++j) {
      outT[(j * 8 + 2 * t) * LDT + r0 + i * 8 + g] = acc[i][j][0];
      outT[(j * 8 + 2 * t + 1) * LDT + r0 + i * 8 + g] = acc[i][j][1];
    }
  __syncwarp();
}

// stage an M x M matrix from the replicated buffer as a B operand: Mat[k][n] = src[k][n] or src[n][k]
template <int MP, bool TRANSPOSE>
__device__ __forceinline__ void stage_mat(double* Mat, const double* __restrict__ src, int tid) {
  constexpr int LDM = TileCfg<MP>::LDM;
  for (int e = tid; e < MP * MP; e += RB) {
    const int k = e / MP, n = e - k * MP;
    Mat[k * LDM + n] = TRANSPOSE ? src[n * MP + k] : src[k * MP + n];
  }
}

template <int MP>
__global__ void __launch_bounds__(RB)
fitc_row1_tile_kernel(const double* __restrict__ X, const double* __restrict__ y, int64_t N, int D, int M,
                      const double* __restrict__ par, const double* __restrict__ small, double* __restrict__ VgT,
                      double* __restrict__ lamg, double* __restrict__ part) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8, LDM = TileCfg<MP>::LDM;
  const SmallLayout lo(MP, D);
  double* Us = sh;                        // [MP][D]
  double* MatA = Us + MP * D;             // [MP][LDM]  B[k][n] = L_A^-1[n][k]
  double* XsT = MatA + MP * LDM;          // [D][LDT]
  double* KT = XsT + (size_t)D * LDT;     // [MP][LDT]   k -> V -> V / sqrt(lambda)
  double* Ys = KT + (size_t)MP * LDT;     // [RB]
  double* Cs = Ys + RB;                   // [MP][MP] + [MP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < MP * D; e += RB) Us[e] = small[lo.us + e];
  stage_mat<MP, true>(MatA, small + lo.lainv, tid);
  for (int e = tid; e < MP * MP + MP; e += RB) Cs[e] = 0.0;
  const double ea = par[0], sn2 = par[1];
  double cacc[MF][MF][2], vacc[MF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    vacc[i][0] = vacc[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < MF; ++j) cacc[i][j][0] = cacc[i][j][1] = 0.0;
  }
  __syncthreads();
  for (int64_t base = (int64_t)blockIdx.x * RB; base < N; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < N;
    for (int d = 0; d < D; ++d) XsT[d * LDT + tid] = live ? X[i * D + d] * par[2 + d] : 0.0;
#pragma unroll 4
    for (int m = 0; m < MP; ++m) {
      double r2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const double df = Us[m * D + d] - XsT[d * LDT + tid];
        r2 = fma(df, df, r2);
      }
      KT[m * LDT + tid] = (m < M && live) ? ea * exp(-0.5 * r2) : 0.0;
    }
    __syncwarp();
    warp_tile_mm<MP>(KT, MatA, KT, warp, lane);                 // V = K L_A^-T
    double q = 0.0;
#pragma unroll 8
    for (int m = 0; m < MP; ++m) {
      const double v = KT[m * LDT + tid];
      q = fma(v, v, q);
      if (live) VgT[(int64_t)m * N + i] = v;
    }
    const double lam = ea - q + sn2;
    if (live) lamg[i] = lam;
    const double rs = live ? 1.0 / sqrt(lam) : 0.0;
#pragma unroll 8
    for (int m = 0; m < MP; ++m) KT[m * LDT + tid] *= rs;
    Ys[tid] = live ? y[i] * rs : 0.0;
    __syncwarp();
    tile_outer<MF, MF, false>(KT, KT, nullptr, warp, lane, cacc);
    tile_col<MF>(KT, Ys, warp, lane, vacc);
    __syncwarp();
  }
  frags_to_smem<MF, MF>(cacc, Cs, MP, warp, lane);
  colfrag_to_smem<MF>(vacc, Cs + MP * MP, warp, lane);
  for (int e = tid; e < MP * MP + MP; e += RB) part[(int64_t)blockIdx.x * (MP * MP + MP) + e] = Cs[e];
}

template <int MP>
__global__ void __launch_bounds__(RB)
fitc_row2_tile_kernel(const double* __restrict__ y, int64_t N, int D, int score, double invN,
                      const double* __restrict__ small, const double* __restrict__ VgT, double* __restrict__ WgT,
                      double* __restrict__ rowv, double* __restrict__ part) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8, LDM = TileCfg<MP>::LDM;
  const SmallLayout lo(MP, D);
  double* MatC = sh;                      // [MP][LDM]  B[k][n] = L_C^-1[n][k]
  double* beta = MatC + MP * LDM;         // [MP]
  double* WT = beta + MP;                 // [MP][LDT]
  double* rbs = WT + (size_t)MP * LDT;    // [RB]
  double* tbs = rbs + RB;                 // [RB]
  double* Cs = tbs + RB;                  // [MP][MP] + [MP]
  double* red = Cs + MP * MP + MP;        // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_mat<MP, true>(MatC, small + lo.lcinv, tid);
  if (tid < MP) beta[tid] = small[lo.beta + tid];
  for (int e = tid; e < MP * MP + MP; e += RB) Cs[e] = 0.0;
  double racc[MF][MF][2], bacc[MF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    bacc[i][0] = bacc[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < MF; ++j) racc[i][j][0] = racc[i][j][1] = 0.0;
  }
  double obj = 0.0;
  __syncthreads();
  double* lamg = rowv;
  double* lb0g = rowv + N;
  double* rbg = rowv + 2 * N;
  double* tbg = rowv + 3 * N;
  double* alg = rowv + 4 * N;
  double* dg = rowv + 5 * N;
  for (int64_t base = (int64_t)blockIdx.x * RB; base < N; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < N;
#pragma unroll 8
    for (int m = 0; m < MP; ++m) WT[m * LDT + tid] = live ? VgT[(int64_t)m * N + i] : 0.0;
    __syncwarp();
    warp_tile_mm<MP>(WT, MatC, WT, warp, lane);                 // W = V L_C^-T
    double rbar = 0.0, tbar = 0.0;
    if (live) {
      double r = 0.0, wb = 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) {
        const double w = WT[m * LDT + tid];
        r = fma(w, w, r);
        wb = fma(w, beta[m], wb);
        WgT[(int64_t)m * N + i] = w;
      }
      const double lam = lamg[i], yi = y[i];
      const double il = 1.0 / lam;
      const double d = il - r * il * il;
      const double alpha = (yi - wb) * il;
      double abar, dbar, lb0 = 0.0;
      if (score == GPS_CRPS) {
        const double s2 = 1.0 / d, s = sqrt(s2), z = alpha * s;
        const double tpm1 = erf(z * INV_SQRT2);
        const double g = z * tpm1 + 2.0 * INV_SQRT_2PI * exp(-0.5 * z * z) - INV_SQRT_PI;
        obj += s * g * invN;
        abar = tpm1 * s2 * invN;
        dbar = -(0.5 * s2 * s * g + 0.5 * tpm1 * alpha * s2 * s2) * invN;
      } else if (score == GPS_LOGS) {
        const double s2 = 1.0 / d;
        obj += (0.5 * alpha * alpha * s2 - 0.5 * log(d) + HALF_LOG_2PI) * invN;
        abar = alpha * s2 * invN;
        dbar = -(0.5 * alpha * alpha * s2 * s2 + 0.5 * s2) * invN;
      } else {
        obj += 0.5 * log(lam) + 0.5 * yi * alpha;
        abar = 0.5 * yi;
        dbar = 0.0;
        lb0 = 0.5 * il;
      }
      lb0 += dbar * (-il * il + 2.0 * r * il * il * il) - abar * alpha * il;
      rbar = -dbar * il * il;
      tbar = -abar * il;
      lb0g[i] = lb0;
      rbg[i] = rbar;
      tbg[i] = tbar;
      alg[i] = alpha;
      dg[i] = d;
    }
    rbs[tid] = rbar;
    tbs[tid] = tbar;
    __syncwarp();
    tile_outer<MF, MF, true>(WT, WT, rbs, warp, lane, racc);
    tile_col<MF>(WT, tbs, warp, lane, bacc);
    __syncwarp();
  }
  frags_to_smem<MF, MF>(racc, Cs, MP, warp, lane);
  colfrag_to_smem<MF>(bacc, Cs + MP * MP, warp, lane);
  const double o = block_sum(obj, red);
  const int len = MP * MP + MP + 1;
  for (int e = tid; e < MP * MP + MP; e += RB) part[(int64_t)blockIdx.x * len + e] = Cs[e];
  if (tid == 0) part[(int64_t)blockIdx.x * len + MP * MP + MP] = o;
}

struct FoldGeom { int64_t lo[4], hi[4]; };   // local row range of each of the four folds

// OBJ = 0: LOO scores (seeds lambda_bar0, r_bar, t_bar come from pass 2 through rowv)
// OBJ = 1: 4-fold DSS: blockIdx.y is the fold; the seeds are formed here from (lambda, alpha, W)
//          and the fold's Hhat_f, h_f:  abar = lam alpha + W'h,  D = -2 Hhat W / lam + alpha h.
template <int MP, int NF, int OBJ>
__global__ void __launch_bounds__(RB)
fitc_row3_tile_kernel(const double* __restrict__ X, const double* __restrict__ y, int64_t N, int D, int M,
                      const double* __restrict__ par, const double* __restrict__ small,
                      const double* __restrict__ VgT, const double* __restrict__ WgT,
                      const double* __restrict__ rowv, double* __restrict__ part, FoldGeom fg) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8, LDM = TileCfg<MP>::LDM;
  constexpr int PC = 8 * NF;
  const SmallLayout lo(MP, D);
  double* Us = sh;                          // [MP][D]
  double* MatCb = Us + MP * D;              // [MP][LDM]  C_bar
  double* MatLC = MatCb + MP * LDM;         // [MP][LDM]  B[k][n] = L_C^-1[k][n]
  double* MatLA = MatLC + MP * LDM;         // [MP][LDM]  B[k][n] = L_A^-1[k][n]
  double* beta = MatLA + MP * LDM;          // [MP]
  double* bbar = beta + MP;                 // [MP]
  double* vyb = bbar + MP;                  // [MP]
  double* XsT = vyb + MP;                   // [PC][LDT]
  double* VT = XsT + (size_t)PC * LDT;      // [MP][LDT]  V
  double* WT = VT + (size_t)MP * LDT;       // [MP][LDT]  W -> W_bar -> V_bar
  double* CT = WT + (size_t)MP * LDT;       // [MP][LDT]  C_bar V  ->  Kuf_bar -> G
  double* Cs = CT;                          // [MP][MP] + [MP][PC]: the block's sums, staged over CT after the loop
  double* red = CT + (size_t)MP * LDT;      // [32]
  // fold matrices of the block objectives only (OBJ = 0 stops here: 2 CTAs per SM at M <= 24, D <= 8)
  double* MatH = red + 32;                  // [MP][LDM]  Hhat_f (OBJ = 1) / H_bar_f (OBJ = 2)
  double* hv = MatH + MP * LDM;             // [MP]       h_f
  double* MatHi = hv + MP;                  // [MP][LDM]  H_f^-1          (OBJ = 2)
  double* gbv = MatHi + MP * LDM;           // [MP]       g_bar_f         (OBJ = 2)
  double* gbs = (OBJ ? gbv + MP : red + 32);   // [D][4]  g_b accumulators, one per warp (warp-reduced every tile)
  static_assert(MP * MP + MP * PC <= MP * LDT, "block sums must fit in one tile");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fold = OBJ ? blockIdx.y : 0;
  const int64_t row_lo = OBJ ? fg.lo[fold] : 0, row_hi = OBJ ? fg.hi[fold] : N;
  if (OBJ == 1) stage_mat<MP, false>(MatH, small + lo.fh + fold * MP * MP, tid);
  if (OBJ == 2) {
    stage_mat<MP, false>(MatH, small + lo.fh2 + fold * MP * MP, tid);
    stage_mat<MP, false>(MatHi, small + lo.fh + fold * MP * MP, tid);
    if (tid < MP) gbv[tid] = small[lo.fgv + fold * MP + tid];
  }
  if (OBJ && tid < MP) hv[tid] = small[lo.fhv + fold * MP + tid];
  for (int e = tid; e < MP * D; e += RB) Us[e] = small[lo.us + e];
  stage_mat<MP, false>(MatCb, small + lo.cbar, tid);
  stage_mat<MP, false>(MatLC, small + lo.lcinv, tid);
  stage_mat<MP, false>(MatLA, small + lo.lainv, tid);
  if (tid < MP) {
    beta[tid] = small[lo.beta + tid];
    bbar[tid] = small[lo.bbar + tid];
    vyb[tid] = small[lo.vyb + tid];
  }
  for (int c = D + 1; c < PC; ++c) XsT[c * LDT + tid] = 0.0;
  if (tid < 4 * D) gbs[tid] = 0.0;
  double sacc[MF][MF][2], pacc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
#pragma unroll
    for (int j = 0; j < MF; ++j) sacc[i][j][0] = sacc[i][j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NF; ++j) pacc[i][j][0] = pacc[i][j][1] = 0.0;
  }
  double sum_lb = 0.0;
  const double ea = par[0];
  const double* lamg = rowv;
  const double* lb0g = rowv + N;
  const double* rbg = rowv + 2 * N;
  const double* tbg = rowv + 3 * N;
  const double* alg = rowv + 4 * N;
  __syncthreads();
  for (int64_t base = row_lo + (int64_t)blockIdx.x * RB; base < row_hi; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < row_hi;
    for (int d = 0; d < D; ++d) XsT[d * LDT + tid] = live ? X[i * D + d] * par[2 + d] : 0.0;
    XsT[D * LDT + tid] = live ? 1.0 : 0.0;
#pragma unroll 8
    for (int m = 0; m < MP; ++m) {
      if (OBJ != 2) VT[m * LDT + tid] = live ? VgT[(int64_t)m * N + i] : 0.0;   // OBJ = 2 borrows VT first
      WT[m * LDT + tid] = live ? WgT[(int64_t)m * N + i] : 0.0;
    }
    __syncwarp();
    const double lam = live ? lamg[i] : 1.0, yi = live ? y[i] : 0.0;
    const double il = 1.0 / lam;
    double lb0, bw = 0.0;
    if (OBJ == 2) {
      // kc seeds: D_i = 2 cbar H^-1 W - (2/lam) H_bar W - mbar h + alpha g_bar,  abar = -mbar lam + W'g_bar
      warp_tile_mm<MP>(WT, MatH, CT, warp, lane);               // H_bar_f W
      const double al = live ? alg[i] : 0.0;
      const double cb = live ? rbg[i] : 0.0, mb = live ? tbg[i] : 0.0;   // cbar_i, mbar_i (pass 2b)
      double q1 = 0.0, wg = 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) {
        const double w = WT[m * LDT + tid];
        q1 = fma(w, CT[m * LDT + tid], q1);
        wg = fma(w, gbv[m], wg);
        bw = fma(bbar[m], w, bw);
        VT[m * LDT + tid] = -2.0 * il * CT[m * LDT + tid];
      }
      __syncwarp();
      warp_tile_mm<MP>(WT, MatHi, CT, warp, lane);              // H_f^-1 W
      const double abar = -mb * lam + wg;
      const double tbar = -abar * il;
      lb0 = live ? (-mb * al + cb + q1 * il * il - abar * al * il) : 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m)
        WT[m * LDT + tid] = live ? (tbar * beta[m] + VT[m * LDT + tid] + 2.0 * cb * CT[m * LDT + tid] - mb * hv[m] +
                                    al * gbv[m]) : 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) VT[m * LDT + tid] = live ? VgT[(int64_t)m * N + i] : 0.0;
      __syncwarp();
    } else if (OBJ) {
      warp_tile_mm<MP>(WT, MatH, CT, warp, lane);               // Hhat_f W
      const double al = live ? alg[i] : 0.0;
      double q1 = 0.0, wh = 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) {
        const double w = WT[m * LDT + tid];
        q1 = fma(w, CT[m * LDT + tid], q1);
        wh = fma(w, hv[m], wh);
        bw = fma(bbar[m], w, bw);
      }
      const double abar = lam * al + wh;                        // dL/dalpha_i
      const double tbar = -abar * il;
      lb0 = live ? (0.5 * il + 0.5 * al * al + q1 * il * il - abar * al * il) : 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m)                               // W_bar = t_bar beta + D_i
        WT[m * LDT + tid] = live ? fma(tbar, beta[m], fma(-2.0 * il, CT[m * LDT + tid], al * hv[m])) : 0.0;
      __syncwarp();
    }
    warp_tile_mm<MP>(VT, MatCb, CT, warp, lane);                // CV = V C_bar
    double s1 = 0.0;
    if (OBJ) {
#pragma unroll 8
      for (int m = 0; m < MP; ++m) s1 = fma(VT[m * LDT + tid], CT[m * LDT + tid], s1);
    } else {
      const double rbar = live ? rbg[i] : 0.0, tbar = live ? tbg[i] : 0.0;
      lb0 = live ? lb0g[i] : 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) {
        const double w = WT[m * LDT + tid];
        s1 = fma(VT[m * LDT + tid], CT[m * LDT + tid], s1);
        bw = fma(bbar[m], w, bw);
        WT[m * LDT + tid] = fma(tbar, beta[m], 2.0 * rbar * w);   // W_bar
      }
    }
    const double lb = live ? (lb0 - bw * yi * il * il - s1 * il * il) : 0.0;
    sum_lb += lb;
    __syncwarp();
    warp_tile_mm<MP>(WT, MatLC, WT, warp, lane);                // L_C^-T W_bar
#pragma unroll 8
    for (int m = 0; m < MP; ++m)
      WT[m * LDT + tid] += vyb[m] * yi * il + 2.0 * CT[m * LDT + tid] * il - 2.0 * lb * VT[m * LDT + tid];   // V_bar
    __syncwarp();
    tile_outer<MF, MF, false>(WT, VT, nullptr, warp, lane, sacc);   // S += V_bar' V
    warp_tile_mm<MP>(WT, MatLA, CT, warp, lane);                // Kuf_bar = V_bar L_A^-1
#pragma unroll 4
    for (int m = 0; m < MP; ++m) {                              // G = Kuf_bar o Kuf (independent exps overlap)
      double r2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const double df = Us[m * D + d] - XsT[d * LDT + tid];
        r2 = fma(df, df, r2);
      }
      CT[m * LDT + tid] = (m < M && live) ? CT[m * LDT + tid] * ea * exp(-0.5 * r2) : 0.0;
    }
    auto gb_row = [&](int d) {                                  // sum_m G_im ((u_md - x_id)/l_d)^2 for this row
      const double xd = XsT[d * LDT + tid];
      double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
      for (int m = 0; m < MP; m += 4) {
        const double d0 = Us[m * D + d] - xd, d1 = Us[(m + 1) * D + d] - xd;
        const double d2 = Us[(m + 2) * D + d] - xd, d3 = Us[(m + 3) * D + d] - xd;
        g0 = fma(CT[m * LDT + tid], d0 * d0, g0);
        g1 = fma(CT[(m + 1) * LDT + tid], d1 * d1, g1);
        g2 = fma(CT[(m + 2) * LDT + tid], d2 * d2, g2);
        g3 = fma(CT[(m + 3) * LDT + tid], d3 * d3, g3);
      }
      return (g0 + g1) + (g2 + g3);
    };
    for (int d = 0; d < D; ++d) {
      const double v = warp_sum(gb_row(d));
      if (lane == 0) gbs[d * 4 + warp] += v;
    }
    __syncwarp();
    tile_outer<MF, NF, false>(CT, XsT, nullptr, warp, lane, pacc);  // P += G' [xs | 1]
    __syncwarp();
  }
  __syncthreads();                                              // every warp is done with its CT rows
  for (int e = tid; e < MP * MP + MP * PC; e += RB) Cs[e] = 0.0;
  __syncthreads();
  frags_to_smem<MF, MF>(sacc, Cs, MP, warp, lane);
  frags_to_smem<MF, NF>(pacc, Cs + MP * MP, PC, warp, lane);
  const int len = MP * MP + MP * PC + D + 1;
  double* out = part + ((int64_t)fold * gridDim.x + blockIdx.x) * len;
  for (int e = tid; e < MP * MP + MP * PC; e += RB) out[e] = Cs[e];
  if (tid < D) out[MP * MP + MP * PC + tid] = (gbs[tid * 4] + gbs[tid * 4 + 1]) + (gbs[tid * 4 + 2] + gbs[tid * 4 + 3]);
  const double sl = block_sum(sum_lb, red);
  if (tid == 0) out[MP * MP + MP * PC + D] = sl;
}

// ---- 4-fold block objectives (DSS, K20:538-582): pass 2 over fold-aligned tiles ---------------------
// blockIdx.y = fold.  W = V L_C^-T and alpha as in pass 2; accumulates, for the fold,
//   P_f = sum W_i W_i' / lam_i (= I - H_f),  g_f = sum W_i alpha_i,  sum log lam_i,  sum lam_i alpha_i^2
// part[(fold * gridDim.x + blockIdx.x)] = [P_f (MP*MP) | g_f (MP) | s1 | s2]
template <int MP>
__global__ void __launch_bounds__(RB)
fitc_row2_block_kernel(const double* __restrict__ y, int64_t N, int D, const double* __restrict__ small,
                       const double* __restrict__ VgT, double* __restrict__ WgT, double* __restrict__ rowv,
                       double* __restrict__ part, FoldGeom fg) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8, LDM = TileCfg<MP>::LDM;
  const SmallLayout lo(MP, D);
  double* MatC = sh;                      // [MP][LDM]  B[k][n] = L_C^-1[n][k]
  double* beta = MatC + MP * LDM;         // [MP]
  double* WT = beta + MP;                 // [MP][LDT]
  double* scs = WT + (size_t)MP * LDT;    // [RB] 1 / lambda
  double* als = scs + RB;                 // [RB] alpha
  double* Cs = als + RB;                  // [MP][MP] + [MP]
  double* red = Cs + MP * MP + MP;        // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fold = blockIdx.y;
  const int64_t row_lo = fg.lo[fold], row_hi = fg.hi[fold];
  stage_mat<MP, true>(MatC, small + lo.lcinv, tid);
  if (tid < MP) beta[tid] = small[lo.beta + tid];
  for (int e = tid; e < MP * MP + MP; e += RB) Cs[e] = 0.0;
  double pacc[MF][MF][2], gacc[MF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    gacc[i][0] = gacc[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < MF; ++j) pacc[i][j][0] = pacc[i][j][1] = 0.0;
  }
  double s1 = 0.0, s2 = 0.0;
  __syncthreads();
  const double* lamg = rowv;
  double* alg = rowv + 4 * N;
  for (int64_t base = row_lo + (int64_t)blockIdx.x * RB; base < row_hi; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < row_hi;
#pragma unroll 8
    for (int m = 0; m < MP; ++m) WT[m * LDT + tid] = live ? VgT[(int64_t)m * N + i] : 0.0;
    __syncwarp();
    warp_tile_mm<MP>(WT, MatC, WT, warp, lane);                 // W = V L_C^-T
    double sc = 0.0, al = 0.0;
    if (live) {
      double wb = 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) {
        const double w = WT[m * LDT + tid];
        wb = fma(w, beta[m], wb);
        WgT[(int64_t)m * N + i] = w;
      }
      const double lam = lamg[i];
      sc = 1.0 / lam;
      al = (y[i] - wb) * sc;
      alg[i] = al;
      s1 += log(lam);
      s2 = fma(lam * al, al, s2);
    }
    scs[tid] = sc;
    als[tid] = al;
    __syncwarp();
    tile_outer<MF, MF, true>(WT, WT, scs, warp, lane, pacc);
    tile_col<MF>(WT, als, warp, lane, gacc);
    __syncwarp();
  }
  frags_to_smem<MF, MF>(pacc, Cs, MP, warp, lane);
  colfrag_to_smem<MF>(gacc, Cs + MP * MP, warp, lane);
  const double t1 = block_sum(s1, red);
  const double t2 = block_sum(s2, red);
  const int len = MP * MP + MP + 2;
  double* out = part + ((int64_t)fold * gridDim.x + blockIdx.x) * len;
  for (int e = tid; e < MP * MP + MP; e += RB) out[e] = Cs[e];
  if (tid == 0) {
    out[MP * MP + MP] = t1;
    out[MP * MP + MP + 1] = t2;
  }
}

// acc[f][e] = sum_b part[(f * nblocks + b)][e]   (grid.y = fold)
__global__ void __launch_bounds__(256)
fitc_reduce_folds_kernel(const double* __restrict__ part, int nblocks, int len, double* __restrict__ acc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  const int f = blockIdx.y;
  double s = 0.0;
#pragma unroll 16
  for (int b = 0; b < nblocks; ++b) s += part[((int64_t)f * nblocks + b) * len + e];
  acc[(int64_t)f * len + e] = s;
}

// Replicated fold algebra of the DSS objective, one warp per fold:
//   H_f = I - P_f = L_H L_H',  h_f = H_f^-1 g_f,  Hhat_f = -1/2 H_f^-1 - 1/2 h_f h_f'
//   obj  = sum_f [ n_f/2 log 2pi + 1/2 s1_f - log|L_H| + 1/2 s2_f + 1/2 g_f' h_f ]
//   beta_bar = -sum_f h_f,   G_W = sum_f ( -2 Hhat_f P_f + h_f g_f' )
// and hands [G_W / 2 | beta_bar | obj] to the C_bar kernel in the slot layout of pass 2
// (S_W = beta beta_bar' + 2 (G_W / 2) + beta_bar beta').
template <int MP>
__global__ void __launch_bounds__(128)
fitc_block_small_kernel(const double* __restrict__ accf, double* __restrict__ small, double* __restrict__ acc2,
                        int D, double fold_rows, int kc_phase_a, int* __restrict__ info) {
  extern __shared__ double sh[];
  const SmallLayout lo(MP, D);
  const int tid = threadIdx.x, lane = tid & 31, f = tid >> 5;
  const int len = MP * MP + MP + 2;
  double* H = sh + (size_t)f * (3 * MP * MP + 4 * MP);   // per warp: H/L_H, T = L_H^-1, Hh = Hhat, vectors
  double* Tm = H + MP * MP;
  double* Hh = Tm + MP * MP;
  double* Li = Hh + MP * MP;
  double* hv = Li + MP;
  double* gv = hv + MP;
  double* contrib = sh + (size_t)4 * (3 * MP * MP + 4 * MP);   // [4][MP*MP + MP + 1]
  const double* P = accf + (size_t)f * len;
  for (int e = lane; e < MP * MP; e += 32) {
    const int r = e / MP, c = e - r * MP;
    H[e] = (r == c ? 1.0 : 0.0) - P[e];
  }
  if (lane < MP) gv[lane] = P[MP * MP + lane];
  __syncwarp();
  const int bad = warp_chol<MP>(H, lane);
  if (bad && lane == 0) atomicCAS(info, 0, 2000000 + bad);
  if (lane < MP) Li[lane] = 1.0 / H[lane * MP + lane];
  __syncwarp();
  double logdet = (lane < MP) ? -log(Li[lane]) : 0.0;
  logdet = warp_sum(logdet);
  double hx = warp_fwd_vec<MP>(H, Li, lane < MP ? gv[lane] : 0.0, lane);
  hx = warp_bwd_vecT<MP>(H, Li, hx, lane);                        // h_f = H^-1 g_f
  if (lane < MP) hv[lane] = hx;
  warp_tri_inverse<MP>(H, Li, Tm, lane);                          // L_H^-1
  __syncwarp();
  double gh = (lane < MP) ? gv[lane] * hx : 0.0;
  gh = warp_sum(gh);
  if (lane < MP) {
    const volatile double* Tv = Tm;
    for (int r = 0; r < MP; ++r) {                                // column `lane` of H^-1 = T' T
      double sacc = 0.0;
      const int k0 = r > lane ? r : lane;
      for (int k = k0; k < MP; ++k) sacc = fma(Tv[k * MP + r], Tv[k * MP + lane], sacc);
      const double hh = kc_phase_a ? sacc : -0.5 * sacc - 0.5 * hv[r] * hx;   // kc: H^-1 itself
      Hh[r * MP + lane] = hh;
      small[lo.fh + f * MP * MP + r * MP + lane] = hh;
    }
    small[lo.fhv + f * MP + lane] = hx;
  }
  __syncwarp();
  if (kc_phase_a) return;                                         // uniform per launch
  double* cf = contrib + (size_t)f * (MP * MP + MP + 1);
  if (lane < MP) {
    for (int r = 0; r < MP; ++r) {                                // column `lane` of -2 Hhat P + h g'
      double sacc = 0.0;
      for (int k = 0; k < MP; ++k) sacc = fma(Hh[r * MP + k], P[k * MP + lane], sacc);
      cf[r * MP + lane] = -2.0 * sacc + hv[r] * gv[lane];
    }
    cf[MP * MP + lane] = -hx;
  }
  if (lane == 0)
    cf[MP * MP + MP] = fold_rows * 0.91893853320467274178 + 0.5 * P[MP * MP + MP] - logdet + 0.5 * P[MP * MP + MP + 1] + 0.5 * gh;
  __syncthreads();
  for (int e = tid; e < MP * MP + MP + 1; e += 128) {
    const double tot = ((contrib[e] + contrib[(MP * MP + MP + 1) + e]) + contrib[2 * (MP * MP + MP + 1) + e]) +
                       contrib[3 * (MP * MP + MP + 1) + e];
    acc2[e] = (e < MP * MP) ? 0.5 * tot : tot;
  }
}

// ---- block CRPS "kc" (K20:669-714): pass 2b over fold-aligned tiles ------------------------------------
// Fold predictive of row i: m_i = y_i - lam_i alpha_i - W_i' h_f, c_i = lam_i + W_i' H_f^-1 W_i.
// Scores it with the closed-form CRPS, stores the seeds mbar_i, cbar_i and accumulates per fold
//   E_f = sum cbar_i W_i W_i',  hbar_f = -sum mbar_i W_i,  obj_f = (1/n_f) sum crps_i.
// part[(fold * gridDim.x + blockIdx.x)] = [E_f (MP*MP) | hbar_f (MP) | obj_f]
template <int MP>
__global__ void __launch_bounds__(RB)
fitc_row2b_kc_kernel(const double* __restrict__ y, int64_t N, int D, const double* __restrict__ small,
                     const double* __restrict__ WgT, double* __restrict__ rowv, double* __restrict__ part,
                     FoldGeom fg, double inv_fold_rows) {
  extern __shared__ double sh[];
  constexpr int MF = MP / 8, LDM = TileCfg<MP>::LDM;
  const SmallLayout lo(MP, D);
  double* MatHi = sh;                     // [MP][LDM]  H_f^-1 (symmetric)
  double* hv = MatHi + MP * LDM;          // [MP]
  double* WT = hv + MP;                   // [MP][LDT]
  double* CT = WT + (size_t)MP * LDT;     // [MP][LDT]  W H_f^-1
  double* cbs = CT + (size_t)MP * LDT;    // [RB] cbar
  double* mbs = cbs + RB;                 // [RB] -mbar
  double* Cs = mbs + RB;                  // [MP][MP] + [MP]
  double* red = Cs + MP * MP + MP;        // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fold = blockIdx.y;
  const int64_t row_lo = fg.lo[fold], row_hi = fg.hi[fold];
  stage_mat<MP, false>(MatHi, small + lo.fh + fold * MP * MP, tid);
  if (tid < MP) hv[tid] = small[lo.fhv + fold * MP + tid];
  for (int e = tid; e < MP * MP + MP; e += RB) Cs[e] = 0.0;
  double eacc[MF][MF][2], hacc[MF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    hacc[i][0] = hacc[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < MF; ++j) eacc[i][j][0] = eacc[i][j][1] = 0.0;
  }
  double obj = 0.0;
  __syncthreads();
  const double* lamg = rowv;
  double* cbg = rowv + 2 * N;   // cbar_i  (slot of r_bar)
  double* mbg = rowv + 3 * N;   // mbar_i  (slot of t_bar)
  const double* alg = rowv + 4 * N;
  for (int64_t base = row_lo + (int64_t)blockIdx.x * RB; base < row_hi; base += (int64_t)gridDim.x * RB) {
    const int64_t i = base + tid;
    const bool live = i < row_hi;
#pragma unroll 8
    for (int m = 0; m < MP; ++m) WT[m * LDT + tid] = live ? WgT[(int64_t)m * N + i] : 0.0;
    __syncwarp();
    warp_tile_mm<MP>(WT, MatHi, CT, warp, lane);                // H_f^-1 W_i (rows)
    double cb = 0.0, mb = 0.0;
    if (live) {
      double q = 0.0, wh = 0.0;
#pragma unroll 8
      for (int m = 0; m < MP; ++m) {
        const double w = WT[m * LDT + tid];
        q = fma(w, CT[m * LDT + tid], q);
        wh = fma(w, hv[m], wh);
      }
      const double lam = lamg[i], al = alg[i], yi = y[i];
      const double cv = lam + q;                                // fold predictive variance (K20:704)
      const double mu = yi - lam * al - wh;                     // fold predictive mean     (K20:698)
      const double sd = sqrt(cv), z = (yi - mu) / sd;
      const double tpm1 = erf(z * INV_SQRT2);
      const double phi2 = 2.0 * INV_SQRT_2PI * exp(-0.5 * z * z);
      obj += sd * (z * tpm1 + phi2 - INV_SQRT_PI) * inv_fold_rows;
      mb = -tpm1 * inv_fold_rows;
      cb = (phi2 - INV_SQRT_PI) / (2.0 * sd) * inv_fold_rows;
      cbg[i] = cb;
      mbg[i] = mb;
    }
    cbs[tid] = cb;
    mbs[tid] = -mb;
    __syncwarp();
    tile_outer<MF, MF, true>(WT, WT, cbs, warp, lane, eacc);    // E_f
    tile_col<MF>(WT, mbs, warp, lane, hacc);                    // hbar_f = -sum mbar W
    __syncwarp();
  }
  frags_to_smem<MF, MF>(eacc, Cs, MP, warp, lane);
  colfrag_to_smem<MF>(hacc, Cs + MP * MP, warp, lane);
  const double o = block_sum(obj, red);
  const int len = MP * MP + MP + 1;
  double* out = part + ((int64_t)fold * gridDim.x + blockIdx.x) * len;
  for (int e = tid; e < MP * MP + MP; e += RB) out[e] = Cs[e];
  if (tid == 0) out[MP * MP + MP] = o;
}

// Replicated fold algebra of kc, phase b (one warp per fold):
//   g_bar = H^-1 h_bar,  H_bar = -H^-1 E H^-1 - sym(g_bar h'),
//   beta_bar = sum_f (-h_bar - P g_bar),  G_W = sum_f ( h h_bar' + 2 H^-1 E + g_bar g' - 2 H_bar P ),  obj = sum_f obj_f
template <int MP>
__global__ void __launch_bounds__(128)
fitc_block_small_kc_kernel(const double* __restrict__ accf, const double* __restrict__ accf2,
                           double* __restrict__ small, double* __restrict__ acc2, int D) {
  extern __shared__ double sh[];
  const SmallLayout lo(MP, D);
  const int tid = threadIdx.x, lane = tid & 31, f = tid >> 5;
  const int len1 = MP * MP + MP + 2, len2 = MP * MP + MP + 1;
  double* Hi = sh + (size_t)f * (3 * MP * MP + 4 * MP);   // H^-1
  double* M1 = Hi + MP * MP;                              // H^-1 E
  double* Hb = M1 + MP * MP;                              // H_bar
  double* hv = Hb + MP * MP;
  double* gb = hv + MP;                                   // g_bar
  double* hb = gb + MP;                                   // h_bar
  double* gv = hb + MP;                                   // g
  double* contrib = sh + (size_t)4 * (3 * MP * MP + 4 * MP);
  const double* P = accf + (size_t)f * len1;
  const double* E = accf2 + (size_t)f * len2;
  for (int e = lane; e < MP * MP; e += 32) Hi[e] = small[lo.fh + f * MP * MP + e];
  if (lane < MP) {
    hv[lane] = small[lo.fhv + f * MP + lane];
    hb[lane] = E[MP * MP + lane];
    gv[lane] = P[MP * MP + lane];
  }
  __syncwarp();
  if (lane < MP) {
    double sg = 0.0;
    for (int k = 0; k < MP; ++k) sg = fma(Hi[lane * MP + k], hb[k], sg);
    gb[lane] = sg;
    small[lo.fgv + f * MP + lane] = sg;
    for (int r = 0; r < MP; ++r) {                          // column `lane` of H^-1 E
      double sacc = 0.0;
      for (int k = 0; k < MP; ++k) sacc = fma(Hi[r * MP + k], E[k * MP + lane], sacc);
      M1[r * MP + lane] = sacc;
    }
  }
  __syncwarp();
  if (lane < MP) {
    for (int r = 0; r < MP; ++r) {                          // column `lane` of H_bar
      double sacc = 0.0;
      for (int k = 0; k < MP; ++k) sacc = fma(M1[r * MP + k], Hi[k * MP + lane], sacc);
      const double hbv = -sacc - 0.5 * (gb[r] * hv[lane] + hv[r] * gb[lane]);
      Hb[r * MP + lane] = hbv;
      small[lo.fh2 + f * MP * MP + r * MP + lane] = hbv;
    }
  }
  __syncwarp();
  double* cf = contrib + (size_t)f * (MP * MP + MP + 1);
  if (lane < MP) {
    for (int r = 0; r < MP; ++r) {
      double sacc = 0.0;
      for (int k = 0; k < MP; ++k) sacc = fma(Hb[r * MP + k], P[k * MP + lane], sacc);
      cf[r * MP + lane] = hv[r] * hb[lane] + 2.0 * M1[r * MP + lane] + gb[r] * gv[lane] - 2.0 * sacc;
    }
    double pg = 0.0;
    for (int k = 0; k < MP; ++k) pg = fma(P[lane * MP + k], gb[k], pg);
    cf[MP * MP + lane] = -hb[lane] - pg;
  }
  if (lane == 0) cf[MP * MP + MP] = E[MP * MP + MP];
  __syncthreads();
  for (int e = tid; e < MP * MP + MP + 1; e += 128) {
    const double tot = ((contrib[e] + contrib[(MP * MP + MP + 1) + e]) + contrib[2 * (MP * MP + MP + 1) + e]) +
                       contrib[3 * (MP * MP + MP + 1) + e];
    acc2[e] = (e < MP * MP) ? 0.5 * tot : tot;
  }
}

// ---- LOO outputs and prediction ----------------------------------------------------------------------
__global__ void fitc_loo_kernel(const double* __restrict__ y, const double* __restrict__ rowv, int64_t N,
                                double* __restrict__ mean, double* __restrict__ var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double alpha = rowv[4 * N + i], d = rowv[5 * N + i];
  mean[i] = y[i] - alpha / d;   // K20:231
  var[i] = 1.0 / d;             // K20:232
}

template <int MP>
__global__ void __launch_bounds__(RB)
fitc_predict_kernel(const double* __restrict__ Xs, int64_t T, int D, int M, const double* __restrict__ par,
                    const double* __restrict__ small, double* __restrict__ mean, double* __restrict__ var) {
  extern __shared__ double sh[];
  const SmallLayout lo(MP, D);
  double* Us = sh;
  double* LA = Us + MP * D;
  double* LAi = LA + MP * MP;
  double* LC = LAi + MP;
  double* LCi = LC + MP * MP;
  double* beta = LCi + MP;
  double* XsT = beta + MP;   // [D][LDT]
  const int tid = threadIdx.x;
  for (int e = tid; e < MP * D; e += RB) Us[e] = small[lo.us + e];
  for (int e = tid; e < MP * MP; e += RB) {
    LA[e] = small[lo.la + e];
    LC[e] = small[lo.lc + e];
  }
  if (tid < MP) {
    LAi[tid] = small[lo.lai + tid];
    LCi[tid] = small[lo.lci + tid];
    beta[tid] = small[lo.beta + tid];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * RB + tid;
  const bool live = i < T;
  for (int d = 0; d < D; ++d) XsT[d * LDT + tid] = live ? Xs[i * D + d] * par[2 + d] : 0.0;
  double v[MP];
  kernel_vec<MP>(Us, XsT, tid, M, D, par[0], v);
  fwd_subst<MP>(LA, LAi, v);
  double q = 0.0;
#pragma unroll
  for (int m = 0; m < MP; ++m) q = fma(v[m], v[m], q);
  fwd_subst<MP>(LC, LCi, v);
  double r = 0.0, mb = 0.0;
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    r = fma(v[m], v[m], r);
    mb = fma(v[m], beta[m], mb);
  }
  if (live) {
    mean[i] = mb;
    var[i] = par[1] + par[0] - q + r;
  }
}

// ---- host-side dispatch -------------------------------------------------------------------------------
size_t smem_row1(int MP, int D) { return ((size_t)MP * D + MP * MP + MP + (size_t)D * LDT + (size_t)MP * LDT + RB + MP * MP + MP) * 8; }
size_t smem_row2(int MP) { return ((size_t)MP * MP + 2 * MP + (size_t)MP * LDT + 2 * RB + MP * MP + MP + 32) * 8; }
size_t smem_row3(int MP, int D, int PC) {
  return ((size_t)MP * D + 3 * MP * MP + 5 * MP + (size_t)PC * LDT + 3 * (size_t)MP * LDT + (size_t)D * RB + MP * MP +
          MP * PC + 32) * 8;
}
size_t smem_pred(int MP, int D) { return ((size_t)MP * D + 2 * MP * MP + 3 * MP + (size_t)D * LDT) * 8; }

// raise the dynamic shared-memory limit of a kernel once per (kernel, size)
template <typename K>
int set_smem(gps_ctx* ctx, K kern, size_t bytes) {
  // one slot per (kernel type K, device): kernels sharing a signature share the slot, hence `last`
  static size_t configured[64] = {};
  static K last[64] = {};
  const int dv = ctx->device & 63;
  if (last[dv] == kern && configured[dv] >= bytes) return GPS_OK;
  GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  last[dv] = kern;
  configured[dv] = bytes;
  return GPS_OK;
}

template <int MP>
int run_row1(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row1(MP, ctx->D);
  GPS_CHECK(set_smem(ctx, fitc_row1_kernel<MP>, sm));
  fitc_row1_kernel<MP><<<f.grid, RB, sm, ctx->stream>>>(ctx->X.p, ctx->y.p, ctx->N, ctx->D, f.M, ctx->params.p,
                                                        f.small.p, f.V.p, f.rowv.p, part);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_row2(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row2(MP);
  GPS_CHECK(set_smem(ctx, fitc_row2_kernel<MP>, sm));
  fitc_row2_kernel<MP><<<f.grid, RB, sm, ctx->stream>>>(ctx->y.p, ctx->N, ctx->D, f.score, 1.0 / (double)f.world_n,
                                                        f.small.p, f.V.p, f.W.p, f.rowv.p, part);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP, int NF>
int run_row3(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row3(MP, ctx->D, 8 * NF);
  GPS_CHECK(set_smem(ctx, (fitc_row3_kernel<MP, NF>), sm));
  fitc_row3_kernel<MP, NF><<<f.grid, RB, sm, ctx->stream>>>(ctx->X.p, ctx->y.p, ctx->N, ctx->D, f.M, ctx->params.p,
                                                            f.small.p, f.V.p, f.W.p, f.rowv.p, part);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

size_t smem_row1t(int MP, int D) { return ((size_t)MP * D + MP * (MP + 4) + (size_t)D * LDT + (size_t)MP * LDT + RB + MP * MP + MP) * 8; }
size_t smem_row2t(int MP) { return ((size_t)MP * (MP + 4) + MP + (size_t)MP * LDT + 2 * RB + MP * MP + MP + 32) * 8; }
size_t smem_row3t(int MP, int D, int PC, int obj) {
  size_t n = (size_t)MP * D + 3 * MP * (MP + 4) + 3 * MP + (size_t)PC * LDT + 3 * (size_t)MP * LDT + 32;
  if (obj) n += (size_t)2 * MP * (MP + 4) + 2 * MP;   // fold matrices of the block objectives
  n += (size_t)4 * D;                                 // per-warp g_b accumulators
  return n * 8;
}
size_t smem_row2kc(int MP) { return ((size_t)MP * (MP + 4) + MP + 2 * (size_t)MP * LDT + 2 * RB + MP * MP + MP + 32) * 8; }
size_t smem_row2b(int MP) { return ((size_t)MP * (MP + 4) + MP + (size_t)MP * LDT + 2 * RB + MP * MP + MP + 32) * 8; }
size_t smem_blocksmall(int MP) { return ((size_t)4 * (3 * MP * MP + 4 * MP) + 4 * (MP * MP + MP + 1)) * 8; }

template <int MP>
int run_row1t(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row1t(MP, ctx->D);
  GPS_CHECK(set_smem(ctx, fitc_row1_tile_kernel<MP>, sm));
  fitc_row1_tile_kernel<MP><<<f.grid, RB, sm, ctx->stream>>>(ctx->X.p, ctx->y.p, ctx->N, ctx->D, f.M, ctx->params.p,
                                                             f.small.p, f.V.p, f.rowv.p, part);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_row2t(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row2t(MP);
  GPS_CHECK(set_smem(ctx, fitc_row2_tile_kernel<MP>, sm));
  fitc_row2_tile_kernel<MP><<<f.grid, RB, sm, ctx->stream>>>(ctx->y.p, ctx->N, ctx->D, f.score,
                                                             1.0 / (double)f.world_n, f.small.p, f.V.p, f.W.p,
                                                             f.rowv.p, part);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP, int NF>
int run_row3t(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row3t(MP, ctx->D, 8 * NF, 0);
  GPS_CHECK(set_smem(ctx, (fitc_row3_tile_kernel<MP, NF, 0>), sm));
  fitc_row3_tile_kernel<MP, NF, 0><<<f.grid, RB, sm, ctx->stream>>>(ctx->X.p, ctx->y.p, ctx->N, ctx->D, f.M, ctx->params.p,
                                                                    f.small.p, f.V.p, f.W.p, f.rowv.p, part, FoldGeom());
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP, int NF>
int run_row3b(gps_ctx* ctx, double* part, const FoldGeom& fg, int gx) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row3t(MP, ctx->D, 8 * NF, 1);
  GPS_CHECK(set_smem(ctx, (fitc_row3_tile_kernel<MP, NF, 1>), sm));
  fitc_row3_tile_kernel<MP, NF, 1><<<dim3(gx, 4), RB, sm, ctx->stream>>>(ctx->X.p, ctx->y.p, ctx->N, ctx->D, f.M,
                                                                         ctx->params.p, f.small.p, f.V.p, f.W.p,
                                                                         f.rowv.p, part, fg);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP, int NF>
int run_row3kc(gps_ctx* ctx, double* part, const FoldGeom& fg, int gx) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row3t(MP, ctx->D, 8 * NF, 1);
  GPS_CHECK(set_smem(ctx, (fitc_row3_tile_kernel<MP, NF, 2>), sm));
  fitc_row3_tile_kernel<MP, NF, 2><<<dim3(gx, 4), RB, sm, ctx->stream>>>(ctx->X.p, ctx->y.p, ctx->N, ctx->D, f.M,
                                                                         ctx->params.p, f.small.p, f.V.p, f.W.p,
                                                                         f.rowv.p, part, fg);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_row2kc(gps_ctx* ctx, double* part, const FoldGeom& fg, int gx, double inv_rows) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row2kc(MP);
  GPS_CHECK(set_smem(ctx, fitc_row2b_kc_kernel<MP>, sm));
  fitc_row2b_kc_kernel<MP><<<dim3(gx, 4), RB, sm, ctx->stream>>>(ctx->y.p, ctx->N, ctx->D, f.small.p, f.W.p, f.rowv.p,
                                                                 part, fg, inv_rows);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_blocksmall_kc(gps_ctx* ctx, const double* accf, const double* accf2, double* acc2) {
  auto& f = ctx->fitc;
  const size_t sm = smem_blocksmall(MP);
  GPS_CHECK(set_smem(ctx, fitc_block_small_kc_kernel<MP>, sm));
  fitc_block_small_kc_kernel<MP><<<1, 128, sm, ctx->stream>>>(accf, accf2, f.small.p, acc2, ctx->D);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_row2b(gps_ctx* ctx, double* part, const FoldGeom& fg, int gx) {
  auto& f = ctx->fitc;
  const size_t sm = smem_row2b(MP);
  GPS_CHECK(set_smem(ctx, fitc_row2_block_kernel<MP>, sm));
  fitc_row2_block_kernel<MP><<<dim3(gx, 4), RB, sm, ctx->stream>>>(ctx->y.p, ctx->N, ctx->D, f.small.p, f.V.p, f.W.p,
                                                                   f.rowv.p, part, fg);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_blocksmall(gps_ctx* ctx, const double* accf, double* acc2, double fold_rows, int kc_phase_a) {
  auto& f = ctx->fitc;
  const size_t sm = smem_blocksmall(MP);
  GPS_CHECK(set_smem(ctx, fitc_block_small_kernel<MP>, sm));
  fitc_block_small_kernel<MP><<<1, 128, sm, ctx->stream>>>(accf, f.small.p, acc2, ctx->D, fold_rows, kc_phase_a,
                                                            ctx->d_info);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <int MP>
int run_pred(gps_ctx* ctx, const double* Xs, int64_t T, double* mean, double* var) {
  auto& f = ctx->fitc;
  const size_t sm = smem_pred(MP, ctx->D);
  GPS_CHECK(set_smem(ctx, fitc_predict_kernel<MP>, sm));
  fitc_predict_kernel<MP><<<(unsigned)((T + RB - 1) / RB), RB, sm, ctx->stream>>>(Xs, T, ctx->D, f.M, ctx->params.p,
                                                                                f.small.p, mean, var);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

int pre_reduce(gps_ctx* ctx, const double** part, int* nblocks, int len);
int pc_of(int D);

template <int MP>
int run_small0(gps_ctx* ctx, const double* dU) {
  auto& f = ctx->fitc;
  fitc_small0_kernel<MP><<<1, 128, 0, ctx->stream>>>(dU, ctx->params.p, f.small.p, f.M, ctx->D, f.jitter, ctx->d_info);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}
template <int MP>
int run_small1(gps_ctx* ctx, const double* part, double* acc1) {
  auto& f = ctx->fitc;
  int nb = f.grid;
  if (part) GPS_CHECK(pre_reduce(ctx, &part, &nb, MP * MP + MP));
  fitc_small1_kernel<MP><<<1, 128, 0, ctx->stream>>>(part, nb, acc1, f.small.p, ctx->D, ctx->d_info);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}
template <int MP>
int run_small2(gps_ctx* ctx, const double* part, double* acc2) {
  auto& f = ctx->fitc;
  int nb = f.grid;
  if (part) GPS_CHECK(pre_reduce(ctx, &part, &nb, MP * MP + MP + 1));
  fitc_small2_kernel<MP><<<1, 128, 0, ctx->stream>>>(part, nb, acc2, f.small.p, f.M, ctx->D, f.score);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}
template <int MP>
int run_small3(gps_ctx* ctx, const double* part, int nblocks, const double* acc2, double* acc3, double* out);

#define MP_DISPATCH(MPV, CALL)                    \
  switch (MPV) {                                  \
    case 8:  { constexpr int MPC = 8;  CALL; } break;  \
    case 16: { constexpr int MPC = 16; CALL; } break;  \
    case 24: { constexpr int MPC = 24; CALL; } break;  \
    case 32: { constexpr int MPC = 32; CALL; } break;  \
    default: return gps_fail(ctx, GPS_EINVAL, "FITC: M padded to %d is not supported", MPV); \
  }

int pc_of(int D) { return (D + 1 <= 8) ? 8 : 16; }

// row passes: tile formulation (default) or the thread-per-row kernels (ctx->fitc_variant == 0)
int do_row1(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  if (f.tile) { MP_DISPATCH(f.MP, GPS_CHECK(run_row1t<MPC>(ctx, part))); }
  else { MP_DISPATCH(f.MP, GPS_CHECK(run_row1<MPC>(ctx, part))); }
  return GPS_OK;
}
int do_row2(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  if (f.tile) { MP_DISPATCH(f.MP, GPS_CHECK(run_row2t<MPC>(ctx, part))); }
  else { MP_DISPATCH(f.MP, GPS_CHECK(run_row2<MPC>(ctx, part))); }
  return GPS_OK;
}
int do_row3(gps_ctx* ctx, double* part) {
  auto& f = ctx->fitc;
  if (pc_of(ctx->D) == 8) {
    if (f.tile) { MP_DISPATCH(f.MP, GPS_CHECK((run_row3t<MPC, 1>(ctx, part)))); }
    else { MP_DISPATCH(f.MP, GPS_CHECK((run_row3<MPC, 1>(ctx, part)))); }
  } else {
    if (f.tile) { MP_DISPATCH(f.MP, GPS_CHECK((run_row3t<MPC, 2>(ctx, part)))); }
    else { MP_DISPATCH(f.MP, GPS_CHECK((run_row3<MPC, 2>(ctx, part)))); }
  }
  return GPS_OK;
}
int len1_of(int MP) { return MP * MP + MP; }
int len2_of(int MP) { return MP * MP + MP + 1; }
int len3_of(int MP, int D) { return MP * MP + MP * pc_of(D) + D + 1; }

template <int MP>
int run_small3(gps_ctx* ctx, const double* part, int nblocks, const double* acc2, double* acc3, double* out) {
  auto& f = ctx->fitc;
  int nb = nblocks;
  if (part) GPS_CHECK(pre_reduce(ctx, &part, &nb, MP * MP + MP * pc_of(ctx->D) + ctx->D + 1));
  fitc_small3_kernel<MP><<<1, 128, 0, ctx->stream>>>(part, nb, acc2, acc3, f.small.p, ctx->params.p, f.M, ctx->D,
                                                     pc_of(ctx->D), f.score, (double)f.world_n, out);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

// For large grids the partials are first summed in REDUCE_GROUPS slices by many CTAs; the
// replicated kernels (or the final reduce) then only add REDUCE_GROUPS rows.
constexpr int REDUCE_GROUPS = 16;
int pre_reduce(gps_ctx* ctx, const double** part, int* nblocks, int len) {
  auto& f = ctx->fitc;
  if (*nblocks <= 2 * REDUCE_GROUPS) return GPS_OK;
  const int per = (*nblocks + REDUCE_GROUPS - 1) / REDUCE_GROUPS;
  dim3 grid((len + 127) / 128, REDUCE_GROUPS);
  fitc_reduce_stage_kernel<<<grid, 128, 0, ctx->stream>>>(*part, *nblocks, len, per, f.part2.p);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  *part = f.part2.p;
  *nblocks = REDUCE_GROUPS;
  return GPS_OK;
}

int reduce_to(gps_ctx* ctx, const double* part, int nblocks, int len, double* acc) {
  GPS_CHECK(pre_reduce(ctx, &part, &nblocks, len));
  fitc_reduce_kernel<<<(len + 255) / 256, 256, 0, ctx->stream>>>(part, nblocks, len, acc);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

}  // namespace

extern "C" {

int gps_fitc_acc_len(int M, int D, int64_t* len1, int64_t* len2, int64_t* len3) {
  if (M > 32 && M <= 4096 && D > 0 && D <= 16) return gps_fitc_large_acc_len(M, D, len1, len2, len3);
  if (M <= 0 || M > 32 || D <= 0 || D > 15) return GPS_EINVAL;
  const int MP = (M + 7) / 8 * 8;
  if (len1) *len1 = len1_of(MP);
  if (len2) *len2 = len2_of(MP);
  if (len3) *len3 = len3_of(MP, D);
  return GPS_OK;
}

// `staged` = called through the public begin / pass1 / pass2 / pass3 / finish protocol, whose passes implement the
// row-additive objectives only: the 4-fold block objectives are rejected there instead of silently evaluating
// a hybrid (the fused gps_fitc_eval drives their fold-aligned passes itself).
static int fitc_begin_impl(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                           int64_t world_n, bool staged) {
  if (!ctx) return GPS_EINVAL;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "fitc: call gps_set_data first");
  if (!theta || !U || score < GPS_CRPS || score > GPS_KC) return gps_fail(ctx, GPS_EINVAL, "fitc: bad arguments");
  if (staged && (score == GPS_DSS || score == GPS_KC))
    return gps_fail(ctx, GPS_EINVAL, "fitc: the staged (caller-side all-reduce) protocol implements crps / logs / nlml only; "
                                     "dss and kc run through gps_fitc_eval on one GPU and gps_fitc_eval_sharded on several");
  if (M > 32) return gps_fitc_large_begin(ctx, theta, U, M, jitter, score, world_n);   // matrix form, same protocol
  const bool block_obj = score == GPS_DSS || score == GPS_KC;
  if (block_obj && ((world_n > 0 ? world_n : ctx->N) % 4 || world_n > ctx->N))
    return gps_fail(ctx, GPS_EINVAL, "fitc dss/kc: needs N %% 4 == 0 (K20:541-543); the row kernels run on one GPU (gps_fitc_eval_sharded shards them)");
  if (block_obj && ctx->fitc_variant == 0)
    return gps_fail(ctx, GPS_EINVAL, "fitc dss/kc: only the tile formulation of the row passes implements them");
  if (M <= 0 || M > 32)
    return gps_fail(ctx, GPS_EINVAL, "fitc: M=%d outside 1..4096", M);
  if (ctx->D > 15) return gps_fail(ctx, GPS_EINVAL, "fitc: D=%d > 15 not supported by the row kernels", ctx->D);
  GPS_CUDA(cudaSetDevice(ctx->device));
  auto& f = ctx->fitc;
  const int D = ctx->D;
  const int MP = (M + 7) / 8 * 8;
  const int64_t N = ctx->N;
  f.M = M; f.MP = MP; f.score = score; f.jitter = jitter; f.world_n = world_n > 0 ? world_n : N;
  f.begun = false; f.pass2_done = false; f.large = false; f.fused = false;
  f.tile = ctx->fitc_variant != 0;
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  if (!ctx->d_info) GPS_CUDA(cudaMalloc(&ctx->d_info, sizeof(int)));
  GPS_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream));
  const SmallLayout lo(MP, D);
  int64_t blocks = (N + RB - 1) / RB;
  const int64_t cap = (int64_t)ctx->sm_count * 4;
  f.grid = (int)(blocks < cap ? blocks : cap);
  GPS_CHECK(gps_ensure(ctx, f.V, (size_t)N * MP));
  GPS_CHECK(gps_ensure(ctx, f.W, (size_t)N * MP));
  GPS_CHECK(gps_ensure(ctx, f.rowv, (size_t)6 * N));
  GPS_CHECK(gps_ensure(ctx, f.small, (size_t)lo.total + M * D + 8 + D + 2 + M * D));
  {
    const size_t nblk = std::max<size_t>((size_t)f.grid, (size_t)4 * ctx->sm_count);
    GPS_CHECK(gps_ensure(ctx, f.part, nblk * len3_of(MP, D)));
  }
  GPS_CHECK(gps_ensure(ctx, f.accf, (size_t)8 * (MP * MP + MP + 2)));
  GPS_CHECK(gps_ensure(ctx, f.part2, (size_t)REDUCE_GROUPS * len3_of(MP, D)));
  GPS_CHECK(gps_ensure(ctx, f.acc1, len1_of(MP)));
  GPS_CHECK(gps_ensure(ctx, f.acc2, len2_of(MP)));
  GPS_CHECK(gps_ensure(ctx, f.acc3, len3_of(MP, D)));
  GPS_CHECK(gps_upload_params(ctx, theta, D, &f.ea, &f.sn2));
  double* dU = f.small.p + lo.total;   // raw inducing inputs
  GPS_CUDA(cudaMemcpyAsync(dU, U, (size_t)M * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  MP_DISPATCH(MP, GPS_CHECK(run_small0<MPC>(ctx, dU)));
  ctx->launches++;
  f.begun = true;
  return GPS_OK;
}

int gps_fitc_begin(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                   int64_t world_n) {
  return fitc_begin_impl(ctx, theta, U, M, jitter, score, world_n, true);
}

int gps_fitc_pass1(gps_ctx* ctx, double* acc1) {
  if (!ctx) return GPS_EINVAL;
  auto& f = ctx->fitc;
  if (!f.begun) return gps_fail(ctx, GPS_ESTATE, "fitc_pass1: call gps_fitc_begin first");
  GPS_CUDA(cudaSetDevice(ctx->device));
  if (f.large) {
    GPS_CHECK(gps_fitc_large_pass1(ctx, acc1));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPS_OK;
  }
  GPS_CHECK(do_row1(ctx, f.part.p));
  ctx->launches++;
  GPS_CHECK(reduce_to(ctx, f.part.p, f.grid, len1_of(f.MP), acc1));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));   // acc1 is handed to the caller's collective
  return GPS_OK;
}

int gps_fitc_pass2(gps_ctx* ctx, const double* acc1, double* acc2) {
  if (!ctx) return GPS_EINVAL;
  auto& f = ctx->fitc;
  if (!f.begun) return gps_fail(ctx, GPS_ESTATE, "fitc_pass2: call gps_fitc_begin first");
  GPS_CUDA(cudaSetDevice(ctx->device));
  if (f.large) {
    GPS_CHECK(gps_fitc_large_pass2(ctx, acc1, acc2, true));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPS_OK;
  }
  MP_DISPATCH(f.MP, GPS_CHECK(run_small1<MPC>(ctx, nullptr, const_cast<double*>(acc1))));
  GPS_CHECK(do_row2(ctx, f.part.p));
  ctx->launches += 2;
  GPS_CHECK(reduce_to(ctx, f.part.p, f.grid, len2_of(f.MP), acc2));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  f.pass2_done = true;
  return GPS_OK;
}

int gps_fitc_pass3(gps_ctx* ctx, const double* acc2, double* acc3) {
  if (!ctx) return GPS_EINVAL;
  auto& f = ctx->fitc;
  if (!f.pass2_done) return gps_fail(ctx, GPS_ESTATE, "fitc_pass3: pass 2 has not run");
  GPS_CUDA(cudaSetDevice(ctx->device));
  if (f.large) {
    GPS_CHECK(gps_fitc_large_pass3(ctx, acc2, acc3));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPS_OK;
  }
  MP_DISPATCH(f.MP, GPS_CHECK(run_small2<MPC>(ctx, nullptr, const_cast<double*>(acc2))));
  GPS_CHECK(do_row3(ctx, f.part.p));
  ctx->launches += 2;
  GPS_CHECK(reduce_to(ctx, f.part.p, f.grid, len3_of(f.MP, ctx->D), acc3));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

static int fitc_finish_impl(gps_ctx* ctx, const double* part, int nblocks, const double* acc2, double* acc3, double* obj,
                           double* grad_theta, double* grad_U) {
  auto& f = ctx->fitc;
  const int D = ctx->D, M = f.M, MP = f.MP;
  const SmallLayout lo(MP, D);
  double* out = f.small.p + lo.total + M * D;   // [obj | g_theta | g_U]
  const int nout = 1 + D + 2 + M * D;
  MP_DISPATCH(MP, GPS_CHECK(run_small3<MPC>(ctx, part, nblocks, acc2, acc3, out)));
  ctx->launches++;
  std::vector<double>& h = f.host_out;
  h.resize(nout + 1);
  GPS_CUDA(cudaMemcpyAsync(h.data(), out, nout * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  int info = 0;
  GPS_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info != 0)
    return gps_fail(ctx, GPS_ENOTPD, "fitc: %s not positive definite at pivot %d",
                    info >= 1000000 ? "I + V V'/lambda" : "K_uu + jitter I", info % 1000000);
  if (obj) *obj = h[0];
  if (grad_theta)
    for (int k = 0; k < D + 2; ++k) grad_theta[k] = h[1 + k];
  if (grad_U)
    for (int k = 0; k < M * D; ++k) grad_U[k] = h[1 + D + 2 + k];
  return GPS_OK;
}

int gps_fitc_finish(gps_ctx* ctx, const double* acc2, const double* acc3, double* obj, double* grad_theta,
                    double* grad_U) {
  if (!ctx) return GPS_EINVAL;
  auto& f = ctx->fitc;
  if (!f.pass2_done) return gps_fail(ctx, GPS_ESTATE, "fitc_finish: passes have not run");
  GPS_CUDA(cudaSetDevice(ctx->device));
  if (f.large) return gps_fitc_large_finish(ctx, acc2, acc3, obj, grad_theta, grad_U);
  return fitc_finish_impl(ctx, nullptr, 0, acc2, const_cast<double*>(acc3), obj, grad_theta, grad_U);
}

int gps_fitc_eval(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                  double* obj, double* grad_theta, double* grad_U) {
  if (!ctx) return GPS_EINVAL;
  if (M >= ctx->fitc_large_min_m) {
    if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "fitc: call gps_set_data first");
    if (!theta || !U) return gps_fail(ctx, GPS_EINVAL, "fitc: bad arguments");
    return gps_fitc_large_eval(ctx, theta, U, M, jitter, score, obj, grad_theta, grad_U);
  }
  if (ctx->fitc_variant == 2 && ctx->N > 0 && theta && U && gps_fitc_fused_supports(ctx, M, score))
    return gps_fitc_fused_eval(ctx, theta, U, M, jitter, score, ctx->N, nullptr, obj, grad_theta, grad_U);
  GPS_CHECK(fitc_begin_impl(ctx, theta, U, M, jitter, score, ctx->N, false));
  auto& f = ctx->fitc;
  f.fused = false;
  // single GPU: the same row kernels; the per-block partials are summed inside the replicated
  // kernels (no separate reduce launches, no host synchronisation between the passes):
  // 7 launches per evaluation.
  GPS_CHECK(do_row1(ctx, f.part.p));
  MP_DISPATCH(f.MP, GPS_CHECK(run_small1<MPC>(ctx, f.part.p, f.acc1.p)));
  if (score == GPS_DSS || score == GPS_KC) {
    // 4-fold block objective: fold-aligned tiles (blockIdx.y = fold), per-fold M x M algebra, then the
    // same C_bar / pass 3 / finish chain with seeds formed from the fold quantities
    const int64_t N = ctx->N, nfr = N / 4;
    FoldGeom fg;
    for (int k = 0; k < 4; ++k) { fg.lo[k] = k * nfr; fg.hi[k] = (k + 1) * nfr; }
    int gx = (int)((nfr + RB - 1) / RB);
    if (gx > ctx->sm_count) gx = ctx->sm_count;
    const int len2b = f.MP * f.MP + f.MP + 2;
    MP_DISPATCH(f.MP, GPS_CHECK(run_row2b<MPC>(ctx, f.part.p, fg, gx)));
    fitc_reduce_folds_kernel<<<dim3((len2b + 255) / 256, 4), 256, 0, ctx->stream>>>(f.part.p, gx, len2b, f.accf.p);
    GPS_LAUNCH_CHECK();
    MP_DISPATCH(f.MP, GPS_CHECK(run_blocksmall<MPC>(ctx, f.accf.p, f.acc2.p, (double)nfr, score == GPS_KC ? 1 : 0)));
    if (score == GPS_KC) {
      // block CRPS: a second fold-aligned pass scores every row and accumulates E_f, hbar_f
      const int len2c = f.MP * f.MP + f.MP + 1;
      double* accf2 = f.accf.p + 4 * len2b;
      MP_DISPATCH(f.MP, GPS_CHECK(run_row2kc<MPC>(ctx, f.part.p, fg, gx, 1.0 / (double)nfr)));
      fitc_reduce_folds_kernel<<<dim3((len2c + 255) / 256, 4), 256, 0, ctx->stream>>>(f.part.p, gx, len2c, accf2);
      GPS_LAUNCH_CHECK();
      MP_DISPATCH(f.MP, GPS_CHECK(run_blocksmall_kc<MPC>(ctx, f.accf.p, accf2, f.acc2.p)));
      ctx->launches += 3;
    }
    f.pass2_done = true;
    f.loo_ok = false;
    ctx->launches += 5;
    if (!(grad_theta || grad_U)) {
      GPS_CUDA(cudaMemsetAsync(f.acc3.p, 0, len3_of(f.MP, ctx->D) * sizeof(double), ctx->stream));
      return fitc_finish_impl(ctx, nullptr, 0, f.acc2.p, f.acc3.p, obj, nullptr, nullptr);
    }
    MP_DISPATCH(f.MP, GPS_CHECK(run_small2<MPC>(ctx, nullptr, f.acc2.p)));
    if (score == GPS_KC) {
      if (pc_of(ctx->D) == 8) {
        MP_DISPATCH(f.MP, GPS_CHECK((run_row3kc<MPC, 1>(ctx, f.part.p, fg, gx))));
      } else {
        MP_DISPATCH(f.MP, GPS_CHECK((run_row3kc<MPC, 2>(ctx, f.part.p, fg, gx))));
      }
    } else if (pc_of(ctx->D) == 8) {
      MP_DISPATCH(f.MP, GPS_CHECK((run_row3b<MPC, 1>(ctx, f.part.p, fg, gx))));
    } else {
      MP_DISPATCH(f.MP, GPS_CHECK((run_row3b<MPC, 2>(ctx, f.part.p, fg, gx))));
    }
    ctx->launches += 2;
    return fitc_finish_impl(ctx, f.part.p, 4 * gx, f.acc2.p, f.acc3.p, obj, grad_theta, grad_U);
  }
  GPS_CHECK(do_row2(ctx, f.part.p));
  f.loo_ok = true;
  f.pass2_done = true;
  ctx->launches += 3;
  if (grad_theta || grad_U) {
    MP_DISPATCH(f.MP, GPS_CHECK(run_small2<MPC>(ctx, f.part.p, f.acc2.p)));
    GPS_CHECK(do_row3(ctx, f.part.p));
    ctx->launches += 2;
    return fitc_finish_impl(ctx, f.part.p, f.grid, f.acc2.p, f.acc3.p, obj, grad_theta, grad_U);
  }
  // objective only: reduce pass 2 and finish with a zero pass-3 accumulator
  GPS_CHECK(reduce_to(ctx, f.part.p, f.grid, len2_of(f.MP), f.acc2.p));
  GPS_CUDA(cudaMemsetAsync(f.acc3.p, 0, len3_of(f.MP, ctx->D) * sizeof(double), ctx->stream));
  return fitc_finish_impl(ctx, nullptr, 0, f.acc2.p, f.acc3.p, obj, nullptr, nullptr);
}

int gps_fitc_loo(gps_ctx* ctx, double* loo_mean, double* loo_var) {
  if (!ctx) return GPS_EINVAL;
  auto& f = ctx->fitc;
  if (!f.pass2_done || !f.loo_ok) return gps_fail(ctx, GPS_ESTATE, "fitc_loo: no CRPS/LOGS/NLML evaluation to report");
  if (!loo_mean || !loo_var) return gps_fail(ctx, GPS_EINVAL, "fitc_loo: null output");
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t N = ctx->N;
  const bool dev = gps_is_device_ptr(loo_mean) && gps_is_device_ptr(loo_var);
  double *dm = loo_mean, *dv = loo_var;
  if (!dev) {
    GPS_CHECK(gps_ensure(ctx, ctx->stage[1], (size_t)N));
    GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)N));
    dm = ctx->stage[1].p;
    dv = ctx->stage[2].p;
  }
  if (f.fused) {
    GPS_CHECK(gps_fitc_fused_loo(ctx, dm, dv));
  } else {
    fitc_loo_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(ctx->y.p, f.rowv.p, N, dm, dv);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
  }
  if (!dev) {
    GPS_CUDA(cudaMemcpyAsync(loo_mean, dm, N * sizeof(double), cudaMemcpyDefault, ctx->stream));
    GPS_CUDA(cudaMemcpyAsync(loo_var, dv, N * sizeof(double), cudaMemcpyDefault, ctx->stream));
  }
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

int gps_fitc_predict(gps_ctx* ctx, const double* Xs, int64_t T, double* mean, double* var) {
  if (!ctx) return GPS_EINVAL;
  auto& f = ctx->fitc;
  if (!f.pass2_done) return gps_fail(ctx, GPS_ESTATE, "fitc_predict: run passes 1 and 2 at this theta, U first");
  if (!Xs || !mean || !var || T < 0) return gps_fail(ctx, GPS_EINVAL, "fitc_predict: bad arguments");
  if (T == 0) return GPS_OK;
  GPS_CUDA(cudaSetDevice(ctx->device));
  const double* dXs;
  GPS_CHECK(gps_stage_in(ctx, Xs, (size_t)T * ctx->D, ctx->stage[0], &dXs));
  const bool dev = gps_is_device_ptr(mean) && gps_is_device_ptr(var);
  double *dm = mean, *dv = var;
  if (!dev) {
    GPS_CHECK(gps_ensure(ctx, ctx->stage[1], (size_t)T));
    GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)T));
    dm = ctx->stage[1].p;
    dv = ctx->stage[2].p;
  }
  if (f.fused) {
    GPS_CHECK(gps_fitc_fused_predict(ctx, dXs, T, dm, dv));
  } else if (f.large) {
    GPS_CHECK(gps_fitc_large_predict(ctx, dXs, T, dm, dv));
  } else {
    MP_DISPATCH(f.MP, GPS_CHECK(run_pred<MPC>(ctx, dXs, T, dm, dv)));
    ctx->launches++;
  }
  if (!dev) {
    GPS_CUDA(cudaMemcpyAsync(mean, dm, T * sizeof(double), cudaMemcpyDefault, ctx->stream));
    GPS_CUDA(cudaMemcpyAsync(var, dv, T * sizeof(double), cudaMemcpyDefault, ctx->stream));
  }
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

}  // extern "C"
