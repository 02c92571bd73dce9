// FITC objective + gradient for M > 32 inducing points: the "matrix form" of the three row passes.
//
// The fused row kernels of gps_fitc.cu keep a row's M-vector in registers, which stops paying
// beyond M = 32.  BASELINE.json's scaling sweep goes to M = 1024, where the work is GEMM-shaped
// (2 N M^2 flops per product, ~10 products per evaluation), so this file runs the same Woodbury
// algebra (oracle/woodbury.py, replacing K20:222-236 / 329-344 / 434-452) on [Mp][Npp] row-major
// matrices with the task-list DMMA tile GEMM of gps_gemm.cu:
//
//   A = Kuu + jitter I = L_A L_A'          blocked POTRF/TRTRI on an M-sized child context
//   V = L_A^-1 Kuf                          GEMM  (lower-triangular k-range)
//   C = I + V diag(1/lam) V'                split-K GEMM with the k-scaling vector, fixed-order reduce
//   W = L_C^-1 V                            GEMM
//   R = W diag(rbar) W',  S = Vbar V'       split-K GEMMs
//   CV = Cbar V,  L_C^-T Wbar,  L_A^-T Vbar GEMMs (upper-triangular k-range for the transposed factors)
//   M x M adjoints (Cholesky adjoint Phi)   tile GEMMs on M x M operands + element-wise kernels
//
// Everything per-row (lambda, d, alpha, the score seeds, lambda_bar) is a column pass over the
// [Mp][Npp] matrices with i as the coalesced index.  Pad rows (m >= M) and pad columns (i >= N)
// hold zeros throughout, so no kernel needs bounds in the GEMMs.  LOO CRPS / log score and NLML on one
// GPU or row-sharded; the 4-fold block objectives (DSS K20:538-587, block CRPS "kc" K20:669-720) on one GPU:
// per fold the M x M matrices P_f = W_f Lam_f^-1 W_f', H_f = I - P_f (factorised on the child context) and
// two [Mp][fold columns] products (H^-1 W_f / Hbar_f W_f) — see the "block objectives" section below.
#include "gps_common.cuh"

namespace {

constexpr int DMAX = 16;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

enum { SM_KUU = 0, SM_LA, SM_LAI, SM_LC, SM_LCI, SM_R, SM_SW, SM_Y, SM_Z, SM_CBAR, SM_S, SM_ABAR, SM_COUNT,
       // block objectives only: L_H, L_H^-1, Hhat / Hbar, one scratch
       SM_LH = SM_COUNT, SM_LHI, SM_HX, SM_X2, SM_COUNT_BLOCK };
enum { RV_LAM = 0, RV_IL, RV_YL, RV_R, RV_ABAR, RV_DBAR, RV_LBAR, RV_RBAR, RV_TBAR, RV_LOOM, RV_LOOV, RV_COUNT };
// block objectives: aliases of slots the LOO scores use (R, DBAR, LOOM, LOOV are free there)
enum { RV_FV = RV_R, RV_MBAR = RV_DBAR, RV_CBARV = RV_LOOM, RV_ROWOBJ = RV_LOOV };
enum { MV_VY = 0, MV_BETA, MV_BBAR, MV_VYBAR, MV_G, MV_H, MV_HBAR, MV_GBAR, MV_COUNT };
// small results copied to the host: [0] obj, [1] sum lambda_bar, then two gradient blocks
constexpr int OUT_OBJ = 0, OUT_LOGDET = 2, OUT_G1 = 8;

}  // namespace

struct gps_fitc_large {
  int64_t Npp = 0;
  int Mp = 0, M = 0, D = 0;
  DevBuf Kuf, V, W, T1, T2;   // [Mp][Npp]
  DevBuf sm, rv, mv, part, out, U, dotp, acc;
  DevBuf fs;                  // block objectives: per-fold scalars [4][2] = (sum log diag L_H, g.h)
  GemmTask* ftasks = nullptr; // block objectives: split-K task lists of the four folds' column ranges
  size_t ftasks_cap = 0;
  gps_ctx::Range t_fold[4];
  int fold_S[4] = {0, 0, 0, 0};
  int64_t fold_key[10] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  int* latch = nullptr;       // device: [which matrix, pivot] of the first failed factorisation of the evaluation
  DevBuf fb;                  // block objectives: per-fold accumulators and replicated M x M matrices
  int64_t row_off = 0;        // first global row of this context's block (row-sharded block objectives)
  gps_allreduce_fn allreduce = nullptr;   // set for the duration of a row-sharded evaluation
  bool block_seeds = false;   // pass 2 left the block objectives' direct dL/dW in T2
  GemmTask* tasks = nullptr;
  size_t tasks_cap = 0;
  gps_ctx::Range t_low, t_up, t_full, t_mm, t_sk_low, t_sk_full;
  int S = 1;                  // split-K chunks
  gps_ctx* ch = nullptr;      // child context for the M-sized factorisations
  bool ready = false;
  int kuf_M = -1;             // live block of Kuf the pad zeros were laid out for
  int64_t kuf_N = -1;
  std::vector<double> h_out;
};

namespace {

// ---- element-wise / reduction kernels ------------------------------------------------------------

// K (Mp x Mp from the Gram kernel, pad zero) -> Kuu copy (no jitter) and A = Kuu + jitter I, identity pad
__global__ void __launch_bounds__(256)
kuu_fix_kernel(double* __restrict__ K, int Mp, int M, double jitter, double* __restrict__ Kuu) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  if (r < M && c < M) {
    const double v = K[e];
    Kuu[e] = v;
    K[e] = r == c ? v + jitter : v;
  } else {
    Kuu[e] = 0.0;
    K[e] = r == c ? 1.0 : 0.0;
  }
}

// dst = tril(src): the blocked factorisation leaves the strict upper block triangle undefined
__global__ void __launch_bounds__(256)
tri_copy_kernel(const double* __restrict__ src, double* __restrict__ dst, int Mp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  dst[e] = c <= r ? src[e] : 0.0;
}

// lam_i = e^a + sn2 - |V_:i|^2 ; il = 1/lam ; yl = y/lam   (pad columns: il = yl = 0)
__global__ void __launch_bounds__(256)
col_lambda_kernel(const double* __restrict__ V, int64_t ld, int M, int64_t N, int64_t Npp,
                  const double* __restrict__ y, const double* __restrict__ par, double* __restrict__ lam,
                  double* __restrict__ il, double* __restrict__ yl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npp) return;
  if (i >= N) {
    lam[i] = 1.0; il[i] = 0.0; yl[i] = 0.0;
    return;
  }
  double q = 0.0;
  for (int m = 0; m < M; ++m) {
    const double v = V[(int64_t)m * ld + i];
    q = fma(v, v, q);
  }
  const double l = par[0] - q + par[1];
  lam[i] = l;
  il[i] = 1.0 / l;
  yl[i] = y[i] / l;
}

// part[c][r] = sum over chunk c of Mat[r][i] x[i]   (grid = rows x chunks; fixed order)
__global__ void __launch_bounds__(256)
rowdot_kernel(const double* __restrict__ Mat, int64_t ld, int64_t n, const double* __restrict__ x,
              double* __restrict__ part) {
  __shared__ double sh[32];
  const double* row = Mat + (int64_t)blockIdx.x * ld;
  const int64_t per = ((n + gridDim.y - 1) / gridDim.y + 1) & ~(int64_t)1;
  const int64_t lo = blockIdx.y * per, hi = lo + per < n ? lo + per : n;
  double s0 = 0.0, s1 = 0.0;
  // rows and chunks start on even offsets of 16-byte aligned rows: 16-byte loads
  for (int64_t i = lo + 2 * threadIdx.x; i + 1 < hi; i += 2 * blockDim.x) {
    const double2 a = *reinterpret_cast<const double2*>(row + i);
    const double2 b = *reinterpret_cast<const double2*>(x + i);
    s0 = fma(a.x, b.x, s0);
    s1 = fma(a.y, b.y, s1);
  }
  if (threadIdx.x == 0 && ((hi - lo) & 1) && hi > lo) s0 = fma(row[hi - 1], x[hi - 1], s0);
  const double s = block_sum(s0 + s1, sh);
  if (threadIdx.x == 0) part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
rowdot_reduce_kernel(const double* __restrict__ part, int rows, int chunks, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += part[(int64_t)c * rows + r];
  out[r] = s;
}

// out[c] = sum_k Mat[k][c] x[k]: 16 columns x 16 k-slices per block, fixed-order combine (grid = Mp / 16)
__global__ void __launch_bounds__(256)
matvec_t_kernel(const double* __restrict__ Mat, int Mp, const double* __restrict__ x, double* __restrict__ out) {
  __shared__ double sh[16][17];
  const int cl = threadIdx.x & 15, ks = threadIdx.x >> 4;
  const int c = blockIdx.x * 16 + cl;
  double s = 0.0;
#pragma unroll 4
  for (int k = ks; k < Mp; k += 16) s = fma(Mat[(int64_t)k * Mp + c], x[k], s);
  sh[ks][cl] = s;
  __syncthreads();
  if (ks == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += sh[q][cl];
    out[c] = t;
  }
}

// out = sum_s part[s] (+ I); lower tiles only are valid when `lower`: the upper triangle is mirrored
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const double* __restrict__ part, int S, int Mp, double* __restrict__ out, int lower,
                     int add_identity) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t MM = (int64_t)Mp * Mp;
  if (e >= MM) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  int64_t src = e;
  if (lower && (c / GPS_TILE) > (r / GPS_TILE)) src = (int64_t)c * Mp + r;
  double s = 0.0;
  for (int k = 0; k < S; ++k) s += part[(int64_t)k * MM + src];
  if (add_identity && r == c) s += 1.0;
  out[e] = s;
}

// r_i = |W_:i|^2, alpha_i = (y_i - W_:i . beta)/lam_i, d_i = 1/lam_i - r_i/lam_i^2
__global__ void __launch_bounds__(256)
col_w_kernel(const double* __restrict__ W, int64_t ld, int M, int64_t N, int64_t Npp,
             const double* __restrict__ beta, const double* __restrict__ y, const double* __restrict__ il,
             double* __restrict__ rr, double* __restrict__ alpha, double* __restrict__ dd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npp) return;
  if (i >= N) {
    rr[i] = 0.0;
    return;
  }
  double r = 0.0, wb = 0.0;
  for (int m = 0; m < M; ++m) {
    const double w = W[(int64_t)m * ld + i];
    r = fma(w, w, r);
    wb = fma(w, beta[m], wb);
  }
  const double l = il[i];
  rr[i] = r;
  alpha[i] = (y[i] - wb) * l;
  dd[i] = l - r * l * l;
}

// NLML over this context's rows: obj share = sum 1/2 log lam + 1/2 y alpha; seeds abar = y/2, dbar = 0
// (the replicated terms N/2 log 2pi + sum log diag L_C are added once, on the host)
__global__ void __launch_bounds__(1024)
nlml_rows_kernel(int64_t N, int64_t Npp, const double* __restrict__ lam, const double* __restrict__ y,
                 const double* __restrict__ alpha, double* __restrict__ abar, double* __restrict__ dbar,
                 double* __restrict__ obj) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < Npp; i += blockDim.x) {
    if (i < N) {
      s += 0.5 * log(lam[i]) + 0.5 * y[i] * alpha[i];
      abar[i] = 0.5 * y[i];
    } else {
      abar[i] = 0.0;
    }
    dbar[i] = 0.0;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) obj[0] = s;
}

__global__ void __launch_bounds__(256)
logdiag_sum_kernel(const double* __restrict__ L, int Mp, int M, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) s += log(L[(int64_t)m * Mp + m]);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}

// out = A + I
__global__ void __launch_bounds__(256)
add_identity_kernel(const double* __restrict__ A, int Mp, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  out[e] = A[e] + ((e / Mp) == (e % Mp) ? 1.0 : 0.0);
}

// lam_bar0, rbar, tbar from the score seeds
__global__ void __launch_bounds__(256)
seed_kernel(int64_t N, int64_t Npp, int nlml, const double* __restrict__ il, const double* __restrict__ rr,
            const double* __restrict__ alpha, const double* __restrict__ abar, const double* __restrict__ dbar,
            double* __restrict__ lbar, double* __restrict__ rbar, double* __restrict__ tbar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npp) return;
  if (i >= N) {
    lbar[i] = 0.0; rbar[i] = 0.0; tbar[i] = 0.0;
    return;
  }
  const double l = il[i], ab = abar[i], db = dbar[i], r = rr[i];
  double lb = nlml ? 0.5 * l : 0.0;
  lb += db * (-l * l + 2.0 * r * l * l * l) - ab * alpha[i] * l;
  lbar[i] = lb;
  rbar[i] = -db * l * l;
  tbar[i] = -ab * l;
}

// SW = beta bbar' + rs R + bbar beta'   (rs = 2 for R = W diag(rbar) W', 1 for the block objectives' G_W)
__global__ void __launch_bounds__(256)
sw_kernel(const double* __restrict__ beta, const double* __restrict__ bbar, const double* __restrict__ R, int Mp,
          double rs, double* __restrict__ SW) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  SW[e] = beta[r] * bbar[c] + rs * R[e] + bbar[r] * beta[c];
}

// Lbar = -tril(Y) (+ diag(1/L_mm) for the log-determinant term of the NLML)
__global__ void __launch_bounds__(256)
lbar_kernel(const double* __restrict__ Y, const double* __restrict__ L, int Mp, int M, int add_diag,
            double* __restrict__ Lbar) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  double v = c <= r ? -Y[e] : 0.0;
  if (add_diag && r == c && r < M) v += 1.0 / L[e];
  Lbar[e] = v;
}

// Z = Phi(P) + Phi(P)' with Phi = lower triangle, diagonal halved
__global__ void __launch_bounds__(256)
phi_sym_kernel(const double* __restrict__ P, int Mp, double* __restrict__ Z) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  Z[e] = c <= r ? P[e] : P[(int64_t)c * Mp + r];
}

// ---- block objectives (4-fold DSS K20:538-587, block CRPS "kc" K20:669-720): element-wise kernels ---------
// Algebra: oracle/woodbury.py::fitc_block_obj_grad.  A fold is a contiguous range [lo, hi) of columns.

constexpr double INV_SQRT_PI = 0.56418958354775628695;
constexpr double INV_SQRT_2PI = 0.39894228040143267794;
constexpr double INV_SQRT2 = 0.70710678118654752440;

// out[i] = src[i] inside [lo, hi), 0 in the rest of the 16-rounded range [lo_r, hi_r) the GEMM k-range covers
__global__ void __launch_bounds__(256)
mask_range_kernel(const double* __restrict__ src, int64_t lo, int64_t hi, int64_t lo_r, int64_t hi_r,
                  double* __restrict__ out) {
  const int64_t i = lo_r + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi_r) return;
  out[i] = (i >= lo && i < hi) ? src[i] : 0.0;
}

// H = I - P
__global__ void __launch_bounds__(256)
h_from_p_kernel(const double* __restrict__ P, int Mp, double* __restrict__ H) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  H[e] = ((e / Mp) == (e % Mp) ? 1.0 : 0.0) - P[e];
}

// fs[0] = sum_m log diag(L_H), fs[1] = g . h
__global__ void __launch_bounds__(256)
fold_scalar_kernel(const double* __restrict__ LH, int Mp, int M, const double* __restrict__ g,
                   const double* __restrict__ h, double* __restrict__ fs) {
  __shared__ double sh[32];
  double s = 0.0, t = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    s += log(LH[(int64_t)m * Mp + m]);
    t = fma(g[m], h[m], t);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) fs[0] = s;
  t = block_sum(t, sh);
  if (threadIdx.x == 0) fs[1] = t;
}

// DSS: Hhat = -1/2 H^-1 - 1/2 h h'
__global__ void __launch_bounds__(256)
hhat_kernel(const double* __restrict__ Hinv, const double* __restrict__ h, int Mp, int M, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  out[e] = (r < M && c < M) ? -0.5 * Hinv[e] - 0.5 * h[r] * h[c] : 0.0;
}

// kc: Hbar = -(H^-1 E H^-1) - 1/2 (gbar h' + h gbar')
__global__ void __launch_bounds__(256)
hbar_kernel(const double* __restrict__ Z, const double* __restrict__ h, const double* __restrict__ gbar, int Mp, int M,
            double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  out[e] = (r < M && c < M) ? -Z[e] - 0.5 * (gbar[r] * h[c] + h[r] * gbar[c]) : 0.0;
}

// x <- -x (hbar = -(W_f mbar))
__global__ void __launch_bounds__(256)
negate_kernel(double* __restrict__ x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = -x[i];
}

// DSS column sweep of one fold: T = Hhat W_f
//   abar_i = lam_i alpha_i + W_i.h ; lam_bar0_i = 1/(2 lam) + alpha^2/2 + (W_i.T_i)/lam^2
//   D_i = -2 T_i/lam_i + h alpha_i ; row objective 1/2 log lam + 1/2 lam alpha^2
__global__ void __launch_bounds__(256)
col_dss_kernel(const double* __restrict__ W, const double* __restrict__ T, double* __restrict__ Dm, int64_t ld, int M,
               int64_t lo, int64_t hi, const double* __restrict__ h, const double* __restrict__ lam,
               const double* __restrict__ alpha, double* __restrict__ abar, double* __restrict__ lbar0,
               double* __restrict__ rowobj) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const double l = lam[i], il = 1.0 / l, al = alpha[i];
  rowobj[i] = 0.5 * log(l) + 0.5 * l * al * al;
  if (!T) return;                                      // objective only
  double hw = 0.0, q = 0.0;
  for (int m = 0; m < M; ++m) {
    const int64_t o = (int64_t)m * ld + i;
    const double w = W[o], t = T[o];
    hw = fma(w, h[m], hw);
    q = fma(w, t, q);
    Dm[o] = fma(-2.0 * il, t, h[m] * al);
  }
  abar[i] = l * al + hw;
  lbar0[i] = 0.5 * il + 0.5 * al * al + q * il * il;
}

// kc column sweep A of one fold: T = H^-1 W_f.  Fold predictive mu_i = y_i - lam_i alpha_i - W_i.h,
// c_i = lam_i + W_i.T_i; CRPS share (mean over the fold), seeds mbar_i, cbar_i; D_i = -h mbar_i + 2 T_i cbar_i
__global__ void __launch_bounds__(256)
col_kca_kernel(const double* __restrict__ W, const double* __restrict__ T, double* __restrict__ Dm, int64_t ld, int M,
               int64_t lo, int64_t hi, const double* __restrict__ h, const double* __restrict__ lam,
               const double* __restrict__ alpha, double inv_nf, double* __restrict__ mbar, double* __restrict__ cbar,
               double* __restrict__ rowobj) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  double hw = 0.0, q = 0.0;
  for (int m = 0; m < M; ++m) {
    const int64_t o = (int64_t)m * ld + i;
    const double w = W[o];
    hw = fma(w, h[m], hw);
    q = fma(w, T[o], q);
  }
  const double l = lam[i];
  const double sd = sqrt(l + q);
  const double z = (l * alpha[i] + hw) / sd;          // (y - mu)/sd
  const double tpm1 = erf(z * INV_SQRT2);
  const double pdf = INV_SQRT_2PI * exp(-0.5 * z * z);
  rowobj[i] = sd * (z * tpm1 + 2.0 * pdf - INV_SQRT_PI) * inv_nf;
  const double mb = -tpm1 * inv_nf;
  const double cb = (2.0 * pdf - INV_SQRT_PI) / (2.0 * sd) * inv_nf;
  mbar[i] = mb;
  cbar[i] = cb;
  for (int m = 0; m < M; ++m) {
    const int64_t o = (int64_t)m * ld + i;
    Dm[o] = fma(2.0 * cb, T[o], -h[m] * mb);
  }
}

// kc column sweep B: T = Hbar W_f.  abar_i = -mbar lam + W_i.gbar ; lam_bar0 = -mbar alpha + cbar + (W_i.T_i)/lam^2 ;
// D_i += gbar alpha_i - 2 T_i/lam_i
__global__ void __launch_bounds__(256)
col_kcb_kernel(const double* __restrict__ W, const double* __restrict__ T, double* __restrict__ Dm, int64_t ld, int M,
               int64_t lo, int64_t hi, const double* __restrict__ gbar, const double* __restrict__ lam,
               const double* __restrict__ alpha, const double* __restrict__ mbar, const double* __restrict__ cbar,
               double* __restrict__ abar, double* __restrict__ lbar0) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const double l = lam[i], il = 1.0 / l, al = alpha[i];
  double gw = 0.0, q = 0.0;
  for (int m = 0; m < M; ++m) {
    const int64_t o = (int64_t)m * ld + i;
    const double w = W[o], t = T[o];
    gw = fma(w, gbar[m], gw);
    q = fma(w, t, q);
    Dm[o] += fma(-2.0 * il, t, gbar[m] * al);
  }
  abar[i] = -mbar[i] * l + gw;
  lbar0[i] = -mbar[i] * al + cbar[i] + q * il * il;
}

// DSS: GW += X (= -2 Hhat P_f) + h g' ; bbar -= h
__global__ void __launch_bounds__(256)
gw_dss_kernel(const double* __restrict__ X, const double* __restrict__ h, const double* __restrict__ g, int Mp,
              double* __restrict__ GW, double* __restrict__ bbar) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  GW[e] += X[e] + h[r] * g[c];
  if (c == 0) bbar[r] -= h[r];
}

// kc: GW += h hbar' + 2 Y (= H^-1 E) + gbar g' + X (= -2 Hbar P_f) ; bbar += -hbar - P_f gbar
__global__ void __launch_bounds__(256)
gw_kc_kernel(const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ P,
             const double* __restrict__ h, const double* __restrict__ hbar, const double* __restrict__ g,
             const double* __restrict__ gbar, int Mp, double* __restrict__ GW, double* __restrict__ bbar) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int r = (int)(e / Mp), c = (int)(e - (int64_t)r * Mp);
  GW[e] += h[r] * hbar[c] + 2.0 * Y[e] + gbar[r] * g[c] + X[e];
  if (c == 0) {
    double s = 0.0;
    for (int k = 0; k < Mp; ++k) s = fma(P[(int64_t)r * Mp + k], gbar[k], s);
    bbar[r] -= hbar[r] + s;
  }
}

// seeds of the shared adjoint chain: tbar = -abar/lam ; lam_bar = lam_bar0 - abar alpha/lam ; rbar = 0
__global__ void __launch_bounds__(256)
block_seed_kernel(int64_t N, int64_t Npp, const double* __restrict__ il, const double* __restrict__ alpha,
                  const double* __restrict__ abar, double* __restrict__ lbar, double* __restrict__ rbar,
                  double* __restrict__ tbar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npp) return;
  rbar[i] = 0.0;
  if (i >= N) {
    lbar[i] = 0.0; tbar[i] = 0.0;
    return;
  }
  const double l = il[i], ab = abar[i];
  lbar[i] -= ab * alpha[i] * l;
  tbar[i] = -ab * l;
}

// objective: rows' sum (+ DSS: sum_f nf/2 log 2pi - sum log diag L_H_f + 1/2 g_f.h_f), fixed order
__global__ void block_obj_kernel(const double* rows, const double* __restrict__ fs, int dss, double nf, double* obj) {
  double v = rows[0];
  if (dss)
    for (int f = 0; f < 4; ++f) v += nf * HALF_LOG_2PI - fs[2 * f] + 0.5 * fs[2 * f + 1];
  obj[0] = v;
}

// pass 3 column sweep: lam_bar -= (bbar . W_:i) y_i/lam^2 + (V_:i . CV_:i)/lam^2 ; W <- Wbar in place
__global__ void __launch_bounds__(256)
col_pass3_kernel(const double* __restrict__ V, const double* __restrict__ CV, double* __restrict__ W, int64_t ld,
                 int M, int64_t N, const double* __restrict__ bbar, const double* __restrict__ beta,
                 const double* __restrict__ y, const double* __restrict__ il, const double* __restrict__ tbar,
                 const double* __restrict__ rbar, double* __restrict__ lbar, const double* __restrict__ Dm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double tb = tbar[i], rb2 = 2.0 * rbar[i];
  double bw = 0.0, s1 = 0.0;
  for (int m = 0; m < M; ++m) {
    const int64_t o = (int64_t)m * ld + i;
    const double w = W[o];
    bw = fma(bbar[m], w, bw);
    s1 = fma(V[o], CV[o], s1);
    // LOO scores: Wbar = beta tbar' + 2 W diag(rbar); block objectives: beta tbar' + D (the direct dL/dW)
    W[o] = Dm ? fma(tb, beta[m], Dm[o]) : fma(rb2, w, tb * beta[m]);
  }
  const double l = il[i];
  lbar[i] -= (bw * y[i] + s1) * l * l;
}

// Vbar = L_C^-T Wbar (already in T2) + vybar yl' + 2 CV diag(il) - 2 V diag(lam_bar)
__global__ void __launch_bounds__(256)
vbar_kernel(double* __restrict__ T2, const double* __restrict__ CV, const double* __restrict__ V, int64_t ld, int M,
            int64_t N, const double* __restrict__ vybar, const double* __restrict__ yl,
            const double* __restrict__ il, const double* __restrict__ lbar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (i >= N || m >= M) return;
  const int64_t o = (int64_t)m * ld + i;
  T2[o] += vybar[m] * yl[i] + 2.0 * CV[o] * il[i] - 2.0 * V[o] * lbar[i];
}

__global__ void __launch_bounds__(1024)
vec_sum_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}

// Kernel-gradient partials of G = Kbar o Kmat ([M][n], row stride ld) against the points P[n][D]:
//   sg[m] = sum_i G_mi,  sd[m][d] = sum_i G_mi (u_md - p_id)/l_d,  sb[d] = sum_mi G_mi ((u_md - p_id)/l_d)^2
// One block = MT rows of G x one strided share of the columns; per-block partials in fixed slots.
template <int MT, int DMX, bool DPOW2>
__global__ void __launch_bounds__(256, 2)
kgrad_kernel(const double* __restrict__ Kbar, const double* __restrict__ Kmat, int64_t ld, int M, int64_t n,
             const double* __restrict__ U, const double* __restrict__ P, int D, const double* __restrict__ par,
             double* __restrict__ part) {
  constexpr int NV = MT * (1 + DMX) + DMX;
  __shared__ double us[MT][DMX];
  __shared__ double ils[DMX];
  __shared__ double xs[DMX][257];
  __shared__ double red[8][NV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * MT;
  if (tid < MT * DMX) {
    const int mm = tid / DMX, d = tid % DMX;
    us[mm][d] = (m0 + mm < M && d < D) ? U[(int64_t)(m0 + mm) * D + d] * par[2 + d] : 0.0;
  }
  if (tid < DMX) ils[tid] = tid < D ? par[2 + tid] : 0.0;
  __syncthreads();
  double sg[MT], sd[MT][DMX], sb[DMX];
#pragma unroll
  for (int mm = 0; mm < MT; ++mm) {
    sg[mm] = 0.0;
#pragma unroll
    for (int d = 0; d < DMX; ++d) sd[mm][d] = 0.0;
  }
#pragma unroll
  for (int d = 0; d < DMX; ++d) sb[d] = 0.0;
  // A tile of 256 points is staged (scaled, transposed) in shared memory with coalesced loads (a
  // thread reading its own row of P would touch 16 cache lines per warp-load); the next tile's
  // operands are fetched into registers while the current one is consumed.
  const int64_t stride = (int64_t)256 * gridDim.y;
  double gn[MT], xr[DMX];
  // element k of this thread in the [256][D] tile: row / shared-memory slot, fixed for the whole kernel
  // (computed once: an integer division by the run-time D inside the loop costs more than the arithmetic)
  int rk[DMX], sk[DMX];
#pragma unroll
  for (int k = 0; k < DMX; ++k) {
    const int e = tid + k * 256;
    const int r = DPOW2 ? e / DMX : e / D, d = DPOW2 ? e % DMX : e - r * D;   // DPOW2: D == DMX, shifts
    rk[k] = r;
    sk[k] = d * 257 + r;
  }
  auto fetch = [&](int64_t base) {
    const int64_t i = base + tid;
#pragma unroll
    for (int mm = 0; mm < MT; ++mm) {
      const int64_t o = (int64_t)(m0 + mm) * ld + i;
      const bool live = i < n && m0 + mm < M;
      const double a = live ? Kbar[o] : 0.0, b = live ? Kmat[o] : 0.0;
      gn[mm] = a * b;
    }
#pragma unroll
    for (int k = 0; k < DMX; ++k)
      xr[k] = (k < D && base + rk[k] < n) ? P[base * D + tid + k * 256] : 0.0;
  };
  int64_t base = (int64_t)blockIdx.y * 256;
  if (base < n) fetch(base);
#pragma unroll 1
  for (; base < n; base += stride) {
    double gv[MT];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DMX; ++k)
      if (k < D) (&xs[0][0])[sk[k]] = xr[k];
#pragma unroll
    for (int mm = 0; mm < MT; ++mm) gv[mm] = gn[mm];
    __syncthreads();
    if (base + stride < n) fetch(base + stride);
#pragma unroll
    for (int mm = 0; mm < MT; ++mm) sg[mm] += gv[mm];
#pragma unroll
    for (int d = 0; d < DMX; ++d) {
      if (d < D) {
        const double xv = xs[d][tid] * ils[d];
#pragma unroll
        for (int mm = 0; mm < MT; ++mm) {
          const double df = us[mm][d] - xv;
          const double gd = gv[mm] * df;
          sd[mm][d] += gd;
          sb[d] = fma(gd, df, sb[d]);
        }
      }
    }
  }
  // block reduction of the NV accumulators
#pragma unroll
  for (int mm = 0; mm < MT; ++mm) {
    const double v = warp_sum(sg[mm]);
    if (lane == 0) red[warp][mm * (1 + DMX)] = v;
#pragma unroll
    for (int d = 0; d < DMX; ++d) {
      const double w = warp_sum(sd[mm][d]);
      if (lane == 0) red[warp][mm * (1 + DMX) + 1 + d] = w;
    }
  }
#pragma unroll
  for (int d = 0; d < DMX; ++d) {
    const double w = warp_sum(sb[d]);
    if (lane == 0) red[warp][MT * (1 + DMX) + d] = w;
  }
  __syncthreads();
  if (tid < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NV + tid] = s;
  }
}

// out[0] = sum G, out[1 + d] = g_b partial, out[1 + DMAX + m*D + d] = -(1/l_d) sum_i G (u - p)/l
template <int MT, int DMX>
__global__ void __launch_bounds__(256)
kgrad_reduce_kernel(const double* __restrict__ part, int gx, int gy, int M, int D, const double* __restrict__ par,
                    double* __restrict__ out) {
  constexpr int NV = MT * (1 + DMX) + DMX;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 1 + DMAX) return;                        // the D + 1 global sums: kgrad_globals_kernel
  {
    const int e = t - (1 + DMAX);
    if (e >= M * D) return;
    const int m = e / D, d = e - m * D;
    const int bx = m / MT, mm = m - bx * MT;
    double s = 0.0;
    for (int by = 0; by < gy; ++by) s += part[((int64_t)by * gx + bx) * NV + mm * (1 + DMX) + 1 + d];
    out[1 + DMAX + e] = -s * par[2 + d];
  }
}

// the D + 1 sums over ALL partial blocks (block 0: sum G -> out[0]; block 1 + d: g_b partial -> out[1 + d]): every
// thread takes a fixed strided share, then a fixed-order block sum (was one thread per sum: 0.18 ms of latency)
template <int MT, int DMX>
__global__ void __launch_bounds__(256)
kgrad_globals_kernel(const double* __restrict__ part, int gx, int gy, double* __restrict__ out) {
  constexpr int NV = MT * (1 + DMX) + DMX;
  __shared__ double sh[32];
  const int64_t nb = (int64_t)gx * gy;
  double s = 0.0;
  if (blockIdx.x == 0) {
    for (int64_t e = threadIdx.x; e < nb * MT; e += blockDim.x) s += part[(e / MT) * NV + (e % MT) * (1 + DMX)];
  } else {
    const int d = blockIdx.x - 1;
    for (int64_t e = threadIdx.x; e < nb; e += blockDim.x) s += part[e * NV + MT * (1 + DMX) + d];
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// prediction columns: mean = Ws_:t . beta, var = sn2 + e^a - |Vs_:t|^2 + |Ws_:t|^2   (K20:76-83)
__global__ void __launch_bounds__(256)
col_predict_kernel(const double* __restrict__ Vs, const double* __restrict__ Ws, int64_t ld, int M, int64_t T,
                   const double* __restrict__ beta, const double* __restrict__ par, double* __restrict__ mean,
                   double* __restrict__ var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  double q = 0.0, r = 0.0, mb = 0.0;
  for (int m = 0; m < M; ++m) {
    const double v = Vs[(int64_t)m * ld + i], w = Ws[(int64_t)m * ld + i];
    q = fma(v, v, q);
    r = fma(w, w, r);
    mb = fma(w, beta[m], mb);
  }
  mean[i] = mb;
  var[i] = par[1] + par[0] - q + r;
}

// ---- host side -------------------------------------------------------------------------------------

inline unsigned blocks_for(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

GemmTask make_task(int a_row, int b_row, int k0, int k1, int c_row, int c_col, int flags = 0) {
  GemmTask t;
  t.a_row = a_row; t.b_row = b_row; t.k0 = k0; t.k1 = k1; t.c_row = c_row; t.c_col = c_col;
  t.flags = flags; t.pad = 0;
  return t;
}

// task lists for the shape (Mp, Npp); cached on the device
int build_tasks(gps_ctx* ctx, gps_fitc_large* fl) {
  const int mt = fl->Mp / GPS_TILE;
  const int64_t nt = fl->Npp / GPS_TILE;
  const int Mp = fl->Mp;
  std::vector<GemmTask> h;
  auto begin = [&](gps_ctx::Range& r) { r.off = h.size(); };
  auto end = [&](gps_ctx::Range& r) { r.cnt = h.size() - r.off; };
  // [Mp][Npp] outputs: row tile fastest so consecutive CTAs share the column panel of the big operand
  begin(fl->t_low);
  for (int64_t j = 0; j < nt; ++j)
    for (int i = 0; i < mt; ++i)
      h.push_back(make_task(i * GPS_TILE, (int)(j * GPS_TILE), 0, (i + 1) * GPS_TILE, i * GPS_TILE, (int)(j * GPS_TILE),
                            GEMM_TRI_END));     // A = a lower-triangular factor inverse: row strips stop at their diagonal
  end(fl->t_low);
  begin(fl->t_up);
  for (int64_t j = 0; j < nt; ++j)
    for (int i = 0; i < mt; ++i)
      h.push_back(make_task(i * GPS_TILE, (int)(j * GPS_TILE), i * GPS_TILE, Mp, i * GPS_TILE, (int)(j * GPS_TILE),
                            GEMM_TRI_BEGIN));   // A' with A lower-triangular: row strips start at their diagonal
  end(fl->t_up);
  begin(fl->t_full);
  for (int64_t j = 0; j < nt; ++j)
    for (int i = 0; i < mt; ++i)
      h.push_back(make_task(i * GPS_TILE, (int)(j * GPS_TILE), 0, Mp, i * GPS_TILE, (int)(j * GPS_TILE)));
  end(fl->t_full);
  begin(fl->t_mm);
  for (int i = 0; i < mt; ++i)
    for (int j = 0; j < mt; ++j)
      h.push_back(make_task(i * GPS_TILE, j * GPS_TILE, 0, Mp, i * GPS_TILE, j * GPS_TILE));
  end(fl->t_mm);
  // split-K over the rows of the data set: the chunk count that fills whole waves of task slots
  // (one slot = one 128 x 128 task = two CTAs of the shipped policy, sm_count slots per wave)
  const int tl = mt * (mt + 1) / 2;
  const int64_t slots = ctx->sm_count;
  int64_t want = 1;
  double best = -1.0;
  const int64_t s_lo = std::min<int64_t>(nt, (slots + tl - 1) / tl), s_hi = std::min<int64_t>(nt, (16 * slots + tl - 1) / tl);
  for (int64_t s = s_lo; s <= s_hi; ++s) {
    const int64_t tasks = s * tl, waves = (tasks + slots - 1) / slots;
    const double eff = (double)tasks / (double)(waves * slots);
    if (eff > best + 1e-9) { best = eff; want = s; }
  }
  if (want > nt) want = nt;
  const int64_t per = (nt + want - 1) / want;
  fl->S = (int)((nt + per - 1) / per);
  begin(fl->t_sk_low);
  for (int s = 0; s < fl->S; ++s) {
    const int k0 = (int)(s * per * GPS_TILE), k1 = (int)std::min<int64_t>(fl->Npp, (s + 1) * per * GPS_TILE);
    for (int i = 0; i < mt; ++i)
      for (int j = 0; j <= i; ++j)
        h.push_back(make_task(i * GPS_TILE, j * GPS_TILE, k0, k1, s * Mp + i * GPS_TILE, j * GPS_TILE));
  }
  end(fl->t_sk_low);
  begin(fl->t_sk_full);
  for (int s = 0; s < fl->S; ++s) {
    const int k0 = (int)(s * per * GPS_TILE), k1 = (int)std::min<int64_t>(fl->Npp, (s + 1) * per * GPS_TILE);
    for (int i = 0; i < mt; ++i)
      for (int j = 0; j < mt; ++j)
        h.push_back(make_task(i * GPS_TILE, j * GPS_TILE, k0, k1, s * Mp + i * GPS_TILE, j * GPS_TILE));
  }
  end(fl->t_sk_full);
  if (h.size() > fl->tasks_cap) {
    if (fl->tasks) cudaFree(fl->tasks);
    fl->tasks = nullptr;
    GPS_CUDA(cudaMalloc(&fl->tasks, h.size() * sizeof(GemmTask)));
    fl->tasks_cap = h.size();
  }
  GPS_CUDA(cudaMemcpyAsync(fl->tasks, h.data(), h.size() * sizeof(GemmTask), cudaMemcpyHostToDevice, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

int ensure_child(gps_ctx* ctx, gps_fitc_large* fl) {
  if (!fl->ch) {
    fl->ch = new gps_ctx();
    fl->ch->device = ctx->device;
    fl->ch->sm_count = ctx->sm_count;
  }
  gps_ctx* ch = fl->ch;
  ch->gemm_variant = ctx->gemm_variant;
  ch->potf2_variant = ctx->potf2_variant;
  ch->stream = ctx->stream;
  ch->own_stream = nullptr;
  ch->time_gemm = false;
  const int r = gps_ensure_ws(ch, fl->Mp);
  if (r != GPS_OK) return gps_fail(ctx, r, "fitc (matrix form): %s", ch->err.c_str());
  return GPS_OK;
}

int setup(gps_ctx* ctx, int M, bool block) {
  if (!ctx->fl) ctx->fl = new gps_fitc_large();
  gps_fitc_large* fl = ctx->fl;
  const int Mp = (int)gps_pad(M);
  const int64_t Npp = ctx->Np;
  const int D = ctx->D;
  const bool reshape = fl->Mp != Mp || fl->Npp != Npp || fl->D != D;
  fl->M = M;
  if (reshape) {
    fl->Mp = Mp; fl->Npp = Npp; fl->D = D;
    const size_t big = (size_t)Mp * Npp;
    GPS_CHECK(gps_ensure(ctx, fl->Kuf, big));
    GPS_CHECK(gps_ensure(ctx, fl->V, big));
    GPS_CHECK(gps_ensure(ctx, fl->W, big));
    GPS_CHECK(gps_ensure(ctx, fl->T1, big));
    GPS_CHECK(gps_ensure(ctx, fl->T2, big));
    GPS_CHECK(gps_ensure(ctx, fl->rv, (size_t)RV_COUNT * Npp));
    GPS_CHECK(gps_ensure(ctx, fl->mv, (size_t)MV_COUNT * Mp));
    GPS_CHECK(gps_ensure(ctx, fl->U, (size_t)Mp * D));
    GPS_CHECK(build_tasks(ctx, fl));
    const size_t kg = (size_t)8 * ctx->sm_count * (8 * (1 + DMAX) + DMAX) + 4096;
    GPS_CHECK(gps_ensure(ctx, fl->part, std::max((size_t)fl->S * Mp * Mp, kg)));
    GPS_CHECK(gps_ensure(ctx, fl->out, (size_t)OUT_G1 + 2 * (1 + DMAX + (size_t)Mp * D)));
  }
  // per-row scalars are laid out with stride N (the layout gps_fitc_loo reads): N may grow inside one padded shape
  GPS_CHECK(gps_ensure(ctx, ctx->fitc.rowv, (size_t)6 * ctx->N));
  // nothing in `sm` outlives an evaluation except the factors begin() writes after this point
  GPS_CHECK(gps_ensure(ctx, fl->sm, (size_t)(block ? SM_COUNT_BLOCK : SM_COUNT) * Mp * Mp));
  // the Gram kernel writes only the live M x N block: the pad rows/columns must read as zeros
  if (reshape || fl->kuf_M != M || fl->kuf_N != ctx->N) {
    GPS_CUDA(cudaMemsetAsync(fl->Kuf.p, 0, (size_t)Mp * Npp * sizeof(double), ctx->stream));
    fl->kuf_M = M;
    fl->kuf_N = ctx->N;
  }
  GPS_CHECK(ensure_child(ctx, fl));
  return GPS_OK;
}

// out[r] = Mat[r][:n] . x for r < rows
int rowdot(gps_ctx* ctx, gps_fitc_large* fl, const double* Mat, int64_t ld, int rows, int64_t n, const double* x,
           double* out) {
  int64_t chunks = (4 * (int64_t)ctx->sm_count + rows - 1) / rows;
  chunks = std::max<int64_t>(1, std::min<int64_t>(chunks, n / 4096 + 1));
  GPS_CHECK(gps_ensure(ctx, fl->dotp, (size_t)rows * chunks));
  rowdot_kernel<<<dim3(rows, (unsigned)chunks), 256, 0, ctx->stream>>>(Mat, ld, n, x, fl->dotp.p);
  GPS_LAUNCH_CHECK();
  rowdot_reduce_kernel<<<blocks_for(rows), 256, 0, ctx->stream>>>(fl->dotp.p, rows, (int)chunks, out);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  return GPS_OK;
}

int mm_gemm(gps_ctx* ctx, gps_fitc_large* fl, int kind, const double* A, const double* B, double* C, double alpha) {
  return gps_gemm_tasks(ctx, kind, A, fl->Mp, B, fl->Mp, C, fl->Mp, alpha, 0.0, nullptr, false,
                        fl->tasks + fl->t_mm.off, fl->t_mm.cnt);
}

int big_gemm(gps_ctx* ctx, gps_fitc_large* fl, int kind, const gps_ctx::Range& r, const double* A, const double* B,
             double* C) {
  return gps_gemm_tasks(ctx, kind, A, fl->Mp, B, fl->Npp, C, fl->Npp, 1.0, 0.0, nullptr, false, fl->tasks + r.off, r.cnt);
}

// out (Mp x Mp) = A diag(dvec) B' over the data rows, split-K with a fixed-order reduction
int splitk(gps_ctx* ctx, gps_fitc_large* fl, const double* A, const double* B, const double* dvec, bool lower,
           bool add_identity, double* out) {
  const gps_ctx::Range& r = lower ? fl->t_sk_low : fl->t_sk_full;
  GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_KC, A, fl->Npp, B, fl->Npp, fl->part.p, fl->Mp, 1.0, 0.0, dvec, false,
                           fl->tasks + r.off, r.cnt));
  splitk_reduce_kernel<<<blocks_for((int64_t)fl->Mp * fl->Mp), 256, 0, ctx->stream>>>(fl->part.p, fl->S, fl->Mp, out,
                                                                                      lower ? 1 : 0, add_identity ? 1 : 0);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

// first failed factorisation of an evaluation: latch[0] = which matrix (1 K_uu + jitter I, 2 I + V Lambda^-1 V',
// 3 + f = H_f of fold f), latch[1] = the failing pivot
__global__ void latch_info_kernel(const int* __restrict__ info, int* __restrict__ latch, int code) {
  if (*info != 0 && latch[0] == 0) {
    latch[0] = code;
    latch[1] = *info;
  }
}

// factor the matrix in ch->Kb: L -> Lout, L^-1 -> Linv (both with an explicit zero upper triangle).  Nothing is read
// back here: a failed pivot (the diagonal kernel substitutes 1 and carries on, so everything downstream stays finite
// or NaN but never hangs) is latched on the device and reported by gps_fitc_large_finish with the result read-back —
// no host synchronisation inside an evaluation, and in row-sharded runs every rank reaches every all-reduce.
int factor(gps_ctx* ctx, gps_fitc_large* fl, double* Lout, double* Linv, int code) {
  gps_ctx* ch = fl->ch;
  const int Mp = fl->Mp;
  GPS_CUDA(cudaMemsetAsync(ch->d_info, 0, sizeof(int), ctx->stream));
  int r = gps_potrf(ch, ch->Kb.p, ch->Xb.p, Mp);
  if (r == GPS_OK) r = gps_trtri(ch, ch->Kb.p, ch->Xb.p, ch->Sb.p, Mp);
  if (r != GPS_OK) return gps_fail(ctx, r, "fitc factorisation %d: %s", code, ch->err.c_str());
  const unsigned nb = blocks_for((int64_t)Mp * Mp);
  tri_copy_kernel<<<nb, 256, 0, ctx->stream>>>(ch->Kb.p, Lout, Mp);
  GPS_LAUNCH_CHECK();
  tri_copy_kernel<<<nb, 256, 0, ctx->stream>>>(ch->Xb.p, Linv, Mp);
  GPS_LAUNCH_CHECK();
  latch_info_kernel<<<1, 1, 0, ctx->stream>>>(ch->d_info, fl->latch, code);
  GPS_LAUNCH_CHECK();
  ctx->launches += 3 + ch->launches;
  ch->launches = 0;
  return GPS_OK;
}

// Abar = 1/2 L^-T (Phi(L' Lbar) + Phi(L' Lbar)') L^-1 with Lbar = -tril(L^-T Sx) (+ diag term)
int chol_adjoint(gps_ctx* ctx, gps_fitc_large* fl, const double* L, const double* Linv, const double* Sx,
                 int add_diag, double* out) {
  const int Mp = fl->Mp;
  const unsigned nb = blocks_for((int64_t)Mp * Mp);
  double* sm = fl->sm.p;
  const size_t MM = (size_t)Mp * Mp;
  double *Y = sm + SM_Y * MM, *Z = sm + SM_Z * MM;
  GPS_CHECK(mm_gemm(ctx, fl, GEMM_MC_MC, Linv, Sx, Y, 1.0));        // L^-T Sx
  lbar_kernel<<<nb, 256, 0, ctx->stream>>>(Y, L, Mp, fl->M, add_diag, Z);
  GPS_LAUNCH_CHECK();
  GPS_CHECK(mm_gemm(ctx, fl, GEMM_MC_MC, L, Z, Y, 1.0));            // L' Lbar
  phi_sym_kernel<<<nb, 256, 0, ctx->stream>>>(Y, Mp, Z);
  GPS_LAUNCH_CHECK();
  GPS_CHECK(mm_gemm(ctx, fl, GEMM_MC_MC, Linv, Z, Y, 1.0));         // L^-T Z
  GPS_CHECK(mm_gemm(ctx, fl, GEMM_KC_MC, Y, Linv, out, 0.5));       // 1/2 (L^-T Z) L^-1
  ctx->launches += 2;
  return GPS_OK;
}

template <int MT, int DMX, bool DPOW2>
int kgrad(gps_ctx* ctx, gps_fitc_large* fl, const double* Kbar, const double* Kmat, int64_t ld, int64_t n,
          const double* P, double* out) {
  const int M = fl->M, D = fl->D;
  const int gx = (M + MT - 1) / MT;
  // column shares: the count that fills whole waves of 2 resident blocks per SM best (2..8 waves)
  const int64_t slots = 2 * (int64_t)ctx->sm_count, cols = (n + 255) / 256;
  int64_t gy = 1;
  double best = -1.0;
  for (int w = 2; w <= 8; ++w) {
    const int64_t c = std::max<int64_t>(1, w * slots / gx);
    const int64_t waves = (gx * c + slots - 1) / slots;
    const double eff = (double)(gx * c) / (double)(waves * slots);
    if (eff > best + 1e-9) { best = eff; gy = c; }
  }
  if (gy > cols) gy = cols;
  constexpr int NV = MT * (1 + DMX) + DMX;
  GPS_CHECK(gps_ensure(ctx, fl->part, (size_t)gx * gy * NV));
  kgrad_kernel<MT, DMX, DPOW2><<<dim3(gx, (unsigned)gy), 256, 0, ctx->stream>>>(Kbar, Kmat, ld, M, n, fl->U.p, P, D, ctx->params.p,
                                                                    fl->part.p);
  GPS_LAUNCH_CHECK();
  const int nthreads = 1 + DMAX + M * D;
  kgrad_reduce_kernel<MT, DMX><<<blocks_for(nthreads), 256, 0, ctx->stream>>>(fl->part.p, gx, (int)gy, M, D, ctx->params.p, out);
  GPS_LAUNCH_CHECK();
  kgrad_globals_kernel<MT, DMX><<<1 + D, 256, 0, ctx->stream>>>(fl->part.p, gx, (int)gy, out);
  GPS_LAUNCH_CHECK();
  ctx->launches += 3;
  return GPS_OK;
}

// D = 8 (the kin40k shape) gets the compile-time tile indexing and four rows per block
int kgrad_any(gps_ctx* ctx, gps_fitc_large* fl, const double* Kbar, const double* Kmat, int64_t ld, int64_t n,
              const double* P, double* out) {
  if (fl->D == 8) return kgrad<4, 8, true>(ctx, fl, Kbar, Kmat, ld, n, P, out);
  if (fl->D < 8) return kgrad<2, 8, false>(ctx, fl, Kbar, Kmat, ld, n, P, out);
  return kgrad<1, 16, false>(ctx, fl, Kbar, Kmat, ld, n, P, out);
}

// ---- block objectives: host side -----------------------------------------------------------------------
// Row-sharded layout: this context holds the rows [row_off, row_off + N) of world_n; fold f is the global range
// [f nf, (f+1) nf), nf = world_n / 4, of which [lo, hi) (possibly empty) lives here.
struct FoldRange { int64_t lo, hi; };

FoldRange fold_range(const gps_ctx* ctx, int f) {
  const gps_fitc_large* fl = ctx->fl;
  const int64_t nf = ctx->fitc.world_n / 4;
  FoldRange r;
  r.lo = std::min<int64_t>(ctx->N, std::max<int64_t>(0, f * nf - fl->row_off));
  r.hi = std::min<int64_t>(ctx->N, std::max<int64_t>(0, (f + 1) * nf - fl->row_off));
  return r;
}

// split-K task lists of the four folds: fold f contracts over its 16-rounded column range, the operands'
// stray columns inside that range are zeroed by the masked k-scaling vector
int build_fold_tasks(gps_ctx* ctx, gps_fitc_large* fl) {
  int64_t key[10];
  for (int f = 0; f < 4; ++f) {
    const FoldRange r = fold_range(ctx, f);
    key[2 * f] = r.lo; key[2 * f + 1] = r.hi;
  }
  key[8] = fl->Mp; key[9] = fl->S;
  if (fl->ftasks && std::equal(key, key + 10, fl->fold_key)) return GPS_OK;
  const int mt = fl->Mp / GPS_TILE, Mp = fl->Mp;
  std::vector<GemmTask> h;
  for (int f = 0; f < 4; ++f) {
    const int64_t lo = key[2 * f], hi = key[2 * f + 1];
    fl->t_fold[f].off = h.size();
    fl->fold_S[f] = 0;
    if (hi > lo) {
      const int64_t lo_r = lo & ~(int64_t)15, hi_r = std::min<int64_t>(fl->Npp, (hi + 15) & ~(int64_t)15);
      const int64_t blocks = (hi_r - lo_r) / 16;
      int64_t Sf = std::max<int64_t>(1, std::min<int64_t>((fl->S + 3) / 4, (blocks + 7) / 8));
      const int64_t per = ((blocks + Sf - 1) / Sf) * 16;
      Sf = (hi_r - lo_r + per - 1) / per;
      fl->fold_S[f] = (int)Sf;
      for (int s = 0; s < (int)Sf; ++s) {
        const int k0 = (int)(lo_r + s * per), k1 = (int)std::min<int64_t>(hi_r, lo_r + (s + 1) * per);
        for (int i = 0; i < mt; ++i)
          for (int j = 0; j <= i; ++j)
            h.push_back(make_task(i * GPS_TILE, j * GPS_TILE, k0, k1, s * Mp + i * GPS_TILE, j * GPS_TILE));
      }
    }
    fl->t_fold[f].cnt = h.size() - fl->t_fold[f].off;
  }
  if (h.size() > fl->ftasks_cap || !fl->ftasks) {
    if (fl->ftasks) cudaFree(fl->ftasks);
    fl->ftasks = nullptr;
    GPS_CUDA(cudaMalloc(&fl->ftasks, std::max<size_t>(1, h.size()) * sizeof(GemmTask)));
    fl->ftasks_cap = std::max<size_t>(1, h.size());
  }
  if (!h.empty()) {
    GPS_CUDA(cudaMemcpyAsync(fl->ftasks, h.data(), h.size() * sizeof(GemmTask), cudaMemcpyHostToDevice, ctx->stream));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  std::copy(key, key + 10, fl->fold_key);
  return GPS_OK;
}

// out (Mp x Mp, full symmetric) = W_f diag(src_f) W_f' over this context's columns [lo, hi) of fold f
int fold_splitk(gps_ctx* ctx, gps_fitc_large* fl, int f, FoldRange r, const double* src, double* out) {
  if (r.hi <= r.lo) {
    GPS_CUDA(cudaMemsetAsync(out, 0, (size_t)fl->Mp * fl->Mp * sizeof(double), ctx->stream));
    return GPS_OK;
  }
  const int64_t lo_r = r.lo & ~(int64_t)15, hi_r = std::min<int64_t>(fl->Npp, (r.hi + 15) & ~(int64_t)15);
  double* fv = fl->rv.p + RV_FV * fl->Npp;
  mask_range_kernel<<<blocks_for(hi_r - lo_r), 256, 0, ctx->stream>>>(src, r.lo, r.hi, lo_r, hi_r, fv);
  GPS_LAUNCH_CHECK();
  GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_KC, fl->W.p, fl->Npp, fl->W.p, fl->Npp, fl->part.p, fl->Mp, 1.0, 0.0, fv, false,
                           fl->ftasks + fl->t_fold[f].off, fl->t_fold[f].cnt));
  splitk_reduce_kernel<<<blocks_for((int64_t)fl->Mp * fl->Mp), 256, 0, ctx->stream>>>(fl->part.p, fl->fold_S[f], fl->Mp,
                                                                                      out, 1, 0);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  return GPS_OK;
}

// out[m] = sum over this context's columns of the fold of W[m][i] src[i]
int fold_rowdot(gps_ctx* ctx, gps_fitc_large* fl, FoldRange r, const double* src, double* out) {
  if (r.hi <= r.lo) {
    GPS_CUDA(cudaMemsetAsync(out, 0, (size_t)fl->Mp * sizeof(double), ctx->stream));
    return GPS_OK;
  }
  const int64_t lo_r = r.lo & ~(int64_t)15, hi_r = std::min<int64_t>(fl->Npp, (r.hi + 15) & ~(int64_t)15);
  double* fv = fl->rv.p + RV_FV * fl->Npp;
  mask_range_kernel<<<blocks_for(hi_r - lo_r), 256, 0, ctx->stream>>>(src, r.lo, r.hi, lo_r, hi_r, fv);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return rowdot(ctx, fl, fl->W.p + lo_r, fl->Npp, fl->Mp, hi_r - lo_r, fv + lo_r, out);
}

// T1[:, the fold's column tiles] = Mat (Mp x Mp) W
int fold_apply(gps_ctx* ctx, gps_fitc_large* fl, FoldRange r, const double* Mat) {
  const int mt = fl->Mp / GPS_TILE;
  const int64_t j0 = r.lo / GPS_TILE, j1 = (r.hi + GPS_TILE - 1) / GPS_TILE;
  return gps_gemm_tasks(ctx, GEMM_KC_MC, Mat, fl->Mp, fl->W.p, fl->Npp, fl->T1.p, fl->Npp, 1.0, 0.0, nullptr, false,
                        fl->tasks + fl->t_full.off + j0 * mt, (size_t)((j1 - j0) * mt));
}

// Pass 2 of the block objectives, after W / alpha / lambda.  Stages (each a loop over the four folds):
//   A  rows:        acc4[f] = [P_f = W_f Lam_f^-1 W_f' | g_f = W_f alpha_f] over this context's rows   -> all-reduce
//   B  replicated:  H_f = I - P_f = L_H L_H', H_f^-1, h_f = H_f^-1 g_f;  rows: T = (Hhat_f | H_f^-1) W_f, column sweep
//                   (fold predictive, objective share, seeds);  DSS: G_W, beta_bar (replicated)
//   C  rows (kc):   acc5[f] = [E_f = W_f diag(cbar) W_f' | hbar_f = -W_f mbar]                         -> all-reduce
//   D  (kc)         replicated: gbar_f, Hbar_f;  rows: T = Hbar_f W_f, second column sweep;  G_W, beta_bar
// and leaves the seeds (tbar, lam_bar0, D in T2) of the shared adjoint chain; acc2 = [G_W | beta_bar | obj].
// One GPU: allreduce == nullptr.
int block_pass2(gps_ctx* ctx, double* acc2, bool want_grad) {
  gps_fitc_large* fl = ctx->fl;
  auto& f = ctx->fitc;
  gps_ctx* ch = fl->ch;
  const int64_t N = ctx->N, Npp = fl->Npp, nf = f.world_n / 4;
  const int Mp = fl->Mp, M = fl->M;
  const size_t MM = (size_t)Mp * Mp, AF = MM + Mp;
  const bool dss = f.score == GPS_DSS;
  gps_allreduce_fn allreduce = fl->allreduce;
  cudaStream_t st = ctx->stream;
  double *sm = fl->sm.p, *rv = fl->rv.p;
  const double* alpha = f.rowv.p + 4 * N;
  const unsigned nbm = blocks_for((int64_t)MM), nbv = blocks_for(Mp);
  GPS_CHECK(build_fold_tasks(ctx, fl));
  GPS_CHECK(gps_ensure(ctx, fl->fs, 8));
  // fold buffer: acc4 [4][MM + Mp] | acc5 [4][MM + Mp] | H^-1 [4][MM] | h [4][Mp] | gbar [4][Mp]
  GPS_CHECK(gps_ensure(ctx, fl->fb, 8 * AF + 4 * MM + 8 * (size_t)Mp));
  double *A4 = fl->fb.p, *A5 = A4 + 4 * AF, *HI = A5 + 4 * AF, *HV = HI + 4 * MM, *GB = HV + 4 * (size_t)Mp;
  double *GW = acc2, *bbar = acc2 + MM, *d_obj = acc2 + MM + Mp;
  GPS_CUDA(cudaMemsetAsync(acc2, 0, (MM + Mp + 2) * sizeof(double), st));
  double *LH = sm + SM_LH * MM, *LHi = sm + SM_LHI * MM, *HX = sm + SM_HX * MM, *X2 = sm + SM_X2 * MM,
         *Y = sm + SM_Y * MM, *Z = sm + SM_Z * MM;
  double *lam = rv + RV_LAM * Npp, *il = rv + RV_IL * Npp, *abar = rv + RV_ABAR * Npp, *lbar = rv + RV_LBAR * Npp,
         *mbar = rv + RV_MBAR * Npp, *cbar = rv + RV_CBARV * Npp, *rowobj = rv + RV_ROWOBJ * Npp;
  FoldRange fr[4];
  for (int fo = 0; fo < 4; ++fo) fr[fo] = fold_range(ctx, fo);
  // ---- A
  for (int fo = 0; fo < 4; ++fo) {
    GPS_CHECK(fold_splitk(ctx, fl, fo, fr[fo], il, A4 + fo * AF));
    GPS_CHECK(fold_rowdot(ctx, fl, fr[fo], alpha, A4 + fo * AF + MM));
  }
  if (allreduce) GPS_CHECK(allreduce(ctx, A4, 4 * AF));
  // ---- B
  for (int fo = 0; fo < 4; ++fo) {
    const FoldRange r = fr[fo];
    const bool rows = r.hi > r.lo;
    const unsigned nbf = blocks_for(std::max<int64_t>(1, r.hi - r.lo));
    double *Pf = A4 + fo * AF, *g = Pf + MM, *Hinv = HI + fo * MM, *h = HV + fo * (size_t)Mp;
    h_from_p_kernel<<<nbm, 256, 0, st>>>(Pf, Mp, ch->Kb.p);
    GPS_LAUNCH_CHECK();
    GPS_CHECK(factor(ctx, fl, LH, LHi, 3 + fo));
    GPS_CHECK(mm_gemm(ctx, fl, GEMM_MC_MC, LHi, LHi, Hinv, 1.0));              // H^-1 = L_H^-T L_H^-1
    matvec_t_kernel<<<Mp / 16, 256, 0, st>>>(Hinv, Mp, g, h);                      // h = H^-1 g (H^-1 symmetric)
    GPS_LAUNCH_CHECK();
    fold_scalar_kernel<<<1, 256, 0, st>>>(LH, Mp, M, g, h, fl->fs.p + 2 * fo);
    GPS_LAUNCH_CHECK();
    ctx->launches += 3;
    if (dss && !want_grad) {                                                   // the rows' share needs no product
      if (rows) {
        col_dss_kernel<<<nbf, 256, 0, st>>>(fl->W.p, nullptr, fl->T2.p, Npp, M, r.lo, r.hi, h, lam, alpha, abar, lbar, rowobj);
        GPS_LAUNCH_CHECK();
        ctx->launches++;
      }
    } else if (dss) {
      hhat_kernel<<<nbm, 256, 0, st>>>(Hinv, h, Mp, M, HX);
      GPS_LAUNCH_CHECK();
      if (rows) {
        GPS_CHECK(fold_apply(ctx, fl, r, HX));                                 // T1 = Hhat W_f
        col_dss_kernel<<<nbf, 256, 0, st>>>(fl->W.p, fl->T1.p, fl->T2.p, Npp, M, r.lo, r.hi, h, lam, alpha, abar, lbar, rowobj);
        GPS_LAUNCH_CHECK();
      }
      GPS_CHECK(mm_gemm(ctx, fl, GEMM_KC_MC, HX, Pf, X2, -2.0));               // -2 Hhat P_f
      gw_dss_kernel<<<nbm, 256, 0, st>>>(X2, h, g, Mp, GW, bbar);
      GPS_LAUNCH_CHECK();
      ctx->launches += 3;
    } else if (rows) {
      GPS_CHECK(fold_apply(ctx, fl, r, Hinv));                                 // T1 = H^-1 W_f
      col_kca_kernel<<<nbf, 256, 0, st>>>(fl->W.p, fl->T1.p, fl->T2.p, Npp, M, r.lo, r.hi, h, lam, alpha, 1.0 / (double)nf,
                                          mbar, cbar, rowobj);
      GPS_LAUNCH_CHECK();
      ctx->launches++;
    }
  }
  // objective: the rows' shares add over the ranks, the fold terms are replicated and added once
  vec_sum_kernel<<<1, 1024, 0, st>>>(rowobj, N, d_obj);
  GPS_LAUNCH_CHECK();
  if (allreduce) GPS_CHECK(allreduce(ctx, d_obj, 2));
  block_obj_kernel<<<1, 1, 0, st>>>(d_obj, fl->fs.p, dss ? 1 : 0, (double)nf, d_obj);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  f.pass2_done = true;
  f.loo_ok = false;
  fl->ready = true;
  fl->block_seeds = want_grad;
  if (!want_grad) return GPS_OK;
  if (!dss) {
    // ---- C
    for (int fo = 0; fo < 4; ++fo) {
      double *E = A5 + fo * AF, *hbar = E + MM;
      GPS_CHECK(fold_rowdot(ctx, fl, fr[fo], mbar, hbar));                     // W_f mbar
      negate_kernel<<<nbv, 256, 0, st>>>(hbar, Mp);                            // hbar = -W_f mbar
      GPS_LAUNCH_CHECK();
      GPS_CHECK(fold_splitk(ctx, fl, fo, fr[fo], cbar, E));                    // E = W_f diag(cbar) W_f'
      ctx->launches++;
    }
    if (allreduce) GPS_CHECK(allreduce(ctx, A5, 4 * AF));
    // ---- D
    for (int fo = 0; fo < 4; ++fo) {
      const FoldRange r = fr[fo];
      const unsigned nbf = blocks_for(std::max<int64_t>(1, r.hi - r.lo));
      double *Pf = A4 + fo * AF, *g = Pf + MM, *Hinv = HI + fo * MM, *h = HV + fo * (size_t)Mp;
      double *E = A5 + fo * AF, *hbar = E + MM, *gbar = GB + fo * (size_t)Mp;
      matvec_t_kernel<<<Mp / 16, 256, 0, st>>>(Hinv, Mp, hbar, gbar);              // gbar = H^-1 hbar
      GPS_LAUNCH_CHECK();
      GPS_CHECK(mm_gemm(ctx, fl, GEMM_KC_MC, Hinv, E, Y, 1.0));                // Y = H^-1 E
      GPS_CHECK(mm_gemm(ctx, fl, GEMM_KC_MC, Y, Hinv, Z, 1.0));                // Z = H^-1 E H^-1
      hbar_kernel<<<nbm, 256, 0, st>>>(Z, h, gbar, Mp, M, HX);
      GPS_LAUNCH_CHECK();
      if (r.hi > r.lo) {
        GPS_CHECK(fold_apply(ctx, fl, r, HX));                                 // T1 = Hbar W_f
        col_kcb_kernel<<<nbf, 256, 0, st>>>(fl->W.p, fl->T1.p, fl->T2.p, Npp, M, r.lo, r.hi, gbar, lam, alpha, mbar, cbar,
                                            abar, lbar);
        GPS_LAUNCH_CHECK();
      }
      GPS_CHECK(mm_gemm(ctx, fl, GEMM_KC_MC, HX, Pf, X2, -2.0));               // -2 Hbar P_f
      gw_kc_kernel<<<nbm, 256, 0, st>>>(X2, Y, Pf, h, hbar, g, gbar, Mp, GW, bbar);
      GPS_LAUNCH_CHECK();
      ctx->launches += 4;
    }
  }
  block_seed_kernel<<<blocks_for(Npp), 256, 0, st>>>(N, Npp, il, alpha, abar, lbar, rv + RV_RBAR * Npp, rv + RV_TBAR * Npp);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

}  // namespace

void gps_fitc_large_free(gps_ctx* ctx) {
  gps_fitc_large* fl = ctx->fl;
  if (!fl) return;
  for (DevBuf* b : {&fl->Kuf, &fl->V, &fl->W, &fl->T1, &fl->T2, &fl->sm, &fl->rv, &fl->mv, &fl->part, &fl->out, &fl->U, &fl->dotp, &fl->acc, &fl->fs, &fl->fb})
    if (b->p) cudaFree(b->p);
  if (fl->tasks) cudaFree(fl->tasks);
  if (fl->ftasks) cudaFree(fl->ftasks);
  if (fl->latch) cudaFree(fl->latch);
  if (fl->ch) gps_ctx_release(fl->ch);
  delete fl;
  ctx->fl = nullptr;
}

// ---- staged evaluation: begin / pass1 / pass2 / pass3 / finish ------------------------------------------
// The same protocol as the fused M <= 32 path (include/gpscore.h): each pass leaves a packed
// accumulator of this context's rows on the device; multi-GPU runs all-reduce it before the next
// pass, gps_fitc_large_eval just chains the passes on its own buffers.
//   acc1 = [C - I (Mp*Mp) | v_y (Mp)]
//   acc2 = [R (Mp*Mp) | beta_bar (Mp) | obj: the rows' share]
//   acc3 = [S (Mp*Mp) | kernel-gradient block of K_uf (1 + 16 + M*D) | sum lambda_bar]

int gps_fitc_large_acc_len(int M, int D, int64_t* len1, int64_t* len2, int64_t* len3) {
  const int64_t Mp = gps_pad(M), MM = Mp * Mp;
  // even lengths: packed back to back, every accumulator stays 16-byte aligned (S is a GEMM operand)
  if (len1) *len1 = MM + Mp;
  if (len2) *len2 = MM + Mp + 2;
  if (len3) *len3 = (MM + (1 + DMAX + (int64_t)M * D) + 2) & ~(int64_t)1;
  return GPS_OK;
}

int gps_fitc_large_begin(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                         int64_t world_n) {
  if (score < GPS_CRPS || score > GPS_KC) return gps_fail(ctx, GPS_EINVAL, "fitc: unknown score %d", score);
  const bool block = score == GPS_DSS || score == GPS_KC;
  if (block && (world_n > 0 ? world_n : ctx->N) % 4 != 0)
    return gps_fail(ctx, GPS_EINVAL, "fitc dss/kc: the four folds need N %% 4 == 0 (K20:541-543)");
  if (ctx->D > DMAX) return gps_fail(ctx, GPS_EINVAL, "fitc: D=%d > %d not supported", ctx->D, DMAX);
  if (M > 4096) return gps_fail(ctx, GPS_EINVAL, "fitc: M=%d > 4096 not supported", M);
  GPS_CUDA(cudaSetDevice(ctx->device));
  auto& f = ctx->fitc;
  f.begun = false; f.pass2_done = false; f.loo_ok = false; f.large = true; f.fused = false;
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  if (!ctx->d_info) GPS_CUDA(cudaMalloc(&ctx->d_info, sizeof(int)));
  GPS_CHECK(setup(ctx, M, block));
  gps_fitc_large* fl = ctx->fl;
  fl->ready = false;
  fl->allreduce = nullptr;
  fl->row_off = 0;
  if (!fl->latch) GPS_CUDA(cudaMalloc(&fl->latch, 2 * sizeof(int)));
  GPS_CUDA(cudaMemsetAsync(fl->latch, 0, 2 * sizeof(int), ctx->stream));
  gps_ctx* ch = fl->ch;
  const int D = ctx->D, Mp = fl->Mp;
  const size_t MM = (size_t)Mp * Mp;
  cudaStream_t st = ctx->stream;
  double ea = 0, sn2 = 0;
  GPS_CHECK(gps_upload_params(ctx, theta, D, &ea, &sn2));
  f.M = M; f.MP = Mp; f.score = score; f.jitter = jitter; f.ea = ea; f.sn2 = sn2;
  f.world_n = world_n > 0 ? world_n : ctx->N;
  GPS_CUDA(cudaMemsetAsync(fl->U.p, 0, (size_t)Mp * D * sizeof(double), st));
  GPS_CUDA(cudaMemcpyAsync(fl->U.p, U, (size_t)M * D * sizeof(double), cudaMemcpyDefault, st));
  double* sm = fl->sm.p;
  // A = Kuu + jitter I = L_A L_A' (replicated)
  GPS_CUDA(cudaMemsetAsync(ch->Kb.p, 0, MM * sizeof(double), st));
  GPS_CHECK(gps_gram_rect(ctx, fl->U.p, M, fl->U.p, M, D, ctx->params.p, ch->Kb.p, Mp));
  kuu_fix_kernel<<<blocks_for((int64_t)MM), 256, 0, st>>>(ch->Kb.p, Mp, M, jitter, sm + SM_KUU * MM);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  GPS_CHECK(factor(ctx, fl, sm + SM_LA * MM, sm + SM_LAI * MM, 1));
  f.begun = true;
  return GPS_OK;
}

// pass 1: Kuf, V, lambda; acc1 = [V diag(1/lam) V' | V (y/lam)] over this context's rows
int gps_fitc_large_pass1(gps_ctx* ctx, double* acc1) {
  gps_fitc_large* fl = ctx->fl;
  const int64_t N = ctx->N, Npp = fl->Npp;
  const int D = ctx->D, Mp = fl->Mp, M = fl->M;
  const size_t MM = (size_t)Mp * Mp;
  cudaStream_t st = ctx->stream;
  const double* par = ctx->params.p;
  double *sm = fl->sm.p, *rv = fl->rv.p;
  GPS_CHECK(gps_gram_rect(ctx, fl->U.p, M, ctx->X.p, N, D, par, fl->Kuf.p, Npp));
  GPS_CHECK(big_gemm(ctx, fl, GEMM_KC_MC, fl->t_low, sm + SM_LAI * MM, fl->Kuf.p, fl->V.p));
  col_lambda_kernel<<<blocks_for(Npp), 256, 0, st>>>(fl->V.p, Npp, M, N, Npp, ctx->y.p, par, rv + RV_LAM * Npp,
                                                     rv + RV_IL * Npp, rv + RV_YL * Npp);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  GPS_CHECK(rowdot(ctx, fl, fl->V.p, Npp, Mp, Npp, rv + RV_YL * Npp, acc1 + MM));
  GPS_CHECK(splitk(ctx, fl, fl->V.p, fl->V.p, rv + RV_IL * Npp, true, false, acc1));
  return GPS_OK;
}

// pass 2: C = I + acc1 -> L_C, beta; W, d, alpha, score + seeds; acc2 = [W diag(rbar) W' | W tbar | obj]
int gps_fitc_large_pass2(gps_ctx* ctx, const double* acc1, double* acc2, bool want_grad) {
  gps_fitc_large* fl = ctx->fl;
  auto& f = ctx->fitc;
  gps_ctx* ch = fl->ch;
  const int64_t N = ctx->N, Npp = fl->Npp;
  const int Mp = fl->Mp, M = fl->M;
  const size_t MM = (size_t)Mp * Mp;
  const bool nlml = f.score == GPS_NLML;
  cudaStream_t st = ctx->stream;
  double *sm = fl->sm.p, *rv = fl->rv.p, *mv = fl->mv.p;
  double* alpha = f.rowv.p + 4 * N;
  double* dd = f.rowv.p + 5 * N;
  const unsigned nbn = blocks_for(Npp);
  add_identity_kernel<<<blocks_for((int64_t)MM), 256, 0, st>>>(acc1, Mp, ch->Kb.p);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  GPS_CHECK(factor(ctx, fl, sm + SM_LC * MM, sm + SM_LCI * MM, 2));
  GPS_CHECK(rowdot(ctx, fl, sm + SM_LCI * MM, Mp, Mp, Mp, acc1 + MM, mv + MV_BETA * Mp));
  GPS_CHECK(big_gemm(ctx, fl, GEMM_KC_MC, fl->t_low, sm + SM_LCI * MM, fl->V.p, fl->W.p));
  col_w_kernel<<<nbn, 256, 0, st>>>(fl->W.p, Npp, M, N, Npp, mv + MV_BETA * Mp, ctx->y.p, rv + RV_IL * Npp,
                                    rv + RV_R * Npp, alpha, dd);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  if (f.score == GPS_DSS || f.score == GPS_KC) return block_pass2(ctx, acc2, want_grad);
  fl->block_seeds = false;
  double* d_obj = acc2 + MM + Mp;
  if (nlml) {
    nlml_rows_kernel<<<1, 1024, 0, st>>>(N, Npp, rv + RV_LAM * Npp, ctx->y.p, alpha, rv + RV_ABAR * Npp,
                                         rv + RV_DBAR * Npp, d_obj);
    GPS_LAUNCH_CHECK();
    logdiag_sum_kernel<<<1, 256, 0, st>>>(sm + SM_LC * MM, Mp, M, fl->out.p + OUT_LOGDET);
    GPS_LAUNCH_CHECK();
    ctx->launches += 2;
  } else {
    // alpha / d are N long (the layout gps_fitc_loo reads); the score kernel pads its outputs to Npp
    // and divides by the global row count, so the ranks' shares add up to the mean of KF:67
    GPS_CHECK(gps_loo_score(ctx, f.score, N, Npp, f.world_n, alpha, dd, ctx->y.p, rv + RV_ABAR * Npp,
                            rv + RV_DBAR * Npp, rv + RV_LOOM * Npp, rv + RV_LOOV * Npp, d_obj));
  }
  f.pass2_done = true;
  f.loo_ok = true;
  fl->ready = true;
  if (!want_grad) return GPS_OK;
  seed_kernel<<<nbn, 256, 0, st>>>(N, Npp, nlml ? 1 : 0, rv + RV_IL * Npp, rv + RV_R * Npp, alpha, rv + RV_ABAR * Npp,
                                   rv + RV_DBAR * Npp, rv + RV_LBAR * Npp, rv + RV_RBAR * Npp, rv + RV_TBAR * Npp);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  GPS_CHECK(rowdot(ctx, fl, fl->W.p, Npp, Mp, Npp, rv + RV_TBAR * Npp, acc2 + MM));
  GPS_CHECK(splitk(ctx, fl, fl->W.p, fl->W.p, rv + RV_RBAR * Npp, true, false, acc2));
  return GPS_OK;
}

// pass 3: C_bar, vy_bar (replicated); lambda_bar, V_bar, Kuf_bar; acc3 = [V_bar V' | kernel gradient | sum lam_bar]
int gps_fitc_large_pass3(gps_ctx* ctx, const double* acc2, double* acc3) {
  gps_fitc_large* fl = ctx->fl;
  auto& f = ctx->fitc;
  const int64_t N = ctx->N, Npp = fl->Npp;
  const int D = ctx->D, Mp = fl->Mp, M = fl->M;
  const size_t MM = (size_t)Mp * Mp;
  const bool nlml = f.score == GPS_NLML;
  cudaStream_t st = ctx->stream;
  double *sm = fl->sm.p, *rv = fl->rv.p, *mv = fl->mv.p;
  const double* bbar = acc2 + MM;
  const size_t glen = 1 + DMAX + (size_t)M * D;
  const bool block = fl->block_seeds;
  sw_kernel<<<blocks_for((int64_t)MM), 256, 0, st>>>(mv + MV_BETA * Mp, bbar, acc2, Mp, block ? 1.0 : 2.0, sm + SM_SW * MM);
  GPS_LAUNCH_CHECK();
  GPS_CHECK(chol_adjoint(ctx, fl, sm + SM_LC * MM, sm + SM_LCI * MM, sm + SM_SW * MM, nlml ? 1 : 0, sm + SM_CBAR * MM));
  matvec_t_kernel<<<Mp / 16, 256, 0, st>>>(sm + SM_LCI * MM, Mp, bbar, mv + MV_VYBAR * Mp);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  GPS_CHECK(big_gemm(ctx, fl, GEMM_KC_MC, fl->t_full, sm + SM_CBAR * MM, fl->V.p, fl->T1.p));     // CV
  col_pass3_kernel<<<blocks_for(N), 256, 0, st>>>(fl->V.p, fl->T1.p, fl->W.p, Npp, M, N, bbar, mv + MV_BETA * Mp,
                                                  ctx->y.p, rv + RV_IL * Npp, rv + RV_TBAR * Npp, rv + RV_RBAR * Npp,
                                                  rv + RV_LBAR * Npp, block ? fl->T2.p : nullptr);
  GPS_LAUNCH_CHECK();
  GPS_CHECK(big_gemm(ctx, fl, GEMM_MC_MC, fl->t_up, sm + SM_LCI * MM, fl->W.p, fl->T2.p));        // L_C^-T Wbar
  vbar_kernel<<<dim3(blocks_for(N), M), 256, 0, st>>>(fl->T2.p, fl->T1.p, fl->V.p, Npp, M, N, mv + MV_VYBAR * Mp,
                                                      rv + RV_YL * Npp, rv + RV_IL * Npp, rv + RV_LBAR * Npp);
  GPS_LAUNCH_CHECK();
  GPS_CHECK(big_gemm(ctx, fl, GEMM_MC_MC, fl->t_up, sm + SM_LAI * MM, fl->T2.p, fl->T1.p));       // Kuf_bar
  GPS_CHECK(splitk(ctx, fl, fl->T2.p, fl->V.p, nullptr, false, false, acc3));
  vec_sum_kernel<<<1, 1024, 0, st>>>(rv + RV_LBAR * Npp, N, acc3 + MM + glen);
  GPS_LAUNCH_CHECK();
  ctx->launches += 3;
  GPS_CHECK(kgrad_any(ctx, fl, fl->T1.p, fl->Kuf.p, Npp, N, ctx->X.p, acc3 + MM));
  return GPS_OK;
}

// finish (replicated): A_bar through the Cholesky adjoint of L_A, Kuu gradient, assembly on the host
int gps_fitc_large_finish(gps_ctx* ctx, const double* acc2, const double* acc3, double* obj, double* grad_theta,
                          double* grad_U) {
  gps_fitc_large* fl = ctx->fl;
  auto& f = ctx->fitc;
  const int D = ctx->D, Mp = fl->Mp, M = fl->M;
  const size_t MM = (size_t)Mp * Mp;
  const size_t glen = 1 + DMAX + (size_t)M * D;
  cudaStream_t st = ctx->stream;
  double* sm = fl->sm.p;
  double* d_out = fl->out.p;
  const bool want_grad = grad_theta || grad_U;
  std::vector<double>& h = fl->h_out;
  h.assign(OUT_G1 + 2 * glen + 2, 0.0);
  GPS_CUDA(cudaMemcpyAsync(&h[OUT_OBJ], acc2 + MM + Mp, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (f.score == GPS_NLML)
    GPS_CUDA(cudaMemcpyAsync(&h[OUT_LOGDET], d_out + OUT_LOGDET, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (want_grad) {
    GPS_CHECK(chol_adjoint(ctx, fl, sm + SM_LA * MM, sm + SM_LAI * MM, acc3, 0, sm + SM_ABAR * MM));
    GPS_CHECK(kgrad_any(ctx, fl, sm + SM_ABAR * MM, sm + SM_KUU * MM, Mp, M, fl->U.p, d_out + OUT_G1 + glen));
    GPS_CUDA(cudaMemcpyAsync(&h[OUT_G1], acc3 + MM, (glen + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    GPS_CUDA(cudaMemcpyAsync(&h[OUT_G1 + glen + 1], d_out + OUT_G1 + glen, glen * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  int latch[2] = {0, 0};
  GPS_CUDA(cudaMemcpyAsync(latch, fl->latch, sizeof latch, cudaMemcpyDeviceToHost, st));
  GPS_CUDA(cudaStreamSynchronize(st));
  if (latch[0] != 0) {
    fl->ready = false;                      // no prediction / LOO read-back from a failed evaluation
    f.loo_ok = false;
    static const char* const what[3] = {"K_uu + jitter I", "I + V Lambda^-1 V'", "I - W_f Lambda_f^-1 W_f' (a fold of the block objective)"};
    return gps_fail(ctx, GPS_ENOTPD, "fitc: %s not positive definite at pivot %d", what[latch[0] > 3 ? 2 : latch[0] - 1], latch[1]);
  }
  double value = h[OUT_OBJ];
  if (f.score == GPS_NLML) value += (double)f.world_n * HALF_LOG_2PI + h[OUT_LOGDET];   // replicated terms, added once
  if (obj) *obj = value;
  if (!want_grad) return GPS_OK;
  const double* g1 = &h[OUT_G1];
  const double sum_lbar = g1[glen];
  const double* g2 = g1 + glen + 1;
  if (grad_theta) {
    grad_theta[0] = g1[0] + f.ea * sum_lbar + g2[0];
    for (int d = 0; d < D; ++d) grad_theta[1 + d] = g1[1 + d] + g2[1 + d];
    grad_theta[D + 1] = f.sn2 * sum_lbar;
  }
  if (grad_U)
    for (int e = 0; e < M * D; ++e) grad_U[e] = g1[1 + DMAX + e] + 2.0 * g2[1 + DMAX + e];
  return GPS_OK;
}

int gps_fitc_large_eval(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                        double* obj, double* grad_theta, double* grad_U) {
  GPS_CHECK(gps_fitc_large_begin(ctx, theta, U, M, jitter, score, ctx->N));
  gps_fitc_large* fl = ctx->fl;
  int64_t l1, l2, l3;
  gps_fitc_large_acc_len(M, ctx->D, &l1, &l2, &l3);
  GPS_CHECK(gps_ensure(ctx, fl->acc, (size_t)(l1 + l2 + l3)));
  double *a1 = fl->acc.p, *a2 = a1 + l1, *a3 = a2 + l2;
  const bool want_grad = grad_theta || grad_U;
  GPS_CHECK(gps_fitc_large_pass1(ctx, a1));
  GPS_CHECK(gps_fitc_large_pass2(ctx, a1, a2, want_grad));
  if (want_grad) GPS_CHECK(gps_fitc_large_pass3(ctx, a2, a3));
  return gps_fitc_large_finish(ctx, a2, a3, obj, grad_theta, grad_U);
}

// row-sharded evaluation: the same chain with the packed accumulators summed over the ranks between the passes.
// row_offset = first global row of this context's block (only the block objectives' fold geometry needs it)
int gps_fitc_large_eval_sharded(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                                int64_t world_n, int64_t row_offset, gps_allreduce_fn allreduce, double* obj,
                                double* grad_theta, double* grad_U) {
  if (M < 1 || M > 4096) return gps_fail(ctx, GPS_EINVAL, "fitc_eval_sharded: M=%d outside 1..4096", M);
  const bool block = score == GPS_DSS || score == GPS_KC;
  if (row_offset < 0 || row_offset + ctx->N > world_n)
    return gps_fail(ctx, GPS_EINVAL, "fitc_eval_sharded: rows [%lld, %lld) outside the %lld global rows", (long long)row_offset,
                    (long long)(row_offset + ctx->N), (long long)world_n);
  GPS_CHECK(gps_fitc_large_begin(ctx, theta, U, M, jitter, score, world_n));
  gps_fitc_large* fl = ctx->fl;
  fl->allreduce = allreduce;
  fl->row_off = row_offset;
  int64_t l1, l2, l3;
  gps_fitc_large_acc_len(M, ctx->D, &l1, &l2, &l3);
  GPS_CHECK(gps_ensure(ctx, fl->acc, (size_t)(l1 + l2 + l3)));
  double *a1 = fl->acc.p, *a2 = a1 + l1, *a3 = a2 + l2;
  const bool want_grad = grad_theta || grad_U;
  int rc = gps_fitc_large_pass1(ctx, a1);
  if (rc == GPS_OK) rc = allreduce(ctx, a1, (size_t)l1);
  if (rc == GPS_OK) rc = gps_fitc_large_pass2(ctx, a1, a2, want_grad);   // block objectives: all-reduces inside, acc2 replicated
  if (rc == GPS_OK && !block) rc = allreduce(ctx, a2, (size_t)l2);
  if (rc == GPS_OK && want_grad) {
    rc = gps_fitc_large_pass3(ctx, a2, a3);
    if (rc == GPS_OK) rc = allreduce(ctx, a3, (size_t)l3);
  }
  fl->allreduce = nullptr;
  if (rc != GPS_OK) return rc;
  return gps_fitc_large_finish(ctx, a2, a3, obj, grad_theta, grad_U);
}

// prediction at the factors of the last evaluation, test rows in chunks
int gps_fitc_large_predict(gps_ctx* ctx, const double* dXs, int64_t T, double* dm, double* dv) {
  gps_fitc_large* fl = ctx->fl;
  if (!fl || !fl->ready) return gps_fail(ctx, GPS_ESTATE, "fitc_predict: no evaluation at this theta, U");
  const int Mp = fl->Mp, M = fl->M, D = fl->D;
  const size_t MM = (size_t)Mp * Mp;
  const int mt = Mp / GPS_TILE;
  // the evaluation's [Mp][Npp] scratch holds the chunk: Ks -> T1, Vs -> T2, Ws -> W
  const int64_t chunk = std::min<int64_t>(fl->Npp, (int64_t)1 << 16);
  const double* par = ctx->params.p;
  double* sm = fl->sm.p;
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t tc = std::min(chunk, T - t0), tp = gps_pad(tc);
    std::vector<GemmTask> tasks;
    for (int64_t j = 0; j < tp / GPS_TILE; ++j)
      for (int i = 0; i < mt; ++i)
        tasks.push_back(make_task(i * GPS_TILE, (int)(j * GPS_TILE), 0, (i + 1) * GPS_TILE, i * GPS_TILE, (int)(j * GPS_TILE),
                                  GEMM_TRI_END));
    GPS_CHECK(gps_upload_tasks2(ctx, tasks));
    GPS_CUDA(cudaMemsetAsync(fl->T1.p, 0, (size_t)Mp * tp * sizeof(double), ctx->stream));
    GPS_CHECK(gps_gram_rect(ctx, fl->U.p, M, dXs + t0 * D, tc, D, par, fl->T1.p, tp));
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, sm + SM_LAI * MM, Mp, fl->T1.p, tp, fl->T2.p, tp, 1.0, 0.0, nullptr, false,
                             ctx->d_tasks2, tasks.size()));
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, sm + SM_LCI * MM, Mp, fl->T2.p, tp, fl->W.p, tp, 1.0, 0.0, nullptr, false,
                             ctx->d_tasks2, tasks.size()));
    col_predict_kernel<<<blocks_for(tc), 256, 0, ctx->stream>>>(fl->T2.p, fl->W.p, tp, M, tc, fl->mv.p + MV_BETA * Mp, par,
                                                                dm + t0, dv + t0);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
  }
  return GPS_OK;
}
