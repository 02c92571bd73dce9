// Scoring-rule kernels: the LOO transform + closed-form Gaussian CRPS / log score with their
// adjoint seeds in one pass (replaces KF:243-245 + crps KF:60-68 / logs KF:52-57 and the part of
// .backward() that differentiates them), the N^2 gradient contraction that rebuilds K_f tiles
// from X instead of reading a stored Gram, symmetric mat-vec, and the
// test-set metric reductions (KF:276-292).  Reductions are warp-shuffle + fixed-order block
// sums: results are bit-reproducible run to run.
#include "gps_common.cuh"
#include "gps_exp.cuh"

namespace {

constexpr double INV_SQRT_PI = 0.56418958354775628695;    // 1/sqrt(pi)
constexpr double INV_SQRT_2PI = 0.39894228040143267794;   // 1/sqrt(2 pi)
constexpr double INV_SQRT2 = 0.70710678118654752440;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

// per-row CRPS of N(mu, s^2) at y given z = (y - mu)/s:  s * g(z)
__device__ __forceinline__ double crps_g(double z, double& two_phi_m1) {
  two_phi_m1 = erf(z * INV_SQRT2);  // 2 Phi(z) - 1
  return z * two_phi_m1 + 2.0 * INV_SQRT_2PI * exp(-0.5 * z * z) - INV_SQRT_PI;
}

// y = A x for symmetric full A [Np, Np]; one warp per row, 16-byte loads.
__global__ void __launch_bounds__(256)
symv_kernel(const double* __restrict__ A, int64_t Np, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Np) return;
  const double2* a = reinterpret_cast<const double2*>(A + row * Np);
  const double2* xv = reinterpret_cast<const double2*>(x);
  double s0 = 0.0, s1 = 0.0;
  for (int64_t j = lane; j < Np / 2; j += 32) {
    const double2 av = a[j];
    const double2 xx = xv[j];
    s0 = fma(av.x, xx.x, s0);
    s1 = fma(av.y, xx.y, s1);
  }
  const double s = warp_sum(s0 + s1);
  if (lane == 0) y[row] = s;
}

__global__ void diag_kernel(const double* __restrict__ A, int64_t Np, double* __restrict__ d, int do_log) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) {
    const double v = A[i * Np + i];
    d[i] = do_log ? log(v) : v;
  }
}

// LOO transform, score value and adjoint seeds.  One block writes the value itself; a larger grid
// (N > 32768) writes one partial per block, summed in block order by partial_sum_kernel:
// deterministic either way.
__global__ void __launch_bounds__(1024)
loo_score_kernel(int score, int64_t N, int64_t Np, int64_t norm_n, const double* __restrict__ alpha,
                 const double* __restrict__ dg, const double* __restrict__ y, double* __restrict__ abar,
                 double* __restrict__ dbar, double* __restrict__ loo_mean, double* __restrict__ loo_var,
                 double* __restrict__ obj) {
  __shared__ double sh[32];
  const double invN = 1.0 / (double)norm_n;   // the mean runs over all rows of the data set (KF:67)
  double sum = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += (int64_t)blockDim.x * gridDim.x) {
    if (i >= N) {
      abar[i] = 0.0;
      dbar[i] = 0.0;
      loo_mean[i] = 0.0;
      loo_var[i] = 1.0;
      continue;
    }
    const double a = alpha[i], d = dg[i];
    const double s2 = 1.0 / d;
    loo_mean[i] = y[i] - a * s2;   // KF:243
    loo_var[i] = s2;               // KF:244
    if (score == GPS_CRPS) {
      const double s = sqrt(s2);
      const double z = a * s;      // (y - mu)/s = (a/d) * sqrt(d)
      double tpm1;
      const double g = crps_g(z, tpm1);
      sum += s * g;
      abar[i] = tpm1 * s2 * invN;
      dbar[i] = -(0.5 * s2 * s * g + 0.5 * tpm1 * a * s2 * s2) * invN;
    } else {
      sum += 0.5 * a * a * s2 - 0.5 * log(d) + HALF_LOG_2PI;
      abar[i] = a * s2 * invN;
      dbar[i] = -(0.5 * a * a * s2 * s2 + 0.5 * s2) * invN;
    }
  }
  sum = block_sum(sum, sh);
  if (threadIdx.x == 0) obj[blockIdx.x] = gridDim.x == 1 ? sum * invN : sum;
}

__global__ void __launch_bounds__(256)
partial_sum_kernel(const double* __restrict__ part, int n, double scale, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s * scale;
}

__global__ void __launch_bounds__(1024)
nlml_value_kernel(int64_t N, const double* __restrict__ logdiag, const double* __restrict__ alpha,
                  const double* __restrict__ y, double* __restrict__ obj) {
  __shared__ double sh[32];
  double sum = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) sum += logdiag[i] + 0.5 * y[i] * alpha[i];
  sum = block_sum(sum, sh);
  if (threadIdx.x == 0) obj[0] = sum + (double)N * HALF_LOG_2PI;   // KF:334
}

// Gradient contraction over the lower 128 x 128 tiles of W:
//   mode 0 (LOO scores): W_ij = -( (u_i a_j + a_i u_j)/2 + S_ij ),  S = K^-1 diag(dbar) K^-1
//   mode 1 (NLML):       W_ij = ( Kinv_ij - a_i a_j ) / 2
// partial[tile][0] = sum W K_f, [1 + d] = sum W K_f ((x_id - x_jd)/l_d)^2, [1 + D] = trace part.
// Off-diagonal tiles count twice (symmetry).
__global__ void __launch_bounds__(256, 2)
grad_contract_kernel(int mode, const double* __restrict__ Mx, int64_t N, int64_t Np,
                     const double* __restrict__ X, int D, const double* __restrict__ par,
                     const double* __restrict__ alpha, const double* __restrict__ u,
                     double* __restrict__ partial) {
  extern __shared__ double sh[];
  constexpr int TS = GPS_TILE;
  double* xi = sh;                        // [D][128]
  double* xj = xi + (size_t)D * TS;       // [D][128]
  double* ai = xj + (size_t)D * TS;       // alpha_i, u_i, alpha_j, u_j : 4 x 128
  double* red = ai + 4 * TS;              // [8 warps][D + 2]
  int bi = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((int64_t)(bi + 1) * (bi + 2) / 2 <= (int64_t)blockIdx.x) ++bi;
  while ((int64_t)bi * (bi + 1) / 2 > (int64_t)blockIdx.x) --bi;
  const int bj = (int)(blockIdx.x - (int64_t)bi * (bi + 1) / 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < TS * D; e += 256) {
    const int r = e / D, d = e - r * D;
    const double il = par[2 + d];
    xi[d * TS + r] = X[((int64_t)bi * TS + r) * D + d] * il;
    xj[d * TS + r] = X[((int64_t)bj * TS + r) * D + d] * il;
  }
  if (tid < TS) {
    ai[tid] = alpha[(int64_t)bi * TS + tid];
    ai[2 * TS + tid] = alpha[(int64_t)bj * TS + tid];
    if (mode == 0) {
      ai[TS + tid] = u[(int64_t)bi * TS + tid];
      ai[3 * TS + tid] = u[(int64_t)bj * TS + tid];
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  const double ea = par[0];
  const double wgt = (bi == bj) ? 1.0 : 2.0;
  const int nred = D + 2;
  // The thread's 8 x 8 block is processed as two 8 x 4 column halves (two CTAs per SM instead of one spilling at 255
  // registers).  Per input dimension
  //   sum_ij G_ij (x_id - x_jd)^2 = sum_i x_id^2 rowsum_i(G) + sum_j x_jd^2 colsum_j(G) - 2 sum_i x_id (G x_d)_i
  // (32 + 12 fmas per half instead of 3 x 32; the relative cancellation of the expansion is ~1e-15 |G| x^2, far inside
  // the 1e-6 gradient tolerance — the objective does not pass through here).
  double rs[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) rs[r] = 0.0;
  double s_a = 0.0, s_tr = 0.0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    double G[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) G[r][c] = 0.0;
    for (int d = 0; d < D; ++d) {
      double a[8], b[4];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = xi[d * TS + ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = xj[d * TS + tx + 16 * (4 * h + c)];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double df = a[r] - b[c];
          G[r][c] = fma(df, df, G[r][c]);
        }
    }
    double cs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int il = ty + 16 * r;
      const int64_t i = (int64_t)bi * TS + il;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jl = tx + 16 * (4 * h + c);
        const int64_t j = (int64_t)bj * TS + jl;
        const double m = Mx[i * Np + j];
        double w;
        if (mode == 0)
          w = -(0.5 * (ai[TS + il] * ai[2 * TS + jl] + ai[il] * ai[3 * TS + jl]) + m);
        else
          w = 0.5 * (m - ai[il] * ai[2 * TS + jl]);
        if (i >= N || j >= N) w = 0.0;
        if (i == j) s_tr += w;
        const double g = wgt * w * ea * exp_neg(-0.5 * G[r][c]);
        G[r][c] = g;
        s_a += g;
        rs[r] += g;
        cs[c] += g;
      }
    }
    for (int d = 0; d < D; ++d) {
      double b[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = xj[d * TS + tx + 16 * (4 * h + c)];
      double s = 0.0, cross = 0.0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const double t0 = fma(G[r][0], b[0], G[r][1] * b[1]), t1 = fma(G[r][2], b[2], G[r][3] * b[3]);
        cross = fma(xi[d * TS + ty + 16 * r], t0 + t1, cross);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) s = fma(b[c] * b[c], cs[c], s);
      s = fma(-2.0, cross, s);
      s = warp_sum(s);
      if (lane == 0) {
        if (h == 0) red[warp * nred + 1 + d] = s;
        else red[warp * nred + 1 + d] += s;
      }
    }
  }
  s_a = warp_sum(s_a);
  s_tr = warp_sum(s_tr);
  if (lane == 0) {
    red[warp * nred] = s_a;
    red[warp * nred + 1 + D] = s_tr;
  }
  for (int d = 0; d < D; ++d) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const double a = xi[d * TS + ty + 16 * r];
      s = fma(a * a, rs[r], s);
    }
    s = warp_sum(s);
    if (lane == 0) red[warp * nred + 1 + d] += s;
  }
  __syncthreads();
  if (tid < nred) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * nred + tid];
    partial[(int64_t)blockIdx.x * nred + tid] = s;
  }
}

// out[c] = sum_t partial[t][c] ; one block per column, fixed order.
__global__ void __launch_bounds__(256)
colsum_kernel(const double* __restrict__ partial, int64_t rows, int ncol, double* __restrict__ out) {
  __shared__ double sh[32];
  const int c = blockIdx.x;
  double s = 0.0;
  for (int64_t t = threadIdx.x; t < rows; t += blockDim.x) s += partial[t * ncol + c];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[c] = s;
}

// test-set reductions (KF:276-292): sums[0] = sum (m-y)^2, [1] = sum (ybar_train - y)^2,
// [2] = sum log-score terms, [3] = sum CRPS terms, [4] = sum trivial-model terms, [5] = #inside +-2sd
__global__ void __launch_bounds__(1024)
metrics_kernel(const double* __restrict__ mean, const double* __restrict__ var, const double* __restrict__ y,
               int64_t n, double ytm, double ytv, double* __restrict__ sums) {
  __shared__ double sh[32];
  double s[6] = {0, 0, 0, 0, 0, 0};
  const double triv0 = 0.5 * log(2.0 * 3.14159265358979323846 * ytv);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = mean[i], c = var[i], yy = y[i];
    const double e = yy - m;
    s[0] += e * e;
    s[1] += (ytm - yy) * (ytm - yy);
    s[2] += e * e / (2.0 * c) + 0.5 * log(c) + HALF_LOG_2PI;
    const double sd = sqrt(c);
    double tpm1;
    const double g = crps_g(e / sd, tpm1);
    s[3] += sd * g;
    s[4] += triv0 + (yy - ytm) * (yy - ytm) / (2.0 * ytv);
    s[5] += ((m + 2.0 * sd - yy) > 0.0 && (yy - (m - 2.0 * sd)) > 0.0) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double v = block_sum(s[k], sh);
    if (threadIdx.x == 0) sums[k] = v;
  }
}

__global__ void __launch_bounds__(1024)
score_kernel(const double* __restrict__ mean, const double* __restrict__ var, const double* __restrict__ y,
             int64_t n, int which, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = mean[i], c = var[i], e = y[i] - m;
    if (which == GPS_CRPS) {
      const double sd = sqrt(c);
      double tpm1;
      s += sd * crps_g(e / sd, tpm1);
    } else {
      s += e * e / (2.0 * c) + 0.5 * log(c) + HALF_LOG_2PI;
    }
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s / (double)n;
}

}  // namespace

int gps_symv(gps_ctx* ctx, const double* A, int64_t Np, const double* x, double* y) {
  symv_kernel<<<(unsigned)((Np + 7) / 8), 256, 0, ctx->stream>>>(A, Np, x, y);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

int gps_diag_extract(gps_ctx* ctx, const double* A, int64_t Np, double* d, int do_log) {
  diag_kernel<<<(unsigned)((Np + 255) / 256), 256, 0, ctx->stream>>>(A, Np, d, do_log);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

int gps_loo_score(gps_ctx* ctx, int score, int64_t N, int64_t Np, int64_t norm_n, const double* alpha, const double* d,
                  const double* y, double* abar, double* dbar, double* loo_mean, double* loo_var,
                  double* obj_dev) {
  if (N <= 32768) {
    loo_score_kernel<<<1, 1024, 0, ctx->stream>>>(score, N, Np, norm_n, alpha, d, y, abar, dbar, loo_mean, loo_var,
                                                  obj_dev);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
    return GPS_OK;
  }
  const int grid = (int)std::min<int64_t>((Np + 1023) / 1024, (int64_t)2 * ctx->sm_count);
  GPS_CHECK(gps_ensure(ctx, ctx->red, (size_t)grid));
  loo_score_kernel<<<grid, 1024, 0, ctx->stream>>>(score, N, Np, norm_n, alpha, d, y, abar, dbar, loo_mean, loo_var,
                                                   ctx->red.p);
  GPS_LAUNCH_CHECK();
  partial_sum_kernel<<<1, 256, 0, ctx->stream>>>(ctx->red.p, grid, 1.0 / (double)norm_n, obj_dev);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  return GPS_OK;
}

int gps_nlml_value(gps_ctx* ctx, int64_t N, int64_t Np, const double* logdiag, const double* alpha,
                   const double* y, double* obj_dev) {
  (void)Np;
  nlml_value_kernel<<<1, 1024, 0, ctx->stream>>>(N, logdiag, alpha, y, obj_dev);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

int gps_grad_contract(gps_ctx* ctx, int mode, const double* Mx, int64_t N, int64_t Np, const double* X, int D,
                      const double* d_par, const double* alpha, const double* u, double* out_dev) {
  const int64_t nb = Np / GPS_TILE;
  const int64_t tiles = nb * (nb + 1) / 2;
  const int nred = D + 2;
  GPS_CHECK(gps_ensure(ctx, ctx->red, (size_t)tiles * nred));
  const size_t smem = ((size_t)2 * D * GPS_TILE + 4 * GPS_TILE + 8 * nred) * sizeof(double);
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(grad_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(((size_t)2 * 64 * GPS_TILE + 4 * GPS_TILE + 8 * 66) * sizeof(double))));
    configured = true;
  }
  grad_contract_kernel<<<(unsigned)tiles, 256, smem, ctx->stream>>>(mode, Mx, N, Np, X, D, d_par, alpha, u,
                                                                   ctx->red.p);
  GPS_LAUNCH_CHECK();
  colsum_kernel<<<nred, 256, 0, ctx->stream>>>(ctx->red.p, tiles, nred, out_dev);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  return GPS_OK;
}

int gps_metrics_kernel(gps_ctx* ctx, const double* mean, const double* var, const double* y, int64_t n,
                       double ytm, double ytv, double* out_dev) {
  metrics_kernel<<<1, 1024, 0, ctx->stream>>>(mean, var, y, n, ytm, ytv, out_dev);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

int gps_score_kernel(gps_ctx* ctx, const double* m, const double* c, const double* y, int64_t n, int which,
                     double* out_dev) {
  score_kernel<<<1, 1024, 0, ctx->stream>>>(m, c, y, n, which, out_dev);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}
