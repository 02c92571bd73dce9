// Hyper-parameter grid sweep (replaces the 50 x 50 contour evaluation of CP:109-144):
// objective-only evaluations of cal_NLML (CP:68-73), cal_m_crps (CP:43-53), wrong_cal_m_crps
// (CP:55-64) and cal_m_logs (CP:75-85) at G independent (length scale, noise s.d.) points on
// 1-D inputs with k = 1 (CP:44).  Grid points are independent: multi-GPU runs split G.
//
// n <= 128 (the reference's n = 20): one CTA per grid point, everything in shared memory —
//   K = rbf + j^2 I, Cholesky, X = L^-1, d = diag(K^-1) = column norms of X, alpha = X'(X y).
// larger n: each point runs the blocked full-GP factorisation (objective only).
//
// In-sample ("wrong") CRPS uses K alpha = y:  mean = k_ff alpha = y - j^2 alpha and
// diag(j^2 I + k_ff - k_ff K^-1 k_ff) = 2 j^2 - j^4 d, algebraically equal to CP:58-59.
#include "gps_common.cuh"

namespace {

constexpr double INV_SQRT_PI = 0.56418958354775628695;
constexpr double INV_SQRT_2PI = 0.39894228040143267794;
constexpr double INV_SQRT2 = 0.70710678118654752440;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

__device__ __forceinline__ double crps_term(double y, double mu, double var) {
  const double s = sqrt(var), z = (y - mu) / s;
  return s * (z * erf(z * INV_SQRT2) + 2.0 * INV_SQRT_2PI * exp(-0.5 * z * z) - INV_SQRT_PI);
}
__device__ __forceinline__ double logs_term(double y, double mu, double var) {
  return (y - mu) * (y - mu) / (2.0 * var) + 0.5 * log(var) + HALF_LOG_2PI;
}

__device__ __forceinline__ double grid_row_value(int which, double y, double alpha, double d, double j2,
                                                 double logl) {
  if (which == GPS_GRID_NLML) return 0.5 * y * alpha + logl + HALF_LOG_2PI;            // CP:71, per row
  if (which == GPS_GRID_CRPS) return crps_term(y, y - alpha / d, 1.0 / d);               // CP:48-51
  if (which == GPS_GRID_LOGS) return logs_term(y, y - alpha / d, 1.0 / d + j2);          // CP:80-83
  return crps_term(y, y - j2 * alpha, 2.0 * j2 - j2 * j2 * d);                           // CP:58-62
}

__global__ void __launch_bounds__(128)
grid_small_kernel(const double* __restrict__ x, const double* __restrict__ y, int n,
                  const double* __restrict__ ls, const double* __restrict__ sds, int64_t G, int which,
                  double* __restrict__ out, int* __restrict__ info) {
  extern __shared__ double sh[];
  const int ld = n + 1;
  double* A = sh;                 // [n][ld]: lower = L, strict upper = (L^-1)' , diag = L_ii
  double* xv = A + (size_t)n * ld;
  double* yv = xv + n;
  double* dinv = yv + n;          // 1 / L_ii = X_ii
  double* tv = dinv + n;          // X y
  double* red = tv + n;           // [32]
  const int tid = threadIdx.x;
  if (tid < n) {
    xv[tid] = x[tid];
    yv[tid] = y[tid];
  }
  for (int64_t gidx = blockIdx.x; gidx < G; gidx += gridDim.x) {
    const double il = 1.0 / ls[gidx], sd = sds[gidx], j2 = sd * sd;
    __syncthreads();
    for (int e = tid; e < n * n; e += 128) {
      const int i = e / n, j = e - i * n;
      const double df = (xv[i] - xv[j]) * il;
      A[i * ld + j] = exp(-0.5 * df * df) + (i == j ? j2 : 0.0);   // CP:19, CP:45
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      if (tid == 0) {
        double p = A[j * ld + j];
        if (!(p > 0.0)) {
          atomicCAS(info, 0, j + 1);
          p = 1.0;
        }
        A[j * ld + j] = sqrt(p);
      }
      __syncthreads();
      if (tid > j && tid < n) {
        const double l = A[tid * ld + j] / A[j * ld + j];
        A[tid * ld + j] = l;
      }
      __syncthreads();
      if (tid > j && tid < n) {
        const double l = A[tid * ld + j];
        for (int c = j + 1; c <= tid; ++c) A[tid * ld + c] -= l * A[c * ld + j];
      }
      __syncthreads();
    }
    if (tid < n) dinv[tid] = 1.0 / A[tid * ld + tid];
    __syncthreads();
    // column tid of X = L^-1, stored in row tid of the strict upper triangle
    if (tid < n) {
      const int j = tid;
      for (int i = j + 1; i < n; ++i) {
        double s = 0.0;
        for (int k = j; k < i; ++k) {
          const double xk = (k == j) ? dinv[j] : A[j * ld + k];
          s = fma(A[i * ld + k], xk, s);
        }
        A[j * ld + i] = -s * dinv[i];
      }
    }
    __syncthreads();
    double d = 0.0, logl = 0.0;
    if (tid < n) {
      // t_k = sum_{i<=k} X[k][i] y_i : column tid of the upper storage
      double t = dinv[tid] * yv[tid];
      for (int i = 0; i < tid; ++i) t = fma(A[i * ld + tid], yv[i], t);
      tv[tid] = t;
      logl = -log(dinv[tid]);
    }
    __syncthreads();
    double val = 0.0;
    if (tid < n) {
      double alpha = dinv[tid] * tv[tid];
      d = dinv[tid] * dinv[tid];
      for (int k = tid + 1; k < n; ++k) {
        const double xk = A[tid * ld + k];   // X[k][tid]
        alpha = fma(xk, tv[k], alpha);
        d = fma(xk, xk, d);
      }
      val = grid_row_value(which, yv[tid], alpha, d, j2, logl);
    }
    const double s = block_sum(val, red);
    if (tid == 0) out[gidx] = (which == GPS_GRID_NLML) ? s : s / (double)n;
  }
}

// large n: per-row values from the blocked path's alpha, d, log diag(L)
__global__ void __launch_bounds__(1024)
grid_rows_kernel(int which, int64_t n, const double* __restrict__ y, const double* __restrict__ alpha,
                 const double* __restrict__ d, const double* __restrict__ logd, double j2,
                 double* __restrict__ out, const int* __restrict__ info, int* __restrict__ latch, int point) {
  __shared__ double red[32];
  if (threadIdx.x == 0 && *info != 0 && *latch == 0) *latch = point + 1;   // the lane's next POTRF clears info
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
    s += grid_row_value(which, y[i], alpha[i], d[i], j2, which == GPS_GRID_NLML ? logd[i] : 0.0);
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = (which == GPS_GRID_NLML) ? s : s / (double)n;
}

}  // namespace

extern "C" int gps_grid_eval(gps_ctx* ctx, const double* x, const double* y, int n, const double* ls,
                             const double* noise_sd, int64_t G, int which, double* out) {
  if (!ctx) return GPS_EINVAL;
  if (!x || !y || !ls || !noise_sd || !out || n <= 0 || G < 0 || which < GPS_GRID_NLML || which > GPS_GRID_LOGS)
    return gps_fail(ctx, GPS_EINVAL, "grid_eval: bad arguments");
  if (G == 0) return GPS_OK;
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  if (!ctx->d_info) GPS_CUDA(cudaMalloc(&ctx->d_info, sizeof(int)));
  GPS_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream));
  if (n <= 128) {
    const double *dx, *dy;
    GPS_CHECK(gps_stage_in(ctx, x, n, ctx->stage[0], &dx));
    GPS_CHECK(gps_stage_in(ctx, y, n, ctx->stage[1], &dy));
    GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)3 * G));
    double* dls = ctx->stage[2].p;
    double* dsd = dls + G;
    double* dout = dsd + G;
    GPS_CUDA(cudaMemcpyAsync(dls, ls, G * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GPS_CUDA(cudaMemcpyAsync(dsd, noise_sd, G * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = ((size_t)n * (n + 1) + 4 * n + 32) * sizeof(double);
    GPS_ONCE_PER_DEVICE(ctx);
    if (!configured) {
      GPS_CUDA(cudaFuncSetAttribute(grid_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(((size_t)128 * 129 + 4 * 128 + 32) * sizeof(double))));
      configured = true;
    }
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    const unsigned grid = (unsigned)(G < cap ? G : cap);
    grid_small_kernel<<<grid, 128, smem, ctx->stream>>>(dx, dy, n, dls, dsd, G, which, dout, ctx->d_info);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
    GPS_CUDA(cudaMemcpyAsync(out, dout, G * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return gps_check_info(ctx);
  }
  // large n: the blocked factorisation per grid point (1-D inputs, a = 0, b = log l, c = 2 log j).
  // At these sizes one evaluation is bound by the serial chain of diagonal blocks (1.6 ms at n = 2048 with
  // most SMs idle), so the points are dealt round-robin to LANES lane contexts, each with its own stream and
  // workspaces: their chains overlap and the GEMM launches of one lane fill the gaps of the others.  Nothing
  // is read back until every lane has drained; a failed factorisation is latched per lane.
  const int64_t Np = gps_pad(n);
  const int LANES = (int)std::min<int64_t>(G, Np <= 4096 ? 8 : (Np <= 8192 ? 2 : 1));
  while ((int)ctx->grid_lanes.size() < LANES) {
    gps_ctx* ln = new gps_ctx();
    ln->device = ctx->device;
    ln->sm_count = ctx->sm_count;
    if (cudaStreamCreateWithFlags(&ln->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ln;
      return gps_fail(ctx, GPS_ECUDA, "grid_eval: cannot create a lane stream");
    }
    ln->stream = ln->own_stream;
    ctx->grid_lanes.push_back(ln);
  }
  GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)G + 8));
  double* dres = ctx->stage[2].p;
  int* dlatch = reinterpret_cast<int*>(dres + G);   // one latch per lane: 1 + index of its first failed point
  GPS_CUDA(cudaMemsetAsync(dlatch, 0, 8 * sizeof(int), ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int l = 0; l < LANES; ++l) {
    gps_ctx* ln = ctx->grid_lanes[l];
    ln->gemm_variant = ctx->gemm_variant;
    ln->potf2_variant = ctx->potf2_variant;
    ln->overlap_trtri = ctx->overlap_trtri;
    ln->time_gemm = false;
    int r = gps_set_data(ln, x, y, n, 1);
    if (r == GPS_OK) r = gps_ensure_ws(ln, Np);
    if (r != GPS_OK) return gps_fail(ctx, r, "grid_eval lane %d: %s", l, ln->err.c_str());
  }
  int rc = GPS_OK;
  for (int64_t g = 0; g < G && rc == GPS_OK; ++g) {
    gps_ctx* ln = ctx->grid_lanes[g % LANES];
    double* v = ln->vecs.p;
    const double theta[3] = {0.0, log(ls[g]), 2.0 * log(noise_sd[g])};
    rc = gps_upload_params(ln, theta, 1, nullptr, nullptr);
    // objective values only: alpha and diag K^-1 come from L^-1 (two triangular sweeps + column sums of squares),
    // so the K^-1 = L^-T L^-1 stage (a third of the flops) is skipped
    if (rc == GPS_OK) rc = gps_factor_and_invert(ln, which == GPS_GRID_NLML, false);
    if (rc == GPS_OK) rc = gps_alpha_from_linv(ln, ln->Xb.p, Np, ln->y.p, v + V_U * Np, v + V_ALPHA * Np, v + V_D * Np);
    if (rc != GPS_OK) {
      gps_fail(ctx, rc, "grid_eval point %lld: %s", (long long)g, ln->err.c_str());
      break;
    }
    grid_rows_kernel<<<1, 1024, 0, ln->stream>>>(which, n, ln->y.p, v + V_ALPHA * Np, v + V_D * Np, v + V_LOGD * Np,
                                                 noise_sd[g] * noise_sd[g], dres + g, ln->d_info,
                                                 dlatch + (g % LANES), (int)g);
    if (cudaGetLastError() != cudaSuccess) rc = gps_fail(ctx, GPS_ECUDA, "grid_eval: launch failed");
  }
  for (int l = 0; l < LANES; ++l) {
    gps_ctx* ln = ctx->grid_lanes[l];
    if (cudaStreamSynchronize(ln->stream) != cudaSuccess && rc == GPS_OK) rc = gps_fail(ctx, GPS_ECUDA, "grid_eval: lane %d failed", l);
    ctx->launches += ln->launches + 0;
    ln->launches = 0;
  }
  ctx->launches += G;
  if (rc != GPS_OK) return rc;
  int latch[8];
  GPS_CUDA(cudaMemcpyAsync(out, dres, G * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(latch, dlatch, sizeof latch, cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t bad = -1;
  for (int l = 0; l < LANES; ++l)
    if (latch[l] && (bad < 0 || latch[l] - 1 < bad)) bad = latch[l] - 1;
  if (bad >= 0)
    return gps_fail(ctx, GPS_ENOTPD, "grid_eval: K + j^2 I not positive definite at grid point %lld (l = %g, j = %g)",
                    (long long)bad, ls[bad], noise_sd[bad]);
  return GPS_OK;
}
