// Prediction at test inputs (diagonal of the predictive covariance only) and the element-wise
// chol_solve twin.
//
// The reference forms the full T x T predictive covariance (cal_mean_and_cov, KF:121-126) and
// keeps its diagonal (KF:273); at T = 30 000 that matrix alone is 7.2 GB.  Here
//   mean_t = k_t' alpha,      var_t = sn2 + e^a - k_t' K^-1 k_t = sn2 + e^a - |L^-1 k_t|^2
// are produced per block of test rows: cross-Gram tile -> DMMA product with the TRIANGULAR factor inverse
// (block k-ranges stop at the diagonal: T N^2 flops, half of a product with the full K^-1) -> fused row
// reduction (dot with alpha, sum of squares).  Rows of the test set are independent, which is what
// multi-GPU runs shard.
#include "gps_common.cuh"

namespace {

// one warp per test row: mean = Ks[t,:] . alpha ; q = |V[t,:]|^2 with V = Ks L^-T ; var = sn2 + ea - q
__global__ void __launch_bounds__(256)
predict_rows_kernel(const double* __restrict__ Ks, const double* __restrict__ V, int64_t ld, int64_t rows,
                    const double* __restrict__ alpha, const double* __restrict__ par,
                    double* __restrict__ mean, double* __restrict__ var) {
  const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= rows) return;
  const double2* k = reinterpret_cast<const double2*>(Ks + t * ld);
  const double2* v = reinterpret_cast<const double2*>(V + t * ld);
  const double2* a = reinterpret_cast<const double2*>(alpha);
  double m = 0.0, q = 0.0;
  for (int64_t j = lane; j < ld / 2; j += 32) {
    const double2 kk = k[j], vv = v[j], aa = a[j];
    m = fma(kk.x, aa.x, m);
    m = fma(kk.y, aa.y, m);
    q = fma(vv.x, vv.x, q);
    q = fma(vv.y, vv.y, q);
  }
  m = warp_sum(m);
  q = warp_sum(q);
  if (lane == 0) {
    mean[t] = m;
    var[t] = par[1] + par[0] - q;
  }
}

// u = L^-1 y on the block-lower-triangular Xinv (row i reads the columns below (i / 128 + 1) * 128; the diagonal blocks
// hold explicit zeros above the diagonal, the strict upper block triangle is undefined): one warp per row
__global__ void __launch_bounds__(256)
trmv_lower_kernel(const double* __restrict__ A, int64_t Np, const double* __restrict__ x, double* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Np) return;
  const int64_t jmax = (row / GPS_TILE + 1) * GPS_TILE;
  const double2* a = reinterpret_cast<const double2*>(A + row * Np);
  const double2* xv = reinterpret_cast<const double2*>(x);
  double s0 = 0.0, s1 = 0.0;
  for (int64_t j = lane; j < jmax / 2; j += 32) {
    const double2 aa = a[j], xx = xv[j];
    s0 = fma(aa.x, xx.x, s0);
    s1 = fma(aa.y, xx.y, s1);
  }
  const double s = warp_sum(s0 + s1);
  if (lane == 0) out[row] = s;
}

// alpha = L^-T u: out[j] = sum over i >= (j / 128) * 128 of Xinv[i][j] u[i].  A block owns 32 columns; its 8 warps take
// the rows i = i0 + w, i0 + w + 8, ...; fixed-order combine in shared memory (grid = Np / 32)
__global__ void __launch_bounds__(256)
trmv_lower_t_kernel(const double* __restrict__ A, int64_t Np, const double* __restrict__ u, double* __restrict__ out) {
  __shared__ double sh[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + lane;
  const int64_t i0 = (j / GPS_TILE) * GPS_TILE;
  double s0 = 0.0, s1 = 0.0;
  int64_t i = i0 + w;
  for (; i + 8 < Np; i += 16) {
    s0 = fma(A[i * Np + j], u[i], s0);
    s1 = fma(A[(i + 8) * Np + j], u[i + 8], s1);
  }
  if (i < Np) s0 = fma(A[i * Np + j], u[i], s0);
  sh[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sh[q][lane];
    out[j] = t;
  }
}

// d[j] = (K^-1)_jj = sum over i >= j of Xinv[i][j]^2, same decomposition as trmv_lower_t_kernel
__global__ void __launch_bounds__(256)
colsq_lower_kernel(const double* __restrict__ A, int64_t Np, double* __restrict__ out) {
  __shared__ double sh[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + lane;
  const int64_t i0 = (j / GPS_TILE) * GPS_TILE;
  double s0 = 0.0, s1 = 0.0;
  int64_t i = i0 + w;
  for (; i + 8 < Np; i += 16) {
    const double a = A[i * Np + j], b = A[(i + 8) * Np + j];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  if (i < Np) {
    const double a = A[i * Np + j];
    s0 = fma(a, a, s0);
  }
  sh[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sh[q][lane];
    out[j] = t;
  }
}

__global__ void pad_identity_kernel(double* __restrict__ K, int64_t n, int64_t Np) {
  const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) K[i * Np + i] = 1.0;
}

}  // namespace

// alpha = K^-1 y = L^-T (L^-1 y) and (optionally) d = diag K^-1 from the factor inverse alone (no LAUUM):
// what prediction and the objective-only grid sweep need
int gps_alpha_from_linv(gps_ctx* ctx, const double* Xinv, int64_t Np, const double* y, double* u, double* alpha, double* d) {
  trmv_lower_kernel<<<(unsigned)((Np + 7) / 8), 256, 0, ctx->stream>>>(Xinv, Np, y, u);
  GPS_LAUNCH_CHECK();
  trmv_lower_t_kernel<<<(unsigned)(Np / 32), 256, 0, ctx->stream>>>(Xinv, Np, u, alpha);
  GPS_LAUNCH_CHECK();
  ctx->launches += 2;
  if (d) {
    colsq_lower_kernel<<<(unsigned)(Np / 32), 256, 0, ctx->stream>>>(Xinv, Np, d);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
  }
  return GPS_OK;
}

extern "C" {

int gps_full_predict(gps_ctx* ctx, const double* theta, const double* Xs, int64_t T, double* mean,
                     double* var) {
  if (!ctx) return GPS_EINVAL;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "full_predict: call gps_set_data first");
  if (!theta || !Xs || !mean || !var || T < 0) return gps_fail(ctx, GPS_EINVAL, "full_predict: bad arguments");
  if (T == 0) return GPS_OK;
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t N = ctx->N, Np = ctx->Np;
  const int D = ctx->D;
  GPS_CHECK(gps_ensure_ws(ctx, Np));
  GPS_CHECK(gps_upload_params(ctx, theta, D, nullptr, nullptr));
  ctx->gemm_events_used = 0;
  // factor and invert the factor only: alpha = L^-T (L^-1 y) by two triangular sweeps, no K^-1 (saves the LAUUM stage)
  GPS_CHECK(gps_factor_and_invert(ctx, false, false));
  GPS_CHECK(gps_check_info(ctx));
  ctx->loo_valid = false;
  GPS_CHECK(gps_alpha_from_linv(ctx, ctx->Xb.p, Np, ctx->y.p, ctx->vecs.p + V_U * Np, ctx->vecs.p + V_ALPHA * Np, nullptr));
  const double* dXs;
  GPS_CHECK(gps_stage_in(ctx, Xs, (size_t)T * D, ctx->stage[0], &dXs));
  const bool dev_out = gps_is_device_ptr(mean) && gps_is_device_ptr(var);
  double *dmean = mean, *dvar = var;
  if (!dev_out) {
    GPS_CHECK(gps_ensure(ctx, ctx->stage[1], (size_t)T));
    GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)T));
    dmean = ctx->stage[1].p;
    dvar = ctx->stage[2].p;
  }
  // alpha = K^-1 y is in the vectors; from here on only L^-1 (Xb, lower block triangle, diagonal blocks with explicit
  // zeros above the diagonal) is needed: the scratch (Sb) takes the cross-Gram block, K^-1's buffer (Kb) its product
  // with L^-T, Np test rows at a time.  Task (ti, tj) contracts over k < (tj + 1) * 128 only; longest tasks first.
  const int64_t chunk = Np;
  std::vector<GemmTask> tasks;
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t rows = (T - t0 < chunk) ? (T - t0) : chunk;
    const int64_t rows_p = gps_pad(rows);
    GPS_CUDA(cudaMemsetAsync(ctx->Sb.p, 0, (size_t)rows_p * Np * sizeof(double), ctx->stream));
    GPS_CHECK(gps_gram_rect(ctx, dXs + t0 * D, rows, ctx->X.p, N, D, ctx->params.p, ctx->Sb.p, Np));
    tasks.clear();
    for (int tj = (int)(Np / GPS_TILE) - 1; tj >= 0; --tj)
      for (int ti = 0; ti < rows_p / GPS_TILE; ++ti) {
        GemmTask t;
        t.a_row = ti * GPS_TILE; t.b_row = tj * GPS_TILE; t.k0 = 0; t.k1 = (tj + 1) * GPS_TILE;
        t.c_row = ti * GPS_TILE; t.c_col = tj * GPS_TILE; t.flags = 0; t.pad = 0;
        tasks.push_back(t);
      }
    GPS_CHECK(gps_upload_tasks2(ctx, tasks));
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_KC, ctx->Sb.p, Np, ctx->Xb.p, Np, ctx->Kb.p, Np, 1.0, 0.0, nullptr,
                             false, ctx->d_tasks2, tasks.size()));
    predict_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(
        ctx->Sb.p, ctx->Kb.p, Np, rows, ctx->vecs.p + V_ALPHA * Np, ctx->params.p, dmean + t0, dvar + t0);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
  }
  if (!dev_out) {
    GPS_CUDA(cudaMemcpyAsync(mean, dmean, T * sizeof(double), cudaMemcpyDefault, ctx->stream));
    GPS_CUDA(cudaMemcpyAsync(var, dvar, T * sizeof(double), cudaMemcpyDefault, ctx->stream));
  }
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

int gps_chol_solve(gps_ctx* ctx, const double* B, const double* A, int64_t n, int64_t nrhs, double* out) {
  if (!ctx) return GPS_EINVAL;
  if (!A || !B || !out || n <= 0 || nrhs <= 0) return gps_fail(ctx, GPS_EINVAL, "chol_solve: bad arguments");
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t Np = gps_pad(n);
  GPS_CHECK(gps_ensure_ws(ctx, Np));
  ctx->loo_valid = false;
  ctx->gemm_events_used = 0;
  // A -> Kb, padded with the identity
  GPS_CUDA(cudaMemsetAsync(ctx->Kb.p, 0, (size_t)Np * Np * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemcpy2DAsync(ctx->Kb.p, Np * sizeof(double), A, n * sizeof(double), n * sizeof(double), n,
                             cudaMemcpyDefault, ctx->stream));
  if (Np > n) {
    pad_identity_kernel<<<(unsigned)((Np - n + 127) / 128), 128, 0, ctx->stream>>>(ctx->Kb.p, n, Np);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
  }
  GPS_CHECK(gps_potrf(ctx, ctx->Kb.p, ctx->Xb.p, Np));
  GPS_CHECK(gps_trtri(ctx, ctx->Kb.p, ctx->Xb.p, ctx->Sb.p, Np));
  GPS_CHECK(gps_lauum(ctx, ctx->Xb.p, ctx->Kb.p, Np));
  GPS_CHECK(gps_check_info(ctx));
  // out = Kinv * B, Np right-hand sides at a time (B block in Sb, product in Xb)
  std::vector<GemmTask> tasks;
  for (int64_t c0 = 0; c0 < nrhs; c0 += Np) {
    const int64_t cols = (nrhs - c0 < Np) ? (nrhs - c0) : Np;
    const int64_t cols_p = gps_pad(cols);
    GPS_CUDA(cudaMemsetAsync(ctx->Sb.p, 0, (size_t)Np * cols_p * sizeof(double), ctx->stream));
    GPS_CUDA(cudaMemcpy2DAsync(ctx->Sb.p, cols_p * sizeof(double), B + c0, nrhs * sizeof(double),
                               cols * sizeof(double), n, cudaMemcpyDefault, ctx->stream));
    tasks.clear();
    for (int ti = 0; ti < Np / GPS_TILE; ++ti)
      for (int tj = 0; tj < cols_p / GPS_TILE; ++tj) {
        GemmTask t;
        t.a_row = ti * GPS_TILE; t.b_row = tj * GPS_TILE; t.k0 = 0; t.k1 = (int)Np;
        t.c_row = ti * GPS_TILE; t.c_col = tj * GPS_TILE; t.flags = 0; t.pad = 0;
        tasks.push_back(t);
      }
    GPS_CHECK(gps_upload_tasks2(ctx, tasks));
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, ctx->Kb.p, Np, ctx->Sb.p, cols_p, ctx->Xb.p, cols_p, 1.0, 0.0,
                             nullptr, false, ctx->d_tasks2, tasks.size()));
    GPS_CUDA(cudaMemcpy2DAsync(out + c0, nrhs * sizeof(double), ctx->Xb.p, cols_p * sizeof(double),
                               cols * sizeof(double), n, cudaMemcpyDefault, ctx->stream));
  }
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

// out[m,n] = A[m,k] B[k,n] for arbitrary sizes (UVA pointers): zero-padded copies to the 128 tile, one task-list
// launch of the DMMA tile GEMM, 2-D copy back.  Serves the same-signature twins of Q (KF:32-39),
// cal_mean_and_cov (KF:121-126) and spgp_cal_mean_and_cov (K20:76-83), whose products the reference does with
// torch.mm.
int gps_matmul(gps_ctx* ctx, const double* A, const double* B, int64_t m, int64_t k, int64_t n, double* out) {
  if (!ctx) return GPS_EINVAL;
  if (!A || !B || !out || m <= 0 || k <= 0 || n <= 0) return gps_fail(ctx, GPS_EINVAL, "matmul: bad arguments");
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t Mp = gps_pad(m), Kp = gps_pad(k), Npp = gps_pad(n);
  GPS_CHECK(gps_ensure(ctx, ctx->stage[0], (size_t)Mp * Kp));
  GPS_CHECK(gps_ensure(ctx, ctx->stage[1], (size_t)Kp * Npp));
  GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)Mp * Npp));
  GPS_CUDA(cudaMemsetAsync(ctx->stage[0].p, 0, (size_t)Mp * Kp * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemsetAsync(ctx->stage[1].p, 0, (size_t)Kp * Npp * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemcpy2DAsync(ctx->stage[0].p, Kp * sizeof(double), A, k * sizeof(double), k * sizeof(double), m,
                             cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaMemcpy2DAsync(ctx->stage[1].p, Npp * sizeof(double), B, n * sizeof(double), n * sizeof(double), k,
                             cudaMemcpyDefault, ctx->stream));
  std::vector<GemmTask> tasks;
  for (int ti = 0; ti < Mp / GPS_TILE; ++ti)
    for (int tj = 0; tj < Npp / GPS_TILE; ++tj) {
      GemmTask t;
      t.a_row = ti * GPS_TILE; t.b_row = tj * GPS_TILE; t.k0 = 0; t.k1 = (int)Kp;
      t.c_row = ti * GPS_TILE; t.c_col = tj * GPS_TILE; t.flags = 0; t.pad = 0;
      tasks.push_back(t);
    }
  GPS_CHECK(gps_upload_tasks2(ctx, tasks));
  ctx->gemm_events_used = 0;
  GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, ctx->stage[0].p, Kp, ctx->stage[1].p, Npp, ctx->stage[2].p, Npp, 1.0, 0.0,
                           nullptr, false, ctx->d_tasks2, tasks.size()));
  GPS_CUDA(cudaMemcpy2DAsync(out, n * sizeof(double), ctx->stage[2].p, Npp * sizeof(double), n * sizeof(double), m,
                             cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

}  // extern "C"
