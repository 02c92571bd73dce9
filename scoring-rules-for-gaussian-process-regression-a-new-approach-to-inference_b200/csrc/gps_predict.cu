// Prediction at test inputs (diagonal of the predictive covariance only) and the element-wise
// chol_solve twin.
//
// The reference forms the full T x T predictive covariance (cal_mean_and_cov, KF:121-126) and
// keeps its diagonal (KF:273); at T = 30 000 that matrix alone is 7.2 GB.  Here
//   mean_t = k_t' alpha,      var_t = sn2 + e^a - k_t' K^-1 k_t
// are produced per block of test rows: cross-Gram tile -> DMMA product with K^-1 -> fused row
// reduction.  Rows of the test set are independent, which is what multi-GPU runs shard.
#include "gps_common.cuh"

namespace {

// one warp per test row: mean = Ks[t,:] . alpha ; q = Ks[t,:] . V[t,:] ; var = sn2 + ea - q
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
    q = fma(kk.x, vv.x, q);
    q = fma(kk.y, vv.y, q);
  }
  m = warp_sum(m);
  q = warp_sum(q);
  if (lane == 0) {
    mean[t] = m;
    var[t] = par[1] + par[0] - q;
  }
}

__global__ void pad_identity_kernel(double* __restrict__ K, int64_t n, int64_t Np) {
  const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) K[i * Np + i] = 1.0;
}

}  // namespace

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
  GPS_CHECK(gps_factor_and_invert(ctx, false));
  GPS_CHECK(gps_check_info(ctx));
  ctx->loo_valid = false;
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
  // after the inversion L^-1 (Xb) and the scratch (Sb) are free: reuse them as the
  // cross-Gram block and its product with K^-1, Np rows at a time.
  const int64_t chunk = Np;
  std::vector<GemmTask> tasks;
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t rows = (T - t0 < chunk) ? (T - t0) : chunk;
    const int64_t rows_p = gps_pad(rows);
    GPS_CUDA(cudaMemsetAsync(ctx->Xb.p, 0, (size_t)rows_p * Np * sizeof(double), ctx->stream));
    GPS_CHECK(gps_gram_rect(ctx, dXs + t0 * D, rows, ctx->X.p, N, D, ctx->params.p, ctx->Xb.p, Np));
    tasks.clear();
    for (int ti = 0; ti < rows_p / GPS_TILE; ++ti)
      for (int tj = 0; tj < Np / GPS_TILE; ++tj) {
        GemmTask t;
        t.a_row = ti * GPS_TILE; t.b_row = tj * GPS_TILE; t.k0 = 0; t.k1 = (int)Np;
        t.c_row = ti * GPS_TILE; t.c_col = tj * GPS_TILE; t.flags = 0; t.pad = 0;
        tasks.push_back(t);
      }
    GPS_CHECK(gps_upload_tasks2(ctx, tasks));
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_KC, ctx->Xb.p, Np, ctx->Kb.p, Np, ctx->Sb.p, Np, 1.0, 0.0, nullptr,
                             false, ctx->d_tasks2, tasks.size()));
    predict_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(
        ctx->Xb.p, ctx->Sb.p, Np, rows, ctx->vecs.p + V_ALPHA * Np, ctx->params.p, dmean + t0, dvar + t0);
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
