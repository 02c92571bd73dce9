// 4-fold block-LOO DSS objective + gradient for the full GP (replaces KF:499-543).
//
// With B = K^-1 (already formed for the LOO scores) the fold predictive of KF:508-530 is
//   m_f = y_f - B_ff^-1 (B y)_f,   cov_f = B_ff^-1,
// so  dss_f = n_f/2 log 2pi - 1/2 log|B_ff| + 1/2 a_f' B_ff^-1 a_f   (a = B y), and
//   dL/dB_ff = Gamma_f = -1/2 (C_f + abar_f abar_f'),  C_f = B_ff^-1,  abar_f = C_f a_f,
//   dL/dK    = -(B Gamma B + sym(u a')),  u = B abar.
// Each B_ff is copied into a tile-padded scratch problem and inverted with the same blocked
// POTRF / TRTRI / LAUUM as the big matrix (a child context of size N/4); Gamma is scattered into a
// block-diagonal N x N buffer with explicit zeros, so  T = Gamma B  and  S = B T  are plain
// task-list tile GEMMs (the k-range of a T tile covers only the folds its rows touch) and the
// gradient contraction is the one the LOO scores use.  The reference sizes every fold with
// index1 = N/4 (KF:521-530): N must be a multiple of 4.
#include "gps_common.cuh"

namespace {

constexpr double HALF_LOG_2PI = 0.91893853320467274178;

// out[Fp][Fp] = B[off.., off..] (nf x nf), identity padding
__global__ void __launch_bounds__(256)
fold_copy_kernel(const double* __restrict__ B, int64_t Np, int64_t off, int64_t nf, int64_t Fp,
                 double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Fp * Fp) return;
  const int64_t r = e / Fp, c = e - r * Fp;
  out[e] = (r < nf && c < nf) ? B[(off + r) * Np + off + c] : (r == c ? 1.0 : 0.0);
}

__global__ void fold_vec_kernel(const double* __restrict__ alpha, int64_t off, int64_t nf, int64_t Fp,
                                double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < Fp) out[r] = (r < nf) ? alpha[off + r] : 0.0;
}

// Gamma[off + r][off + c] = -1/2 (C_f[r][c] + abar_f[r] abar_f[c]);  abar_full[off + r] = abar_f[r]
__global__ void __launch_bounds__(256)
fold_gamma_kernel(const double* __restrict__ Cf, const double* __restrict__ abf, int64_t Fp, int64_t nf,
                  int64_t off, int64_t Np, double* __restrict__ Gamma, double* __restrict__ abar_full) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nf * nf) return;
  const int64_t r = e / nf, c = e - r * nf;
  Gamma[(off + r) * Np + off + c] = -0.5 * (Cf[r * Fp + c] + abf[r] * abf[c]);
  if (c == 0) abar_full[off + r] = abf[r];
}

// obj += n_f/2 log 2pi - sum log diag(L_f) + 1/2 a_f' abar_f      (one block, fixed order)
__global__ void __launch_bounds__(1024)
fold_value_kernel(const double* __restrict__ logdiag, const double* __restrict__ af,
                  const double* __restrict__ abf, int64_t nf, double* __restrict__ obj, int first) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t r = threadIdx.x; r < nf; r += blockDim.x) s += HALF_LOG_2PI - logdiag[r] + 0.5 * af[r] * abf[r];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) obj[0] = (first ? 0.0 : obj[0]) + s;
}

// obj = sum over folds of their values, in fold order (one thread: fixed order)
__global__ void fold_sum_kernel(const double* __restrict__ fv, int64_t stride, int64_t at, int folds, double* __restrict__ obj) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int f = 0; f < folds; ++f) s += fv[f * stride + at];
    obj[0] = s;
  }
}

}  // namespace

static int full_dss_body(gps_ctx* ctx, double* par_obj, double* par_gsum, bool want_grad);

int gps_full_dss(gps_ctx* ctx, double* par_obj, double* par_gsum, bool want_grad) {
  const int rc = full_dss_body(ctx, par_obj, par_gsum, want_grad);
  if (rc != GPS_OK && ctx->dss_fork) {
    // error exit: the fold streams may still hold queued work that reads the parent's buffers; join them into
    // the caller's stream so that nothing issued later on this context can overtake it
    for (size_t f = 0; f < ctx->fold_lanes.size() && f < 4; ++f)
      if (cudaEventRecord(ctx->dss_join[f], ctx->fold_lanes[f]->stream) == cudaSuccess)
        cudaStreamWaitEvent(ctx->stream, ctx->dss_join[f], 0);
    cudaGetLastError();
  }
  return rc;
}

static int full_dss_body(gps_ctx* ctx, double* par_obj, double* par_gsum, bool want_grad) {
  const int64_t N = ctx->N, Np = ctx->Np;
  const int D = ctx->D;
  constexpr int FOLDS = 4;
  if (N % FOLDS) return gps_fail(ctx, GPS_EINVAL, "dss: the reference's 4-fold code needs N %% 4 == 0 (KF:521-530), N=%lld", (long long)N);
  const int64_t nf = N / FOLDS, Fp = gps_pad(nf);
  double* v = ctx->vecs.p;
  // Four child contexts, one per fold, each with its own stream and workspaces: the N/4-sized factorisations
  // reuse the blocked drivers (own task lists) and, being bound by their serial diagonal-block chains, run
  // side by side instead of one after the other.
  while ((int)ctx->fold_lanes.size() < FOLDS) {
    gps_ctx* ln = new gps_ctx();
    ln->device = ctx->device;
    ln->sm_count = ctx->sm_count;
    if (cudaStreamCreateWithFlags(&ln->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ln;
      return gps_fail(ctx, GPS_ECUDA, "dss: cannot create a fold stream");
    }
    ln->stream = ln->own_stream;
    ctx->fold_lanes.push_back(ln);
  }
  if (!ctx->dss_fork) {
    GPS_CUDA(cudaEventCreateWithFlags(&ctx->dss_fork, cudaEventDisableTiming));
    for (auto& e : ctx->dss_join) GPS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  GPS_CHECK(gps_ensure(ctx, ctx->Gb, (size_t)Np * Np));
  GPS_CHECK(gps_ensure(ctx, ctx->fold_vecs, (size_t)FOLDS * (3 * Fp + 8)));
  GPS_CUDA(cudaMemsetAsync(ctx->Gb.p, 0, (size_t)Np * Np * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemsetAsync(v + V_ABAR * Np, 0, Np * sizeof(double), ctx->stream));
  GPS_CUDA(cudaEventRecord(ctx->dss_fork, ctx->stream));
  int rc = GPS_OK;
  for (int f = 0; f < FOLDS && rc == GPS_OK; ++f) {
    gps_ctx* ch = ctx->fold_lanes[f];
    ch->gemm_variant = ctx->gemm_variant;
    ch->potf2_variant = ctx->potf2_variant;
    ch->overlap_trtri = ctx->overlap_trtri;
    ch->time_gemm = false;
    rc = gps_ensure_ws(ch, Fp);
    if (rc != GPS_OK) return gps_fail(ctx, rc, "dss fold %d: %s", f, ch->err.c_str());
    cudaStream_t fs = ch->stream;
    GPS_CUDA(cudaStreamWaitEvent(fs, ctx->dss_fork, 0));
    double* af = ctx->fold_vecs.p + (size_t)f * (3 * Fp + 8);   // alpha_f padded
    double* abf = af + Fp;                                        // abar_f
    double* ldf = abf + Fp;                                       // log diag(L_f)
    double* objf = ldf + Fp;                                      // this fold's value
    const int64_t off = f * nf;
    fold_copy_kernel<<<(unsigned)((Fp * Fp + 255) / 256), 256, 0, fs>>>(ctx->Kb.p, Np, off, nf, Fp, ch->Kb.p);
    GPS_LAUNCH_CHECK();
    fold_vec_kernel<<<(unsigned)((Fp + 255) / 256), 256, 0, fs>>>(v + V_ALPHA * Np, off, nf, Fp, af);
    GPS_LAUNCH_CHECK();
    ctx->launches += 2;
    rc = gps_potrf(ch, ch->Kb.p, ch->Xb.p, Fp);
    if (rc == GPS_OK) rc = gps_diag_extract(ch, ch->Kb.p, Fp, ldf, 1);
    if (rc == GPS_OK) rc = gps_trtri(ch, ch->Kb.p, ch->Xb.p, ch->Sb.p, Fp);
    if (rc == GPS_OK) rc = gps_lauum(ch, ch->Xb.p, ch->Kb.p, Fp);          // C_f = B_ff^-1
    if (rc == GPS_OK) rc = gps_symv(ch, ch->Kb.p, Fp, af, abf);              // abar_f = C_f a_f
    if (rc != GPS_OK) return gps_fail(ctx, rc, "dss fold %d: %s", f, ch->err.c_str());
    fold_value_kernel<<<1, 1024, 0, fs>>>(ldf, af, abf, nf, objf, 1);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
    if (want_grad) {
      fold_gamma_kernel<<<(unsigned)((nf * nf + 255) / 256), 256, 0, fs>>>(ch->Kb.p, abf, Fp, nf, off, Np, ctx->Gb.p,
                                                                          v + V_ABAR * Np);
      GPS_LAUNCH_CHECK();
      ctx->launches++;
    }
    GPS_CUDA(cudaEventRecord(ctx->dss_join[f], fs));
  }
  for (int f = 0; f < FOLDS; ++f) GPS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->dss_join[f], 0));
  // value = sum of the fold values in fold order; a failed fold factorisation is reported like a failed K
  fold_sum_kernel<<<1, 32, 0, ctx->stream>>>(ctx->fold_vecs.p, 3 * Fp + 8, 3 * Fp, FOLDS, par_obj);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  for (int f = 0; f < FOLDS; ++f) {
    gps_ctx* ch = ctx->fold_lanes[f];
    int info = 0;
    GPS_CUDA(cudaMemcpyAsync(&info, ch->d_info, sizeof(int), cudaMemcpyDeviceToHost, ch->stream));
    GPS_CUDA(cudaStreamSynchronize(ch->stream));
    ctx->launches += ch->launches;
    ch->launches = 0;
    if (info != 0) return gps_fail(ctx, GPS_ENOTPD, "dss: fold %d block of K^-1 not positive definite at pivot %d", f, info);
  }
  if (!want_grad) return GPS_OK;
  // u = B abar
  GPS_CHECK(gps_symv(ctx, ctx->Kb.p, Np, v + V_ABAR * Np, v + V_U * Np));
  // T = Gamma B: tile (i, j) only needs the k-tiles covered by the folds that rows of tile i touch
  const int nb = (int)(Np / GPS_TILE);
  std::vector<GemmTask> tasks;
  tasks.reserve((size_t)nb * nb);
  for (int i = 0; i < nb; ++i) {
    const int64_t r0 = (int64_t)i * GPS_TILE, r1 = std::min<int64_t>(N, r0 + GPS_TILE);
    int k0 = 0, k1 = 0;
    if (r0 < N) {
      const int64_t f0 = r0 / nf, f1 = (r1 - 1) / nf;
      k0 = (int)((f0 * nf) / GPS_TILE * GPS_TILE);
      k1 = (int)std::min<int64_t>(Np, gps_pad((f1 + 1) * nf));
    }
    for (int j = 0; j < nb; ++j) {
      GemmTask t;
      t.a_row = i * GPS_TILE; t.b_row = j * GPS_TILE; t.k0 = k0; t.k1 = k1;
      t.c_row = i * GPS_TILE; t.c_col = j * GPS_TILE; t.flags = 0; t.pad = 0;
      tasks.push_back(t);
    }
  }
  GPS_CHECK(gps_upload_tasks2(ctx, tasks));
  GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, ctx->Gb.p, Np, ctx->Kb.p, Np, ctx->Xb.p, Np, 1.0, 0.0, nullptr, false,
                           ctx->d_tasks2, tasks.size()));
  // S = B T (lower tiles, full k) — the task list of the symmetric product, B operand now [k][n]
  GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, ctx->Kb.p, Np, ctx->Xb.p, Np, ctx->Sb.p, Np, 1.0, 0.0, nullptr, false,
                           ctx->d_tasks + ctx->symprod.off, ctx->symprod.cnt));
  GPS_CHECK(gps_grad_contract(ctx, 0, ctx->Sb.p, N, Np, ctx->X.p, D, ctx->params.p, v + V_ALPHA * Np, v + V_U * Np,
                              par_gsum));
  return GPS_OK;
}
