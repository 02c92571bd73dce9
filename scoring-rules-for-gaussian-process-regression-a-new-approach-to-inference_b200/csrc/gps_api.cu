// extern "C" boundary of libgpscore.so (see include/gpscore.h) — context, data staging and the
// full-GP evaluation / prediction drivers.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "gps_common.cuh"

int gps_fail(gps_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

int gps_ensure(gps_ctx* ctx, DevBuf& b, size_t n) {
  if (b.n >= n && b.p) return GPS_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.n = 0;
  cudaError_t e = cudaMalloc(&b.p, n * sizeof(double));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return gps_fail(ctx, GPS_ENOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(double),
                    cudaGetErrorString(e));
  }
  b.n = n;
  return GPS_OK;
}

bool gps_is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}


// Input staging: returns a device pointer holding `n` doubles of `p` (p itself if already on the
// device, else a copy in `tmp`).
int gps_stage_in(gps_ctx* ctx, const double* p, size_t n, DevBuf& tmp, const double** out) {
  if (n == 0) { *out = p; return GPS_OK; }
  if (gps_is_device_ptr(p)) { *out = p; return GPS_OK; }
  GPS_CHECK(gps_ensure(ctx, tmp, n));
  GPS_CUDA(cudaMemcpyAsync(tmp.p, p, n * sizeof(double), cudaMemcpyDefault, ctx->stream));
  *out = tmp.p;
  return GPS_OK;
}

int gps_ensure_ws(gps_ctx* ctx, int64_t Np) {
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  if (!ctx->d_info) GPS_CUDA(cudaMalloc(&ctx->d_info, sizeof(int)));
  if (ctx->ws_Np != Np) {
    GPS_CHECK(gps_ensure(ctx, ctx->Kb, (size_t)Np * Np));
    GPS_CHECK(gps_ensure(ctx, ctx->Xb, (size_t)Np * Np));
    GPS_CHECK(gps_ensure(ctx, ctx->Sb, (size_t)Np * Np));
    GPS_CHECK(gps_ensure(ctx, ctx->vecs, (size_t)V_COUNT * Np));
    GPS_CHECK(gps_build_tasks(ctx, Np));
    ctx->ws_Np = Np;
    ctx->loo_valid = false;
  }
  return GPS_OK;
}

int gps_upload_tasks2(gps_ctx* ctx, const std::vector<GemmTask>& h) {
  if (h.size() > ctx->tasks2_cap) {
    if (ctx->d_tasks2) cudaFree(ctx->d_tasks2);
  if (ctx->d_tickets) cudaFree(ctx->d_tickets);
    ctx->d_tasks2 = nullptr;
    GPS_CUDA(cudaMalloc(&ctx->d_tasks2, h.size() * sizeof(GemmTask)));
    ctx->tasks2_cap = h.size();
  }
  GPS_CUDA(cudaMemcpyAsync(ctx->d_tasks2, h.data(), h.size() * sizeof(GemmTask), cudaMemcpyHostToDevice,
                           ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));  // h may be a temporary
  return GPS_OK;
}

int gps_upload_params(gps_ctx* ctx, const double* theta, int D, double* ea_out, double* sn2_out) {
  double h[2 + 64];
  if (D > 64) return gps_fail(ctx, GPS_EINVAL, "D=%d exceeds 64", D);
  h[0] = exp(theta[0]);
  h[1] = exp(theta[D + 1]);
  for (int d = 0; d < D; ++d) h[2 + d] = exp(-theta[1 + d]);
  GPS_CUDA(cudaMemcpyAsync(ctx->params.p, h, (2 + D) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // h is on the stack: the copy from pageable memory is staged before the call returns
  if (ea_out) *ea_out = h[0];
  if (sn2_out) *sn2_out = h[1];
  return GPS_OK;
}

namespace {

void gemm_timing_begin(gps_ctx* ctx) { ctx->gemm_events_used = 0; }

int gemm_timing_end(gps_ctx* ctx) {
  double ms = 0;
  for (size_t i = 0; i < ctx->gemm_events_used; ++i) {
    float t = 0;
    GPS_CUDA(cudaEventElapsedTime(&t, ctx->gemm_events[i].first, ctx->gemm_events[i].second));
    ms += t;
  }
  ctx->last_gemm_ms = ms;
  ctx->last_gemm_launches = (int64_t)ctx->gemm_events_used;
  return GPS_OK;
}

}  // namespace

static int stage_mark(gps_ctx* ctx, int which) {
  if (!ctx->stage_ev[which]) GPS_CUDA(cudaEventCreate(&ctx->stage_ev[which]));
  GPS_CUDA(cudaEventRecord(ctx->stage_ev[which], ctx->stream));
  return GPS_OK;
}

// K = ARD(X, X) + sn2 I  ->  L, L^-1, K^-1 (in Kb), alpha.  logdiag optionally.
int gps_factor_and_invert(gps_ctx* ctx, bool want_logdet, bool want_kinv) {
  const int64_t N = ctx->N, Np = ctx->Np;
  double* v = ctx->vecs.p;
  GPS_CHECK(stage_mark(ctx, gps_ctx::ST_BEGIN));
  GPS_CHECK(gps_gram_sym(ctx, ctx->X.p, N, Np, ctx->D, ctx->params.p, ctx->Kb.p));
  GPS_CHECK(stage_mark(ctx, gps_ctx::ST_GRAM));
  if (ctx->overlap_trtri == 1) {
    // POTRF and the inversion merges overlap: the two stages are reported as one ("potrf" = both, "trtri" = 0)
    GPS_CHECK(gps_potrf_trtri(ctx, ctx->Kb.p, ctx->Xb.p, ctx->Sb.p, Np));
    if (want_logdet) GPS_CHECK(gps_diag_extract(ctx, ctx->Kb.p, Np, v + V_LOGD * Np, 1));
    GPS_CHECK(stage_mark(ctx, gps_ctx::ST_POTRF));
  } else {
    GPS_CHECK(gps_potrf(ctx, ctx->Kb.p, ctx->Xb.p, Np));
    if (want_logdet) GPS_CHECK(gps_diag_extract(ctx, ctx->Kb.p, Np, v + V_LOGD * Np, 1));
    GPS_CHECK(stage_mark(ctx, gps_ctx::ST_POTRF));
    GPS_CHECK(gps_trtri(ctx, ctx->Kb.p, ctx->Xb.p, ctx->Sb.p, Np));
  }
  GPS_CHECK(stage_mark(ctx, gps_ctx::ST_TRTRI));
  if (!want_kinv) return GPS_OK;            // prediction: L^-1 (Xb) is all it needs
  GPS_CHECK(gps_lauum(ctx, ctx->Xb.p, ctx->Kb.p, Np));
  GPS_CHECK(stage_mark(ctx, gps_ctx::ST_LAUUM));
  GPS_CHECK(gps_symv(ctx, ctx->Kb.p, Np, ctx->y.p, v + V_ALPHA * Np));
  return GPS_OK;
}

// free a child context (it borrows the parent's stream)
void gps_ctx_release(gps_ctx* ch) {
  ch->stream = nullptr;
  ch->own_stream = nullptr;
  DevBuf* cb[] = {&ch->Kb, &ch->Xb, &ch->Sb, &ch->vecs, &ch->red, &ch->params};
  for (DevBuf* b : cb)
    if (b->p) cudaFree(b->p);
  if (ch->d_info) cudaFree(ch->d_info);
  if (ch->d_tasks) cudaFree(ch->d_tasks);
  if (ch->d_tickets) cudaFree(ch->d_tickets);
  for (auto* v : {&ch->potrf_events, &ch->tile_events, &ch->below_events, &ch->trailA1_events})
    for (auto e : *v) cudaEventDestroy(e);
  for (cudaStream_t s : {ch->panel_stream, ch->panel2_stream, ch->trail_stream, ch->tri_stream})
    if (s) cudaStreamDestroy(s);
  for (cudaEvent_t e : {ch->fork_ev, ch->join_trail_ev, ch->join_tri_ev})
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ch->stage_ev)
    if (e) cudaEventDestroy(e);
  for (auto& pr : ch->trace) cudaEventDestroy(pr.second);
  delete ch;
}

extern "C" {

const char* gps_version(void) { return "gpscore-b200 0.1 (sm_100a)"; }

int gps_create(int device, gps_ctx** out) {
  if (!out) return GPS_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return GPS_ENODEVICE;  // there is no CPU fallback
  }
  if (device < 0 || device >= n) return GPS_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return GPS_ECUDA;
  gps_ctx* ctx = new gps_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    delete ctx;
    return GPS_ECUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return GPS_OK;
}

int gps_set_stream(gps_ctx* ctx, void* cuda_stream) {
  if (!ctx) return GPS_EINVAL;
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return GPS_OK;
}

int gps_set_gemm_timing(gps_ctx* ctx, int on) {
  if (!ctx) return GPS_EINVAL;
  ctx->time_gemm = on != 0;
  ctx->gemm_events_used = 0;
  return GPS_OK;
}

void gps_destroy(gps_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->X, &ctx->y, &ctx->Kb, &ctx->Xb, &ctx->Sb, &ctx->vecs, &ctx->red, &ctx->params, &ctx->Gb, &ctx->fold_vecs, &ctx->descend_buf,
                    &ctx->fitc.V, &ctx->fitc.W, &ctx->fitc.rowv, &ctx->fitc.small, &ctx->fitc.part, &ctx->fitc.part2, &ctx->fitc.accf,
                    &ctx->fitc.acc1, &ctx->fitc.acc2, &ctx->fitc.acc3, &ctx->stage[0], &ctx->stage[1],
                    &ctx->stage[2], &ctx->stage[3]};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (ctx->d_info) cudaFree(ctx->d_info);
  if (ctx->d_tasks) cudaFree(ctx->d_tasks);
  if (ctx->d_tasks2) cudaFree(ctx->d_tasks2);
  if (ctx->d_tickets) cudaFree(ctx->d_tickets);
  for (auto& pr : ctx->trace) cudaEventDestroy(pr.second);
  for (auto& pr : ctx->gemm_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (cudaEvent_t e : {ctx->dss_fork, ctx->dss_join[0], ctx->dss_join[1], ctx->dss_join[2], ctx->dss_join[3]})
    if (e) cudaEventDestroy(e);
  for (gps_ctx* ln : ctx->fold_lanes) ctx->grid_lanes.push_back(ln);   // released with the grid lanes below
  for (gps_ctx* ln : ctx->grid_lanes) {
    cudaStream_t s = ln->own_stream;
    for (DevBuf* b : {&ln->X, &ln->y})
      if (b->p) cudaFree(b->p);
    gps_ctx_release(ln);
    if (s) cudaStreamDestroy(s);
  }
  gps_fitc_large_free(ctx);
  gps_fitc_fused_free(ctx);
  gps_comm_free(ctx);
  for (auto* v : {&ctx->potrf_events, &ctx->tile_events, &ctx->below_events, &ctx->trailA1_events})
    for (auto e : *v) cudaEventDestroy(e);
  if (ctx->panel2_stream) cudaStreamDestroy(ctx->panel2_stream);
  for (auto e : ctx->stage_ev) if (e) cudaEventDestroy(e);
  if (ctx->panel_stream) cudaStreamDestroy(ctx->panel_stream);
  if (ctx->trail_stream) cudaStreamDestroy(ctx->trail_stream);
  if (ctx->tri_stream) cudaStreamDestroy(ctx->tri_stream);
  for (cudaEvent_t e : {ctx->fork_ev, ctx->join_trail_ev, ctx->join_tri_ev})
    if (e) cudaEventDestroy(e);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* gps_last_error(gps_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t gps_launch_count(gps_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gps_last_gemm_ms(gps_ctx* ctx, double* ms, int64_t* launches) {
  if (!ctx) return GPS_EINVAL;
  if (ms) *ms = ctx->last_gemm_ms;
  if (launches) *launches = ctx->last_gemm_launches;
  return GPS_OK;
}

int gps_last_stage_ms(gps_ctx* ctx, double* ms7) {
  if (!ctx || !ms7) return GPS_EINVAL;
  if (!ctx->stage_valid) return gps_fail(ctx, GPS_ESTATE, "last_stage_ms: no CRPS/LOGS obj+grad evaluation to report");
  for (int k = 1; k < gps_ctx::ST_COUNT; ++k) ms7[k - 1] = ctx->last_stage_ms[k];
  return GPS_OK;
}

int gps_set_data(gps_ctx* ctx, const double* X, const double* y, int64_t N, int D) {
  if (!ctx) return GPS_EINVAL;
  if (!X || !y || N <= 0 || D <= 0 || D > 64) return gps_fail(ctx, GPS_EINVAL, "set_data: bad N=%lld D=%d", (long long)N, D);
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t Np = gps_pad(N);
  GPS_CHECK(gps_ensure(ctx, ctx->X, (size_t)Np * D));
  GPS_CHECK(gps_ensure(ctx, ctx->y, (size_t)Np));
  GPS_CUDA(cudaMemsetAsync(ctx->X.p, 0, (size_t)Np * D * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemsetAsync(ctx->y.p, 0, (size_t)Np * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(ctx->X.p, X, (size_t)N * D * sizeof(double), cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(ctx->y.p, y, (size_t)N * sizeof(double), cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->N = N;
  ctx->Np = Np;
  ctx->D = D;
  ctx->loo_valid = false;
  ctx->fitc.begun = false;
  ctx->fitc.pass2_done = false;
  ctx->fitc.fused = false;
  return GPS_OK;
}

// Everything of one full-GP evaluation that runs on the device, enqueued on the context's streams; the kernel
// parameters (e^a, e^c, 1/l) are already in ctx->params.  Leaves the objective in params[PAR_OBJ] and the raw
// gradient sums in params[PAR_GSUM ..] (the noise entry still lacks its chain-rule factor e^c).
static int full_eval_enqueue(gps_ctx* ctx, int score, bool want_grad, bool* stages_full) {
  const int64_t N = ctx->N, Np = ctx->Np;
  const int D = ctx->D;
  // objective only (crps / logs / nlml): alpha and diag K^-1 follow from L^-1 alone, the K^-1 = L^-T L^-1 stage is skipped
  const bool want_kinv = want_grad || score == GPS_DSS;
  GPS_CHECK(gps_factor_and_invert(ctx, score == GPS_NLML, want_kinv));
  double* v = ctx->vecs.p;
  double* par = ctx->params.p;
  *stages_full = false;
  ctx->stage_valid = false;
  if (!want_kinv)
    GPS_CHECK(gps_alpha_from_linv(ctx, ctx->Xb.p, Np, ctx->y.p, v + V_U * Np, v + V_ALPHA * Np,
                                  score == GPS_NLML ? nullptr : v + V_D * Np));
  if (score == GPS_DSS) {
    ctx->loo_valid = false;
    GPS_CHECK(gps_full_dss(ctx, par + PAR_OBJ, par + PAR_GSUM, want_grad));
  } else if (score != GPS_NLML) {
    if (want_kinv) GPS_CHECK(gps_diag_extract(ctx, ctx->Kb.p, Np, v + V_D * Np, 0));
    GPS_CHECK(gps_loo_score(ctx, score, N, Np, N, v + V_ALPHA * Np, v + V_D * Np, ctx->y.p, v + V_ABAR * Np,
                            v + V_DBAR * Np, v + V_LOOM * Np, v + V_LOOV * Np, par + PAR_OBJ));
    ctx->loo_valid = true;
    if (want_grad) {
      GPS_CHECK(gps_symv(ctx, ctx->Kb.p, Np, v + V_ABAR * Np, v + V_U * Np));
      GPS_CHECK(stage_mark(ctx, gps_ctx::ST_SCORE));
      GPS_CHECK(gps_symprod(ctx, ctx->Kb.p, v + V_DBAR * Np, ctx->Xb.p, ctx->Sb.p, Np));   // L^-1 in Xb is dead after LAUUM
      GPS_CHECK(stage_mark(ctx, gps_ctx::ST_SYMPROD));
      GPS_CHECK(gps_grad_contract(ctx, 0, ctx->Sb.p, N, Np, ctx->X.p, D, par, v + V_ALPHA * Np, v + V_U * Np,
                                  par + PAR_GSUM));
      GPS_CHECK(stage_mark(ctx, gps_ctx::ST_CONTRACT));
      *stages_full = true;
    }
  } else {
    ctx->loo_valid = false;
    GPS_CHECK(gps_nlml_value(ctx, N, Np, v + V_LOGD * Np, v + V_ALPHA * Np, ctx->y.p, par + PAR_OBJ));
    if (want_grad)
      GPS_CHECK(gps_grad_contract(ctx, 1, ctx->Kb.p, N, Np, ctx->X.p, D, par, v + V_ALPHA * Np, nullptr,
                                  par + PAR_GSUM));
  }
  return GPS_OK;
}

int gps_full_eval(gps_ctx* ctx, const double* theta, int score, double* obj, double* grad) {
  if (!ctx) return GPS_EINVAL;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "full_eval: call gps_set_data first");
  if (!theta || !obj || score < GPS_CRPS || score > GPS_DSS) return gps_fail(ctx, GPS_EINVAL, "full_eval: bad arguments (kc is a FITC objective)");
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t Np = ctx->Np;
  const int D = ctx->D;
  GPS_CHECK(gps_ensure_ws(ctx, Np));
  double ea, sn2;
  GPS_CHECK(gps_upload_params(ctx, theta, D, &ea, &sn2));
  gemm_timing_begin(ctx);
  bool stages_full = false;
  GPS_CHECK(full_eval_enqueue(ctx, score, grad != nullptr, &stages_full));
  double* par = ctx->params.p;
  double h[8 + 66];
  GPS_CUDA(cudaMemcpyAsync(h, par + PAR_OBJ, (size_t)(PAR_GSUM - PAR_OBJ + D + 2) * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CHECK(gps_check_info(ctx));  // synchronises the stream
  GPS_CHECK(gemm_timing_end(ctx));
  if (stages_full) {
    for (int k = 1; k < gps_ctx::ST_COUNT; ++k) {
      float ms = 0;
      GPS_CUDA(cudaEventElapsedTime(&ms, ctx->stage_ev[k - 1], ctx->stage_ev[k]));
      ctx->last_stage_ms[k] = ms;
    }
    ctx->stage_valid = true;
  }
  *obj = h[0];
  if (grad) {
    const double* gs = h + (PAR_GSUM - PAR_OBJ);
    grad[0] = gs[0];
    for (int d = 0; d < D; ++d) grad[1 + d] = gs[1 + d];
    grad[D + 1] = sn2 * gs[1 + D];
  }
  return GPS_OK;
}

// ---- the optimiser loop KF:237-260 with theta resident on the device ------------------------------------------------
// theta_dev[D + 2] -> params (e^a, e^c, 1/l_d): what gps_upload_params does on the host
__global__ void theta_params_kernel(const double* __restrict__ th, int D, double* __restrict__ par) {
  const int t = threadIdx.x;
  if (t == 0) par[0] = exp(th[0]);
  else if (t == 1) par[1] = exp(th[D + 1]);
  else if (t - 2 < D) par[t] = exp(-th[1 + (t - 2)]);
}
// theta -= lr * grad (noise entry with its e^c factor), trace[it] = objective before the step.  A failed
// factorisation (info != 0; the next iteration's factorisation clears the flag) freezes theta and is latched.
__global__ void full_update_kernel(double* __restrict__ th, const double* __restrict__ par, int D, double lr,
                                   double* __restrict__ trace, int it, const int* __restrict__ info, int* __restrict__ fail) {
  if (*info != 0 || fail[0] != 0) {
    if (threadIdx.x == 0 && fail[0] == 0) {
      fail[0] = *info;
      fail[1] = it;
    }
    return;
  }
  const int t = threadIdx.x;
  const double* gs = par + PAR_GSUM;
  if (t < D + 1) th[t] -= lr * gs[t];
  else if (t == D + 1) th[t] -= lr * par[1] * gs[t];
  if (t == 0) trace[it] = par[PAR_OBJ];
}

int gps_full_descend(gps_ctx* ctx, double* theta, int score, double lr_theta, int iters, double* obj_trace) {
  if (!ctx) return GPS_EINVAL;
  if (!theta || iters < 0) return gps_fail(ctx, GPS_EINVAL, "full_descend: bad arguments");
  const int P = ctx->D + 2;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "full_descend: call gps_set_data first");
  if (score < GPS_CRPS || score > GPS_DSS) return gps_fail(ctx, GPS_EINVAL, "full_descend: bad score");
  if (score == GPS_DSS || iters == 0) {
    // the 4-fold objective checks its fold factorisations on the host: one evaluation per round trip
    std::vector<double> g(P);
    for (int it = 0; it < iters; ++it) {
      double obj = 0.0;
      GPS_CHECK(gps_full_eval(ctx, theta, score, &obj, g.data()));
      if (obj_trace) obj_trace[it] = obj;
      for (int k = 0; k < P; ++k) theta[k] -= lr_theta * g[k];   // KF:254-257
    }
    return GPS_OK;
  }
  // theta stays on the device: per iteration the parameter kernel, the whole evaluation and the update kernel are
  // enqueued back to back; one synchronisation and read-back after the last iteration
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int D = ctx->D;
  GPS_CHECK(gps_ensure_ws(ctx, ctx->Np));
  GPS_CHECK(gps_ensure(ctx, ctx->descend_buf, (size_t)P + (size_t)iters + 4));
  double* d_th = ctx->descend_buf.p;
  double* d_trace = d_th + P;
  int* d_fail = reinterpret_cast<int*>(d_trace + iters);
  GPS_CUDA(cudaMemcpyAsync(d_th, theta, (size_t)P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GPS_CUDA(cudaMemsetAsync(d_fail, 0, 2 * sizeof(int), ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));   // theta may be pageable host memory
  ctx->gemm_events_used = 0;
  const bool timing = ctx->time_gemm;
  ctx->time_gemm = false;
  int rc = GPS_OK;
  for (int it = 0; it < iters && rc == GPS_OK; ++it) {
    theta_params_kernel<<<1, 96, 0, ctx->stream>>>(d_th, D, ctx->params.p);
    bool sf = false;
    rc = full_eval_enqueue(ctx, score, true, &sf);
    if (rc != GPS_OK) break;
    full_update_kernel<<<1, 96, 0, ctx->stream>>>(d_th, ctx->params.p, D, lr_theta, d_trace, it, ctx->d_info, d_fail);
    ctx->launches += 2;
  }
  ctx->time_gemm = timing;
  ctx->stage_valid = false;
  GPS_CHECK(rc);
  GPS_LAUNCH_CHECK();
  int fail[2] = {0, 0};
  GPS_CUDA(cudaMemcpyAsync(theta, d_th, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(fail, d_fail, sizeof fail, cudaMemcpyDeviceToHost, ctx->stream));
  if (obj_trace)
    GPS_CUDA(cudaMemcpyAsync(obj_trace, d_trace, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->loo_valid = false;
  if (fail[0] != 0)
    return gps_fail(ctx, GPS_ENOTPD, "full_descend: K + sn2 I not positive definite at pivot %d in iteration %d", fail[0], fail[1]);
  return GPS_OK;
}

int gps_fitc_descend(gps_ctx* ctx, double* theta, double* U, int M, double jitter, int score, double lr_theta,
                     double lr_u, int iters, double* obj_trace) {
  if (!ctx) return GPS_EINVAL;
  if (!theta || !U || M <= 0 || iters < 0) return gps_fail(ctx, GPS_EINVAL, "fitc_descend: bad arguments");
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "fitc_descend: call gps_set_data first");
  if (ctx->fitc_variant == 2 && M < ctx->fitc_large_min_m && gps_fitc_fused_supports(ctx, M, score))
    return gps_fitc_fused_descend(ctx, theta, U, M, jitter, score, lr_theta, lr_u, iters, obj_trace);   // device-resident loop
  const int P = ctx->D + 2, Q = M * ctx->D;
  std::vector<double> g(P), gU(Q);
  for (int it = 0; it < iters; ++it) {
    double obj = 0.0;
    GPS_CHECK(gps_fitc_eval(ctx, theta, U, M, jitter, score, &obj, g.data(), gU.data()));
    if (obj_trace) obj_trace[it] = obj;
    for (int k = 0; k < P; ++k) theta[k] -= lr_theta * g[k];   // K20:244-246
    for (int k = 0; k < Q; ++k) U[k] -= lr_u * gU[k];          // K20:247 / K20:350
  }
  return GPS_OK;
}

int gps_full_loo(gps_ctx* ctx, double* loo_mean, double* loo_var) {
  if (!ctx) return GPS_EINVAL;
  if (!ctx->loo_valid) return gps_fail(ctx, GPS_ESTATE, "full_loo: no CRPS/LOGS evaluation to report");
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t N = ctx->N, Np = ctx->Np;
  if (loo_mean)
    GPS_CUDA(cudaMemcpyAsync(loo_mean, ctx->vecs.p + V_LOOM * Np, N * sizeof(double), cudaMemcpyDefault, ctx->stream));
  if (loo_var)
    GPS_CUDA(cudaMemcpyAsync(loo_var, ctx->vecs.p + V_LOOV * Np, N * sizeof(double), cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

int gps_test_metrics(gps_ctx* ctx, const double* mean, const double* var, const double* y, int64_t n,
                     double ytm, double ytv, double* out) {
  if (!ctx) return GPS_EINVAL;
  if (!mean || !var || !y || !out || n <= 0) return gps_fail(ctx, GPS_EINVAL, "test_metrics: bad arguments");
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  const double *dm, *dv, *dy;
  GPS_CHECK(gps_stage_in(ctx, mean, n, ctx->stage[0], &dm));
  GPS_CHECK(gps_stage_in(ctx, var, n, ctx->stage[1], &dv));
  GPS_CHECK(gps_stage_in(ctx, y, n, ctx->stage[2], &dy));
  GPS_CHECK(gps_metrics_kernel(ctx, dm, dv, dy, n, ytm, ytv, ctx->params.p + PAR_OBJ));
  double s[6];
  GPS_CUDA(cudaMemcpyAsync(s, ctx->params.p + PAR_OBJ, sizeof s, cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  // out = {mse, smse, logs, crps, msll, coverage} followed by the six raw sums (for row-sharded callers)
  out[0] = s[0] / n;
  out[1] = s[0] / s[1];
  out[2] = s[2] / n;
  out[3] = s[3] / n;
  out[4] = (s[2] - s[4]) / n;
  out[5] = s[5] / n;
  for (int k = 0; k < 6; ++k) out[6 + k] = s[k];
  return GPS_OK;
}

int gps_score(gps_ctx* ctx, const double* m, const double* c, const double* y, int64_t n, int which,
              double* out) {
  if (!ctx) return GPS_EINVAL;
  if (!m || !c || !y || !out || n <= 0 || (which != GPS_CRPS && which != GPS_LOGS))
    return gps_fail(ctx, GPS_EINVAL, "score: bad arguments");
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  const double *dm, *dv, *dy;
  GPS_CHECK(gps_stage_in(ctx, m, n, ctx->stage[0], &dm));
  GPS_CHECK(gps_stage_in(ctx, c, n, ctx->stage[1], &dv));
  GPS_CHECK(gps_stage_in(ctx, y, n, ctx->stage[2], &dy));
  GPS_CHECK(gps_score_kernel(ctx, dm, dv, dy, n, which, ctx->params.p + PAR_OBJ));
  GPS_CUDA(cudaMemcpyAsync(out, ctx->params.p + PAR_OBJ, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

int gps_ard(gps_ctx* ctx, const double* x, int64_t n, const double* xp, int64_t m, int D, double a,
            const double* b, int nb, double* out) {
  if (!ctx) return GPS_EINVAL;
  if (!x || !xp || !b || !out || n < 0 || m < 0 || D <= 0 || D > 64 || (nb != 1 && nb != D))
    return gps_fail(ctx, GPS_EINVAL, "ard: bad arguments");
  if (n == 0 || m == 0) return GPS_OK;
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  double th[66];
  th[0] = a;
  for (int d = 0; d < D; ++d) th[1 + d] = b[nb == 1 ? 0 : d];
  th[D + 1] = 0.0;
  GPS_CHECK(gps_upload_params(ctx, th, D, nullptr, nullptr));
  const double *dx, *dxp;
  GPS_CHECK(gps_stage_in(ctx, x, (size_t)n * D, ctx->stage[0], &dx));
  GPS_CHECK(gps_stage_in(ctx, xp, (size_t)m * D, ctx->stage[1], &dxp));
  double* dout = out;
  const bool dev_out = gps_is_device_ptr(out);
  if (!dev_out) {
    GPS_CHECK(gps_ensure(ctx, ctx->stage[2], (size_t)n * m));
    dout = ctx->stage[2].p;
  }
  GPS_CHECK(gps_gram_rect(ctx, dx, n, dxp, m, D, ctx->params.p, dout, m));
  if (!dev_out)
    GPS_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * m * sizeof(double), cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

}  // extern "C"
