// Stage-level test hooks and FP64 issue-rate micro-benchmarks (include/gpscore_debug.h).
#include "../../include/gpscore_debug.h"
#include "gps_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
dmma_peak_kernel(int iters, double* out) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256)
dfma_peak_kernel(int iters, double* out) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace

extern "C" {

int gps_dbg_gemm(gps_ctx* ctx, int kind, const double* A, const double* B, double* C, int64_t Mp,
                 int64_t Npp, int64_t Kp, double alpha, double beta, const double* dvec, int mirror) {
  if (!ctx) return GPS_EINVAL;
  if (Mp % GPS_TILE || Npp % GPS_TILE || Kp % GPS_TILE) return gps_fail(ctx, GPS_EINVAL, "dbg_gemm: dims must be multiples of 128");
  GPS_CUDA(cudaSetDevice(ctx->device));
  std::vector<GemmTask> tasks;
  for (int ti = 0; ti < Mp / GPS_TILE; ++ti)
    for (int tj = 0; tj < Npp / GPS_TILE; ++tj) {
      if (mirror && tj > ti) continue;
      GemmTask t;
      t.a_row = ti * GPS_TILE; t.b_row = tj * GPS_TILE; t.k0 = 0; t.k1 = (int)Kp;
      t.c_row = ti * GPS_TILE; t.c_col = tj * GPS_TILE; t.flags = 0; t.pad = 0;
      tasks.push_back(t);
    }
  GPS_CHECK(gps_upload_tasks2(ctx, tasks));
  ctx->gemm_events_used = 0;
  const bool prev_timing = ctx->time_gemm;
  ctx->time_gemm = true;
  const int64_t lda = (kind == 2) ? Mp : Kp;
  const int64_t ldb = (kind == 0) ? Kp : Npp;
  const int rc = gps_gemm_tasks(ctx, kind, A, lda, B, ldb, C, Npp, alpha, beta, dvec, mirror != 0, ctx->d_tasks2,
                                tasks.size());
  ctx->time_gemm = prev_timing;
  GPS_CHECK(rc);
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  GPS_CUDA(cudaEventElapsedTime(&ms, ctx->gemm_events[0].first, ctx->gemm_events[0].second));
  ctx->last_gemm_ms = ms;
  ctx->last_gemm_launches = 1;
  return GPS_OK;
}

int gps_dbg_factor(gps_ctx* ctx, const double* A, int64_t n, double* L, double* Linv, double* Ainv) {
  if (!ctx) return GPS_EINVAL;
  if (!A || n <= 0) return gps_fail(ctx, GPS_EINVAL, "dbg_factor: bad arguments");
  GPS_CUDA(cudaSetDevice(ctx->device));
  const int64_t Np = gps_pad(n);
  GPS_CHECK(gps_ensure_ws(ctx, Np));
  ctx->loo_valid = false;
  ctx->gemm_events_used = 0;
  GPS_CUDA(cudaMemsetAsync(ctx->Kb.p, 0, (size_t)Np * Np * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemsetAsync(ctx->Xb.p, 0, (size_t)Np * Np * sizeof(double), ctx->stream));
  GPS_CUDA(cudaMemcpy2DAsync(ctx->Kb.p, Np * sizeof(double), A, n * sizeof(double), n * sizeof(double), n,
                             cudaMemcpyDefault, ctx->stream));
  std::vector<double> ones((size_t)(Np - n), 1.0);
  if (Np > n)
    GPS_CUDA(cudaMemcpy2DAsync(ctx->Kb.p + n * Np + n, (Np + 1) * sizeof(double), ones.data(), sizeof(double),
                               sizeof(double), Np - n, cudaMemcpyDefault, ctx->stream));
  GPS_CHECK(gps_potrf(ctx, ctx->Kb.p, ctx->Xb.p, Np));
  GPS_CHECK(gps_check_info(ctx));
  if (L)
    GPS_CUDA(cudaMemcpy2DAsync(L, n * sizeof(double), ctx->Kb.p, Np * sizeof(double), n * sizeof(double), n,
                               cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (Linv || Ainv) {
    GPS_CHECK(gps_trtri(ctx, ctx->Kb.p, ctx->Xb.p, ctx->Sb.p, Np));
    if (Linv)
      GPS_CUDA(cudaMemcpy2DAsync(Linv, n * sizeof(double), ctx->Xb.p, Np * sizeof(double), n * sizeof(double), n,
                                 cudaMemcpyDefault, ctx->stream));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (Ainv) {
    GPS_CHECK(gps_lauum(ctx, ctx->Xb.p, ctx->Kb.p, Np));
    GPS_CUDA(cudaMemcpy2DAsync(Ainv, n * sizeof(double), ctx->Kb.p, Np * sizeof(double), n * sizeof(double), n,
                               cudaMemcpyDefault, ctx->stream));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return GPS_OK;
}

int gps_dbg_fp64_peak(gps_ctx* ctx, int iters, double* dmma_tflops, double* dfma_tflops) {
  if (!ctx) return GPS_EINVAL;
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_ensure(ctx, ctx->params, PAR_LEN));
  const int blocks = ctx->sm_count * 4;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    GPS_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    dmma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(iters, ctx->params.p + 256);
    GPS_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    GPS_LAUNCH_CHECK();
    GPS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  }
  // per warp instruction: 8*8*4 FMA = 512 flop
  if (dmma_tflops) *dmma_tflops = (double)blocks * 8 /*warps*/ * 16.0 * iters * 512.0 / (ms * 1e-3) / 1e12;
  for (int rep = 0; rep < 2; ++rep) {
    GPS_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    dfma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(iters, ctx->params.p + 256);
    GPS_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    GPS_LAUNCH_CHECK();
    GPS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  }
  if (dfma_tflops) *dfma_tflops = (double)blocks * 256 * 16.0 * iters * 2.0 / (ms * 1e-3) / 1e12;
  ctx->launches += 4;
  return GPS_OK;
}

int gps_dbg_set_variant(gps_ctx* ctx, int what, int value) {
  if (!ctx) return GPS_EINVAL;
  if (what == 0) ctx->gemm_variant = value;
  else if (what == 1) ctx->potf2_variant = value;
  else if (what == 2) ctx->fitc_variant = value;
  else if (what == 3) ctx->fitc_large_min_m = value;
  else if (what == 4) ctx->overlap_trtri = value;
  else if (what == 6) ctx->trace_on = value != 0;
  else if (what == 7) ctx->chain_strip = value;
  else if (what == 8) ctx->gemm_auto_strip = value;
  else if (what == 10) ctx->tri_strip = value;
  else if (what == 11) ctx->cap_trtri = value;
  else if (what == 12) ctx->cap_trail = value;
  else if (what == 13) ctx->potrf_left = value;
  else if (what == 15) {
    ctx->trtri_rowwise = value;
    ctx->ws_Np = 0;
  }
  else if (what == 14) {
    if (value < 2 || value > 32) return gps_fail(ctx, GPS_EINVAL, "dbg_set_variant: outer block outside 2..32 tiles");
    ctx->potrf_ob = value;
    ctx->ws_Np = 0;   // task lists and events are rebuilt on the next evaluation
  }
  else if (what == 9) {
    if (value < 30 || value > 90) return gps_fail(ctx, GPS_EINVAL, "dbg_set_variant: split percentage outside 30..90");
    ctx->trtri_split_pct = value;
    ctx->ws_Np = 0;   // task lists are rebuilt on the next evaluation
  }
  else return gps_fail(ctx, GPS_EINVAL, "dbg_set_variant: unknown knob %d", what);
  return GPS_OK;
}

int gps_dbg_trace(gps_ctx* ctx, int cap, int* codes, double* ms) {
  if (!ctx || !codes || !ms) return GPS_EINVAL;
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  int n = 0;
  for (size_t i = 0; i < ctx->trace_used && n < cap; ++i, ++n) {
    float t = 0;
    GPS_CUDA(cudaEventSynchronize(ctx->trace[i].second));
    GPS_CUDA(cudaEventElapsedTime(&t, ctx->trace[0].second, ctx->trace[i].second));
    codes[n] = ctx->trace[i].first;
    ms[n] = t;
  }
  return n;
}

int gps_dbg_potf2_phases(gps_ctx* ctx, int64_t* cycles17) {
  if (!ctx || !cycles17) return GPS_EINVAL;
  GPS_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->potf2_prof) {
    GPS_CUDA(cudaMalloc(&ctx->potf2_prof, 32 * sizeof(long long)));
    GPS_CUDA(cudaMemset(ctx->potf2_prof, 0, 32 * sizeof(long long)));
    for (int k = 0; k < 17; ++k) cycles17[k] = 0;
    return GPS_OK;   // armed: the next factorisations record their phase stamps
  }
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  long long h[17];
  GPS_CUDA(cudaMemcpy(h, ctx->potf2_prof, sizeof h, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 17; ++k) cycles17[k] = h[k] - h[0];
  return GPS_OK;
}

int gps_dbg_launch_floor(gps_ctx* ctx, int launches, int reps, double* us) {
  if (!ctx || !us || launches < 0 || reps < 1) return GPS_EINVAL;
  return gps_launch_floor_us(ctx, launches, reps, us);
}

int gps_dbg_fused_phases(gps_ctx* ctx, int64_t* ns48) {
  if (!ctx || !ns48) return GPS_EINVAL;
  return gps_fitc_fused_phases(ctx, reinterpret_cast<long long*>(ns48));
}

int gps_dbg_gram(gps_ctx* ctx, const double* theta, double* K) {
  if (!ctx) return GPS_EINVAL;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "dbg_gram: no data");
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_ensure_ws(ctx, ctx->Np));
  GPS_CHECK(gps_upload_params(ctx, theta, ctx->D, nullptr, nullptr));
  GPS_CHECK(gps_gram_sym(ctx, ctx->X.p, ctx->N, ctx->Np, ctx->D, ctx->params.p, ctx->Kb.p));
  // lower tiles only were written: copy the lower triangle, caller mirrors
  GPS_CUDA(cudaMemcpy2DAsync(K, ctx->N * sizeof(double), ctx->Kb.p, ctx->Np * sizeof(double),
                             ctx->N * sizeof(double), ctx->N, cudaMemcpyDefault, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->loo_valid = false;
  return GPS_OK;
}

}  // extern "C"
