// Shared declarations for libgpscore (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/gpscore.h"

#define GPS_TILE 128  // matrix tile edge: every N x N buffer is padded to a multiple of it
#define GPS_POTRF_OB 8  // POTRF outer block column = 8 tiles (k = 1024 trailing updates; 4, 12 and 16 measured slower)

struct GemmTask {  // one 128 x 128 output tile of a tile-GEMM launch
  int32_t a_row;   // first row (KC operand) / first column (MC operand) of the A panel
  int32_t b_row;   // same for B
  int32_t k0, k1;  // contraction range in elements, multiples of 16
  int32_t c_row, c_col;
  int32_t flags;   // GEMM_TRI_END / GEMM_TRI_BEGIN: the k-range ends / begins with a diagonal 128-block of a triangular A
  int32_t pad;
};
// A strip of rows [h, h + TM) of the tile then only needs k < k1 - 128 + h + TM (lower-triangular block last: the rest
// of the strip's rows in that block are explicit zeros) resp. k >= k0 + h (block of a transposed lower factor first).
// The skipped products are exact zeros, so results are bit-identical; it matters where a task has few k-tiles
// (matrix-form FITC with 1-8 tile rows).
enum { GEMM_TRI_END = 1, GEMM_TRI_BEGIN = 2 };

struct DevBuf {
  double* p = nullptr;
  size_t n = 0;  // doubles
};

struct gps_fitc_large;
struct gps_fitc_fused;
struct gps_comm;

struct gps_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;      // stream all work is enqueued on
  cudaStream_t own_stream = nullptr;  // the context's own stream
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;
  int fitc_variant = 2;               // 0: thread-per-row FITC passes, 1: tile (DMMA) formulation, 2: fused three-kernel path (gps_fitc_fused.cu)
  long long* potf2_prof = nullptr;    // device buffer for clock64 phase stamps of the diagonal kernel (debug)
  int potf2_variant = 1;              // 0: register-cyclic diagonal kernel, 1: 32-blocked DMMA diagonal kernel
  int gemm_variant = 9;               // tile-GEMM policy (see gps_gemm.cu): 9 = TMA-fed operand ring (default), 6 = the same tile with cp.async; switched by gps_dbg_set_variant
  int gemm_strip_policy = 0;          // 32 / 16: row-strip policy, set around the few-tile launches of POTRF's chain
  int chain_strip = 16;               // A/B knob 7: strip height used on the chain (0 = the normal policy)
  int gemm_auto_strip = 1;            // A/B knob 8: strip policies for every launch with too few tasks to fill the SMs
  // GEMM timing of the last full eval
  bool time_gemm = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
  size_t gemm_events_used = 0;
  // stage boundaries of the last full-GP evaluation (events on the main stream)
  enum { ST_BEGIN = 0, ST_GRAM, ST_POTRF, ST_TRTRI, ST_LAUUM, ST_SCORE, ST_SYMPROD, ST_CONTRACT, ST_COUNT };
  cudaEvent_t stage_ev[ST_COUNT] = {};
  double last_stage_ms[ST_COUNT] = {};
  bool stage_valid = false;
  double last_gemm_ms = 0;
  int64_t last_gemm_launches = 0;

  // ---- training data -------------------------------------------------------------------------
  int64_t N = 0;   // rows
  int64_t Np = 0;  // N padded to GPS_TILE
  int D = 0;
  DevBuf X;        // [Np, D] row-major, pad rows zero
  DevBuf y;        // [Np], pad zero

  // ---- full-GP workspaces (allocated lazily) -------------------------------------------------
  int64_t ws_Np = 0;
  DevBuf Kb;       // K -> L (lower, in place) -> K^-1 (full symmetric)
  DevBuf Xb;       // L^-1 (lower block triangle; diagonal blocks fully defined)
  DevBuf Sb;       // scratch for TRTRI, then S = K^-1 diag(dbar) K^-1 (lower tiles)
  DevBuf vecs;     // alpha, d, abar, dbar, u, loo_mean, loo_var, logdiag : 8 x Np
  DevBuf Gb;       // block-diagonal Gamma of the 4-fold DSS gradient (allocated on first use)
  DevBuf fold_vecs;
  std::vector<gps_ctx*> grid_lanes;   // lane contexts of the large-n grid sweep (own streams and workspaces)
  std::vector<gps_ctx*> fold_lanes;   // one lane context per DSS fold
  cudaEvent_t dss_fork = nullptr, dss_join[4] = {};
  DevBuf red;      // reduction scratch
  DevBuf descend_buf;   // device-resident theta | objective trace | failure latch of gps_full_descend
  DevBuf params;   // device copy of theta-derived parameters
  int* d_info = nullptr;       // device: first failing pivot (0 = ok)
  GemmTask* d_tasks = nullptr; // device task lists (cached per ws_Np)
  size_t tasks_cap = 0;
  GemmTask* d_tasks2 = nullptr; // per-call task lists (prediction, chol_solve)
  size_t tasks2_cap = 0;
  DevBuf stage[4];             // staging for host-resident arguments of the element-wise entry points
  std::vector<GemmTask> h_tasks;
  bool loo_valid = false;
  // cached task-list layout for ws_Np (offsets into d_tasks)
  struct Range { size_t off = 0, cnt = 0; };
  std::vector<Range> potrf_panel, potrf_inner, potrf_innerB, potrf_innerL, potrf_trailA, potrf_trailA1, potrf_trailB;
  cudaStream_t panel_stream = nullptr;          // high-priority stream for the POTRF look-ahead
  std::vector<cudaEvent_t> potrf_events, tile_events, below_events, trailA1_events;
  cudaStream_t panel2_stream = nullptr;         // POTRF panel work below the diagonal block (high priority)
  std::vector<Range> trtri_p, trtri_x;
  // TRTRI re-ordered for the overlapped driver: launches in issue order, each tagged with the POTRF
  // outer step after which its operands are final
  struct TriLaunch { int step; int phase; Range r; };
  std::vector<TriLaunch> trtri_sched;
  cudaStream_t trail_stream = nullptr;          // POTRF trailing updates (mid priority)
  cudaStream_t tri_stream = nullptr;            // overlapped TRTRI (lowest priority)
  cudaEvent_t fork_ev = nullptr, join_trail_ev = nullptr, join_tri_ev = nullptr;
  int overlap_trtri = 1;                        // 0: POTRF then TRTRI back to back (A/B knob)
  int gemm_grid_cap = 0;                        // > 0: TMA tile GEMMs larger than this many CTAs run persistent with that grid (slots left free for the chain)
  int cap_trtri = 0, cap_trail = 0;             // A/B knobs 11 / 12: grid caps of the overlapped TRTRI merges / the POTRF trailing updates
  int* d_tickets = nullptr;                     // dynamic-scheduling counters of the persistent launches (ring of 256)
  unsigned ticket_seq = 0;
  int trtri_rowwise = 0;                        // A/B knob 15: X phases of the inversion tree's right spine released row group by row group
  int potrf_ob = GPS_POTRF_OB;                  // POTRF outer block column in tiles (A/B knob 14)
  int potrf_left = 1;                           // A/B knob 13: rows below the diagonal block updated left-looking (one k <= 896 update per tile column instead of up to seven k = 128 updates; default since round 2: 66.2 -> 65.6 ms)
  int tri_strip = 0;                            // strip policy of the TRTRI merges issued behind POTRF (A/B knob 10; 0 = the normal policy)
  int trtri_split_pct = 50;                     // share of a large TRTRI node's tiles that goes to its left child (A/B knob 9)
  // debug timeline of the factorisation lanes (knob 6): (code, event) pairs, code = lane * 1000 + outer step
  bool trace_on = false;
  std::vector<std::pair<int, cudaEvent_t>> trace;
  size_t trace_used = 0;
  Range lauum, symprod;

  // ---- FITC state ----------------------------------------------------------------------------
  struct Fitc {
    int M = 0, MP = 0, score = 0;
    double jitter = 0, ea = 0, sn2 = 0;
    int64_t world_n = 0;
    int grid = 0;
    DevBuf V, W;          // [MP, N] (tile kernels) or [N, MP] (thread-per-row kernels)
    DevBuf rowv;          // per-row scalars: lam, lam_bar0, rbar, tbar, alpha, d : 6 x N
    DevBuf small;         // replicated small matrices (layout in gps_fitc.cu)
    DevBuf part;          // per-block partial accumulators
    DevBuf part2;         // first-stage sums of the partials (large grids)
    DevBuf acc1, acc2, acc3;  // single-GPU accumulators
    bool begun = false, pass2_done = false, tile = true, loo_ok = false;
    bool large = false;   // last evaluation ran the matrix form (gps_fitc_large.cu)
    bool fused = false;   // last evaluation ran the fused three-kernel path (gps_fitc_fused.cu)
    DevBuf accf;          // per-fold accumulators of the block objectives
    std::vector<double> host_out;
  } fitc;
  gps_fitc_large* fl = nullptr;   // matrix-form FITC state (M > 32)
  gps_fitc_fused* fu = nullptr;   // fused three-kernel FITC state (M <= 31)
  gps_comm* comm = nullptr;       // NCCL communicator of row-sharded runs (gps_comm.cu)
  int fitc_large_min_m = 33;      // M from which gps_fitc_eval uses the matrix form (debug knob)
};

int gps_fail(gps_ctx* c, int code, const char* fmt, ...);

#define GPS_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return gps_fail(ctx, GPS_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                      cudaGetErrorString(e__));                                                \
  } while (0)

#define GPS_CHECK(call)          \
  do {                           \
    int r__ = (call);            \
    if (r__ != GPS_OK) return r__; \
  } while (0)

#define GPS_LAUNCH_CHECK() GPS_CUDA(cudaGetLastError())

// function attributes (dynamic shared-memory limits) are per device: one flag per device ordinal
#define GPS_ONCE_PER_DEVICE(ctx)        \
  static bool once__[64] = {};          \
  bool& configured = once__[(ctx)->device & 63]

int gps_ensure(gps_ctx* ctx, DevBuf& b, size_t n);
bool gps_is_device_ptr(const void* p);
int gps_stage_in(gps_ctx* ctx, const double* p, size_t n, DevBuf& tmp, const double** out);
int gps_ensure_ws(gps_ctx* ctx, int64_t Np);
int gps_upload_params(gps_ctx* ctx, const double* theta, int D, double* ea_out, double* sn2_out);
int gps_upload_tasks2(gps_ctx* ctx, const std::vector<GemmTask>& h);
int gps_factor_and_invert(gps_ctx* ctx, bool want_logdet, bool want_kinv = true);
// gps_predict.cu
int gps_alpha_from_linv(gps_ctx* ctx, const double* Xinv, int64_t Np, const double* y, double* u, double* alpha, double* d);
int gps_full_dss(gps_ctx* ctx, double* par_obj, double* par_gsum, bool want_grad);
void gps_ctx_release(gps_ctx* child);
// gps_fitc_large.cu
void gps_fitc_large_free(gps_ctx* ctx);
int gps_fitc_large_eval(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                        double* obj, double* grad_theta, double* grad_U);
int gps_fitc_large_predict(gps_ctx* ctx, const double* dXs, int64_t T, double* dm, double* dv);
int gps_fitc_large_acc_len(int M, int D, int64_t* len1, int64_t* len2, int64_t* len3);
int gps_fitc_large_begin(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                         int64_t world_n);
int gps_fitc_large_pass1(gps_ctx* ctx, double* acc1);
int gps_fitc_large_pass2(gps_ctx* ctx, const double* acc1, double* acc2, bool want_grad);
int gps_fitc_large_pass3(gps_ctx* ctx, const double* acc2, double* acc3);
int gps_fitc_large_finish(gps_ctx* ctx, const double* acc2, const double* acc3, double* obj, double* grad_theta,
                          double* grad_U);
// gps_fitc_fused.cu
typedef int (*gps_allreduce_fn)(gps_ctx* ctx, double* buf, size_t n);   // in-place sum over the ranks, on ctx->stream
bool gps_fitc_fused_supports(const gps_ctx* ctx, int M, int score);
void gps_fitc_fused_free(gps_ctx* ctx);
int gps_fitc_fused_eval(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                        int64_t world_n, gps_allreduce_fn allreduce, double* obj, double* grad_theta, double* grad_U);
int gps_fitc_fused_descend(gps_ctx* ctx, double* theta, double* U, int M, double jitter, int score, double lr_theta,
                           double lr_u, int iters, double* obj_trace, int64_t world_n = 0, gps_allreduce_fn allreduce = nullptr);
int gps_fitc_fused_loo(gps_ctx* ctx, double* dm, double* dv);
int gps_fitc_fused_predict(gps_ctx* ctx, const double* dXs, int64_t T, double* dm, double* dv);
int gps_launch_floor_us(gps_ctx* ctx, int launches, int reps, double* us);
int gps_fitc_fused_phases(gps_ctx* ctx, long long* out48);
int gps_fitc_large_eval_sharded(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                                int64_t world_n, int64_t row_offset, gps_allreduce_fn allreduce, double* obj,
                                double* grad_theta, double* grad_U);
// gps_comm.cu
void gps_comm_free(gps_ctx* ctx);
int gps_comm_allreduce(gps_ctx* ctx, double* buf, size_t n);
// peer-memory exchange area of the fused FITC kernels (one-shot all-reduce inside the pass kernels over NVLink):
// every rank owns [2 parities][world slots][GPS_P2P_SLOT doubles] followed by [2][world] sequence flags, and holds
// peer mappings (cudaIpc) of all the others.  peers == nullptr: not available, the NCCL path is used.
constexpr int GPS_P2P_SLOT = 1664;   // >= the largest packed accumulator of the fused path (M = 31, D = 16: 1585)
struct gps_p2p_view {
  double** peers;            // device array [world] of exchange-area base pointers (own entry = local)
  int rank, world;
  unsigned long long seq;    // sequence number of the NEXT exchange (the host advances it per launch)
};
bool gps_comm_p2p_view(gps_ctx* ctx, gps_p2p_view* v, int exchanges);
// offsets (doubles) into ctx->params and rows of ctx->vecs
constexpr int PAR_OBJ = 128;
constexpr int PAR_GSUM = 136;
constexpr int PAR_LEN = 512;
enum { V_ALPHA = 0, V_D, V_ABAR, V_DBAR, V_U, V_LOOM, V_LOOV, V_LOGD, V_COUNT };

static inline int64_t gps_pad(int64_t n) { return (n + GPS_TILE - 1) / GPS_TILE * GPS_TILE; }

// ---- device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0. `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (w == 0) {
    v = lane < nw ? sh[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// ---- modules -------------------------------------------------------------------------------------
// gps_gram.cu
int gps_gram_sym(gps_ctx* ctx, const double* X, int64_t N, int64_t Np, int D, const double* d_par, double* K);
int gps_gram_rect(gps_ctx* ctx, const double* x, int64_t n, const double* xp, int64_t m, int D,
                  const double* d_par, double* out, int64_t ldo);
// gps_gemm.cu
enum { GEMM_KC_KC = 0, GEMM_KC_MC = 1, GEMM_MC_MC = 2 };
int gps_gemm_tasks(gps_ctx* ctx, int kind, const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* C, int64_t ldc, double alpha, double beta, const double* dvec, bool mirror,
                   const GemmTask* d_tasks, size_t ntasks);
// gps_chol.cu
int gps_build_tasks(gps_ctx* ctx, int64_t Np);
int gps_potrf(gps_ctx* ctx, double* K, double* Xinv, int64_t Np);       // K -> L in place; diag-block inverses -> Xinv
int gps_trtri(gps_ctx* ctx, const double* L, double* Xinv, double* scratch, int64_t Np);
// POTRF with the TRTRI merges issued on a low-priority stream as soon as their block columns are final
int gps_potrf_trtri(gps_ctx* ctx, double* K, double* Xinv, double* scratch, int64_t Np);
int gps_lauum(gps_ctx* ctx, const double* Xinv, double* Kinv, int64_t Np);
int gps_symprod(gps_ctx* ctx, const double* Kinv, const double* dvec, double* scratch, double* S, int64_t Np);
int gps_check_info(gps_ctx* ctx);
// gps_score.cu
int gps_symv(gps_ctx* ctx, const double* A, int64_t Np, const double* x, double* y);
int gps_diag_extract(gps_ctx* ctx, const double* A, int64_t Np, double* d, int do_log);
int gps_loo_score(gps_ctx* ctx, int score, int64_t N, int64_t Np, int64_t norm_n, const double* alpha, const double* d,
                  const double* y, double* abar, double* dbar, double* loo_mean, double* loo_var,
                  double* obj_dev);
int gps_nlml_value(gps_ctx* ctx, int64_t N, int64_t Np, const double* logdiag, const double* alpha,
                   const double* y, double* obj_dev);
int gps_grad_contract(gps_ctx* ctx, int mode, const double* Mx, int64_t N, int64_t Np, const double* X, int D,
                      const double* d_par, const double* alpha, const double* u, double* out_dev);
int gps_metrics_kernel(gps_ctx* ctx, const double* mean, const double* var, const double* y, int64_t n,
                       double ytm, double ytv, double* out_dev);
int gps_score_kernel(gps_ctx* ctx, const double* m, const double* c, const double* y, int64_t n, int which,
                     double* out_dev);
