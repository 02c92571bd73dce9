/*
 * gpscore.h — C-ABI of libgpscore.so: the GP scoring-rule objective+gradient hot path on B200.
 *
 * The reference (polarlightman/Scoring-rules-for-Gaussian-process-regression-…) exposes no
 * FFI or plugin interface: the path is loop-body code inside four Python scripts
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds with ctypes
 * (INTEGRATION.md shows the stub); every entry point cites the reference statements it
 * replaces.  Abbreviations: KF = kin40k-FULL-compare.py, K20 = KIN40K-COMPARE-ALL-FITC-20.py,
 * SF/SC = the SIMPLE scripts, CP = contour-plot.R.
 *
 * Conventions
 *  - plain C types only; no torch types cross this boundary.
 *  - hyper-parameters are passed as theta = [a, b_1..b_D, c] with a = log sf^2 (KF:9),
 *    b_d = log l_d (KF:10-12, NOT log l^2) and c = log sn^2 (KF:239).  theta, inducing inputs,
 *    objective values and gradients are HOST pointers (they are a few doubles).
 *  - large arrays (X, y, test inputs, predictions, matrices of the element-wise entry points)
 *    are "UVA pointers": device memory or host memory, copied with cudaMemcpyDefault.
 *    All matrices are row-major float64.
 *  - every function returns 0 on success or a GPS_E* code; gps_last_error() gives the text.
 *    GPS_ENOTPD reports "matrix not positive definite" so that a host wrapper can raise
 *    RuntimeError the way torch.potrf did (the convention KF:726 / K20:784 rely on).
 *  - one host thread drives a context; work is enqueued on the context's stream and the
 *    call synchronises before returning host scalars.
 *  - there is no CPU fallback: gps_create fails with GPS_ENODEVICE without a CUDA device.
 */
#ifndef GPSCORE_H
#define GPSCORE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gps_ctx gps_ctx;

enum { GPS_OK = 0, GPS_EINVAL = 1, GPS_ECUDA = 2, GPS_ENOTPD = 3, GPS_ENODEVICE = 4, GPS_ENOMEM = 5,
       GPS_ESTATE = 6 };

/* score selector: LOO-CRPS (KF:245), LOO log score (KF:424), negative log marginal
 * likelihood (KF:331-334), 4-fold block-LOO DSS (KF:499-538, K20:538-582; needs 4 | N), 4-fold block CRPS "kc"
 * (K20:669-714; FITC only).  The FITC block objectives run through gps_fitc_eval (any M <= 4096) and, row-sharded,
 * through gps_fitc_eval_sharded. */
enum { GPS_CRPS = 0, GPS_LOGS = 1, GPS_NLML = 2, GPS_DSS = 3, GPS_KC = 4 };

/* ---- context ------------------------------------------------------------------------------- */
int gps_create(int device, gps_ctx** out);
void gps_destroy(gps_ctx* ctx);
const char* gps_last_error(gps_ctx* ctx);
/* version string of the library ("gpscore-b200 <n>") */
const char* gps_version(void);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t gps_launch_count(gps_ctx* ctx);
/* run on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the context's own; NULL restores the context's stream.  Lets the caller bracket calls
 * with its own CUDA events. */
int gps_set_stream(gps_ctx* ctx, void* cuda_stream);
/* switch the per-launch CUDA-event timing of the tile GEMMs on (1) or off (0, default) */
int gps_set_gemm_timing(gps_ctx* ctx, int on);
/* device milliseconds the dominant dense kernels (DMMA tile GEMM) spent inside the last
 * gps_full_eval, measured with CUDA events on the context's stream, and their launch count */
int gps_last_gemm_ms(gps_ctx* ctx, double* ms, int64_t* launches);
/* device milliseconds of the seven stages of the last CRPS / LOGS objective+gradient evaluation,
 * from CUDA events on the context's stream: ms7 = {Gram, POTRF, TRTRI, LAUUM, alpha+scores+u,
 * K^-1 diag(dbar) K^-1, gradient contraction} */
int gps_last_stage_ms(gps_ctx* ctx, double* ms7);

/* ---- training data (replaces train_x / train_y of KF:208-209, K20:198-199) ----------------- */
/* X[N,D], y[N] row-major UVA pointers; copied into context-owned, tile-padded buffers and all
 * N x N workspaces for the full GP are (re)allocated lazily on the first full-GP call. */
int gps_set_data(gps_ctx* ctx, const double* X, const double* y, int64_t N, int D);

/* ---- full GP: objective + gradient (replaces KF:239-252, KF:329-339, KF:416-428) ----------- */
/* theta[D+2] host; obj[1] host; grad[D+2] host (may be NULL: objective only). */
int gps_full_eval(gps_ctx* ctx, const double* theta, int score, double* obj, double* grad);
/* LOO predictive mean / variance of the last CRPS or LOGS evaluation (mean_term, cov_term of
 * KF:243-244); UVA pointers of length N. */
int gps_full_loo(gps_ctx* ctx, double* loo_mean, double* loo_var);
/* predictive mean and the diagonal of the predictive covariance at T test inputs
 * (replaces KF:267-273 → cal_mean_and_cov KF:121-126; only the diagonal is formed). */
int gps_full_predict(gps_ctx* ctx, const double* theta, const double* Xs, int64_t T, double* mean,
                     double* var);

/* ---- FITC: objective + gradient (replaces K20:222-236, K20:329-344, K20:434-452) ----------- */
/* Woodbury O(N M^2) evaluation of the dense big_Q path, in three row passes.  U[M,D] host.
 * grad_theta[D+2], grad_U[M*D] host.  For row-sharded multi-GPU runs the three passes are also
 * exposed separately so the host can all-reduce the packed accumulators between them
 * (acc buffers are DEVICE pointers owned by the caller; their lengths come from
 * gps_fitc_acc_len).  world_n is the global number of rows (the mean in KF:67 divides by it).
 * M <= 31, CRPS / LOGS / NLML: gps_fitc_eval is three fused kernels (preamble + row pass + reduction each) and one
 * stream synchronisation; DSS / KC and M = 32 use the staged row kernels.  32 < M <= 4096: the matrix form on the
 * tile-GEMM engine, all five scores (gps_fitc_predict works after any of them, gps_fitc_loo after CRPS / LOGS;
 * DSS / KC add, per fold, two M x M split-K products, one factorisation and one or two [M][N/4] products).  The staged
 * begin / pass1 / pass2 / pass3 / finish protocol (caller-side all-reduce of the accumulators) implements the
 * row-additive objectives only and rejects DSS / KC; new code should use gps_fitc_eval_sharded below. */
int gps_fitc_eval(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                  double* obj, double* grad_theta, double* grad_U);
int gps_fitc_acc_len(int M, int D, int64_t* len1, int64_t* len2, int64_t* len3);
int gps_fitc_begin(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                   int64_t world_n);
int gps_fitc_pass1(gps_ctx* ctx, double* acc1);           /* acc1 = [C - I (M*M) | v_y (M)]          */
int gps_fitc_pass2(gps_ctx* ctx, const double* acc1, double* acc2); /* acc2 = [R (M*M) | beta_bar (M) | obj] */
int gps_fitc_pass3(gps_ctx* ctx, const double* acc2, double* acc3); /* acc3 = [S | P = G [xs|1] | g_b rows (D) | sum lam_bar] */
int gps_fitc_finish(gps_ctx* ctx, const double* acc2, const double* acc3, double* obj, double* grad_theta,
                    double* grad_U);
/* ---- row-sharded FITC with the collective inside the library (FITC rows split over ranks, NCCL
 *      all-reduce of the M x M accumulators and gradient partials) ------------------------------------------------
 * One process per GPU.  Rank 0 obtains a 128-byte NCCL unique id (gps_comm_unique_id) and hands it to the other
 * ranks by any means (the Python layer broadcasts it over torch.distributed); every rank then calls
 * gps_comm_init on its context.  gps_fitc_eval_sharded evaluates ONE problem whose world_n rows are spread over
 * the ranks in contiguous blocks (this context's gps_set_data holds this rank's block): the three row passes
 * run with ncclAllReduce(sum, double) of the packed accumulators in between, all enqueued on the context's
 * stream — no host synchronisation before the final result read-back.  Every rank returns the same objective
 * and gradients.  CRPS / LOGS / NLML: M <= 31 runs the fused kernels, larger M the matrix form.  DSS / KC (4 | world_n):
 * the matrix form for every M; the folds are quarters of the GLOBAL row order, so the ranks' blocks must be
 * consecutive in rank order (the row counts are exchanged through the communicator and checked against world_n);
 * per evaluation two more all-reduces carry the folds' [P_f | g_f] (and for KC [E_f | hbar_f]) accumulators.
 * libnccl is bound with dlopen at the first gps_comm_* call (gps_comm_set_library names a specific copy). */
int gps_comm_set_library(const char* path);
int gps_comm_unique_id(void* out128);
int gps_comm_init(gps_ctx* ctx, const void* uid128, int rank, int world);
int gps_comm_info(gps_ctx* ctx, int* rank, int* world, int* nccl_version);
int gps_comm_destroy(gps_ctx* ctx);
/* Transport of the M <= 31 row-sharded evaluation.  1 (default when gps_comm_init could map the ranks' exchange
 * areas through cudaIpc): the all-reduce of each pass runs INSIDE the pass kernel — the CTA that completes the
 * rank's total stores it into every peer's exchange area over NVLink, publishes a sequence flag, waits for the
 * peers' flags and sums the slots in rank order — so a sharded evaluation is the same three launches as a single-GPU
 * one.  0: ncclAllReduce between the kernels.  *active (may be NULL) receives the transport in effect. */
int gps_comm_set_transport(gps_ctx* ctx, int transport, int* active);
int gps_comm_allreduce_sum(gps_ctx* ctx, double* buf, int64_t n);   /* in-place sum of a DEVICE buffer over the ranks */
int gps_fitc_eval_sharded(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                          int64_t world_n, double* obj, double* grad_theta, double* grad_U);
/* the loop K20:219-251 on a row-sharded problem (M <= 31, CRPS / LOGS / NLML): theta and U resident on every rank's
 * device, exchange inside the pass kernels (or NCCL between them), one synchronisation at the end */
int gps_fitc_descend_sharded(gps_ctx* ctx, double* theta, double* U, int M, double jitter, int score, int64_t world_n,
                             double lr_theta, double lr_u, int iters, double* obj_trace);
int gps_fitc_loo(gps_ctx* ctx, double* loo_mean, double* loo_var);
/* replaces K20:270-277 → spgp_cal_mean_and_cov K20:76-83 (diagonal only). Needs a finished
 * gps_fitc_eval / pass1+pass2 at the same theta, U (uses its L_A, L_C, beta). */
int gps_fitc_predict(gps_ctx* ctx, const double* Xs, int64_t T, double* mean, double* var);

/* ---- the scripts' optimiser loop in one call (replaces KF:237-260 / K20:219-251) ------------- */
/* `iters` steps of fixed-step gradient descent, theta -= lr_theta * grad (and U -= lr_u * grad_U for
 * FITC: the two learning rates of K20:326-327), without returning to the caller between steps.
 * theta[D+2] and U[M*D] are host arrays updated in place; obj_trace (host, may be NULL) receives
 * the objective before each step.  Stops early with GPS_ENOTPD if a factorisation fails.
 * gps_fitc_descend with M <= 31 and CRPS / LOGS / NLML keeps theta and U RESIDENT ON THE DEVICE: every iteration
 * is three evaluation kernels and one update kernel enqueued back to back, with a single synchronisation and
 * read-back after the last one (a failed factorisation freezes the parameters at that iteration). */
int gps_full_descend(gps_ctx* ctx, double* theta, int score, double lr_theta, int iters, double* obj_trace);
int gps_fitc_descend(gps_ctx* ctx, double* theta, double* U, int M, double jitter, int score, double lr_theta,
                     double lr_u, int iters, double* obj_trace);

/* ---- scoring of predictions (replaces KF:276-292: mse, SMSE KF:128-134, logs KF:52-57,
 *      crps KF:60-68, trivial_loss KF:110-119, coverage KF:288-292) ---------------------------- */
/* mean, var, y: UVA length n.  ytrain_mean / ytrain_var: mean and UNBIASED variance of the
 * training targets (KF:113-114).  out[12] host = {mse, smse, logs, crps, msll, coverage} followed by
 * the six raw sums they are formed from (what row-sharded callers all-reduce). */
int gps_test_metrics(gps_ctx* ctx, const double* mean, const double* var, const double* y, int64_t n,
                     double ytrain_mean, double ytrain_var, double* out);

/* ---- element-wise twins of the reference helpers (same argument meaning) -------------------- */
/* ARD(x, xp, a, b) KF:7-23: out[n,m] = e^a exp(-0.5 sum_d ((x_d - xp_d)/e^{b_d})^2).
 * b host, nb = 1 (broadcast, KF:8) or D. */
int gps_ard(gps_ctx* ctx, const double* x, int64_t n, const double* xp, int64_t m, int D, double a,
            const double* b, int nb, double* out);
/* chol_solve(B, A) KF:25-29: out[n,nrhs] = A^-1 B for SPD A[n,n]. */
int gps_chol_solve(gps_ctx* ctx, const double* B, const double* A, int64_t n, int64_t nrhs, double* out);
/* torch.mm twin for the products inside Q (KF:38), cal_mean_and_cov (KF:124-125) and spgp_cal_mean_and_cov
 * (K20:81-82): out[m,n] = A[m,k] B[k,n], UVA pointers, any sizes (padded internally to the 128 tile). */
int gps_matmul(gps_ctx* ctx, const double* A, const double* B, int64_t m, int64_t k, int64_t n, double* out);
/* crps(m, c, y) KF:60-68 / logs(m, c, y) KF:52-57: c is the VARIANCE. which = GPS_CRPS | GPS_LOGS. */
int gps_score(gps_ctx* ctx, const double* m, const double* c, const double* y, int64_t n, int which,
              double* out);

/* ---- hyper-parameter grid sweep (replaces CP:109-144: cal_NLML CP:68-73, cal_m_crps CP:43-53,
 *      wrong_cal_m_crps CP:55-64, cal_m_logs CP:75-85 over a (length scale, noise s.d.) grid) -- */
enum { GPS_GRID_NLML = 0, GPS_GRID_CRPS = 1, GPS_GRID_WRONG_CRPS = 2, GPS_GRID_LOGS = 3 };
/* 1-D inputs x[n], targets y[n] (UVA); ls[G], noise_sd[G] host: the G grid points in natural
 * parameters (k = 1 as in CP:44); out[G] host. */
int gps_grid_eval(gps_ctx* ctx, const double* x, const double* y, int n, const double* ls,
                  const double* noise_sd, int64_t G, int which, double* out);

#ifdef __cplusplus
}
#endif
#endif /* GPSCORE_H */
