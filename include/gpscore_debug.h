/*
 * gpscore_debug.h — stage-level entry points of libgpscore.so used by the GPU unit tests and by
 * the micro-benchmarks that measure the FP64 roofline denominators.  Not part of the drop-in
 * boundary (see gpscore.h); kept exported so each dense stage can be checked on its own.
 */
#ifndef GPSCORE_DEBUG_H
#define GPSCORE_DEBUG_H
#include "gpscore.h"
#ifdef __cplusplus
extern "C" {
#endif

/* C[Mp,Np'] = alpha * op(A) op(B) + beta * C over all 128x128 tiles with the full k range.
 * kind 0: A[Mp,Kp] . B[Np',Kp]'   kind 1: A[Mp,Kp] . B[Kp,Np']   kind 2: A[Kp,Mp]' . B[Kp,Np'].
 * All dimensions multiples of 128, DEVICE pointers, dense row-major (ld = row length).
 * dvec (may be NULL, kind 0 only) scales the contraction index.  mirror (kind 2, square): also
 * write the transposed tile, lower tiles only. */
int gps_dbg_gemm(gps_ctx* ctx, int kind, const double* A, const double* B, double* C, int64_t Mp,
                 int64_t Npp, int64_t Kp, double alpha, double beta, const double* dvec, int mirror);

/* Factor the SPD matrix A[n,n] (UVA) with the blocked path; any of the outputs may be NULL:
 * L[n,n] (lower Cholesky factor), Linv[n,n] = L^-1, Ainv[n,n] = A^-1 (full symmetric). */
int gps_dbg_factor(gps_ctx* ctx, const double* A, int64_t n, double* L, double* Linv, double* Ainv);

/* Issue-rate micro-benchmarks on all SMs: DMMA m8n8k4 and DFMA, TFLOP/s each. */
int gps_dbg_fp64_peak(gps_ctx* ctx, int iters, double* dmma_tflops, double* dfma_tflops);

/* Tuning knobs for A/B measurements: what = 0 selects the tile-GEMM policy
 * (0: BK16 x 4 stages, 1: + fragment double-buffering, 2: BK32 x 3 stages, 3: BK32 + double-buffering,
 * 4: BK16 with 16 warps, 5: BK32 with 16 warps, 6: 64 x 128 CTA tile, 4 warps, two CTAs per SM (round-1 default),
 * 7: 6 + double-buffering, 8: 6 with 64 x 32 warp tiles, 9: the tile of 6 fed by TMA — cp.async.bulk.tensor.2d into a 4-stage
 * mbarrier ring, 128-byte hardware swizzle — = default since round 2; launches with a k-scaling vector and the row-strip
 * policies stay on cp.async); what = 1 selects the diagonal-block kernel of
 * POTRF (0: register-cyclic, 1: 32-blocked DMMA = default); what = 2 selects the FITC row passes
 * (0: thread-per-row, 1: tile/DMMA formulation = default); what = 3 sets the number of inducing points from
 * which gps_fitc_eval switches to the matrix form (default 33; lower it to A/B the two paths at M <= 32);
 * what = 4 selects the schedule of the full-GP factorisation: 1 = POTRF and the TRTRI merges overlapped on
 * priority streams (default), 0 = POTRF (with its look-ahead lanes) then TRTRI, 2 = every launch on the caller's
 * stream, one at a time (used to time launches alone); what = 6 arms the lane timeline (gps_dbg_trace);
 * what = 7 sets the row-strip height of the few-tile launches on POTRF's serial chain (16 = default, 32, 0 = the
 * normal policy); what = 8 switches the automatic strip policies for under-filled launches off (0) or on (1); what = 9 sets the share (per cent)
 * of a large TRTRI node's tiles that goes to its left child (50 = halving); what = 10 sets the row-strip height of the TRTRI
 * merges issued behind POTRF (0 = the normal policy, default; 32 / 16 measured slower); what = 11 / 12 run the overlapped TRTRI
 * merges / the POTRF trailing updates as persistent launches with that many CTAs (0 = off, default; measured slower);
 * what = 13: rows below POTRF's diagonal blocks updated left-looking inside a block column (1 = default) or by a rank-128
 * update per tile step (0, the round-1 order); what = 14 sets the POTRF outer block width in tiles (default 8); what = 15
 * releases the X phases of the inversion tree's right spine row group by row group (0 = default; 1 measured slower). */
int gps_dbg_set_variant(gps_ctx* ctx, int what, int value);

/* Timeline of the factorisation lanes of the last full-GP evaluation (arm with gps_dbg_set_variant(ctx, 6, 1)):
 * codes[i] = lane * 1000 + outer step (lane 1 diagonal-block chain, 2 rows below, 3 trailing update, 4 inversion
 * merges), ms[i] = time since the factorisation started.  Returns the number of entries (<= cap). */
int gps_dbg_trace(gps_ctx* ctx, int cap, int* codes, double* ms);

/* clock64 phase stamps of the last diagonal-block kernel launch (first call arms the recording):
 * cycles17[k] = cycles since kernel start at phase boundary k (load, 4 x {factor, panel, update},
 * sub-block inverses, off-diagonal inverse, write-back). */
int gps_dbg_potf2_phases(gps_ctx* ctx, int64_t* cycles17);

/* Launch floor of an evaluation call: `launches` empty kernels on the context's stream followed by one stream
 * synchronisation, best wall-clock microseconds of `reps` repetitions.  bench.py prints it as the roofline of the
 * latency-bound FITC M = 20, N = 10^4 point. */
int gps_dbg_launch_floor(gps_ctx* ctx, int launches, int reps, double* us);

/* Phase timeline of the last fused FITC evaluation (M <= 31): globaltimer nanoseconds, 16 slots per kernel
 * (pass 1 at 0, pass 2 at 16, pass 3 at 32): +0 kernel start, +1 preamble done, +2 row loop done, +3 CTA partial
 * formed (all by CTA 0), +4 total reduced (last CTA), +5 finishing step done (pass 3).  The first call arms the
 * recording and returns zeros. */
int gps_dbg_fused_phases(gps_ctx* ctx, int64_t* ns48);

/* Training Gram K = ARD(X,X) + sn2 I of the current data set at theta, N x N (UVA out). */
int gps_dbg_gram(gps_ctx* ctx, const double* theta, double* K);

#ifdef __cplusplus
}
#endif
#endif
