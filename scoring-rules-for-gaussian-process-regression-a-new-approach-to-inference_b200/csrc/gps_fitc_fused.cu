// FITC objective + gradient for M <= 31 inducing points: the three Woodbury row passes (SURVEY.md App. A.2,
// K20:222-251 / K20:329-354 / K20:434-462) as THREE kernels and nothing else.  Each kernel
//
//   preamble   the replicated M x M algebra of its stage, redone by every CTA in shared memory (Cholesky and
//              triangular inverse by one warp with rows in registers, every M x M x M product on DMMA)
//   row pass   a persistent loop over 32-row (pass 3: 16-row) warp tiles, entirely in registers: the
//              m8n8k4 accumulator layout of  Out' = Mat . K'  ("m on the lane group, rows on the thread
//              in the group") IS the A/B operand layout of the outer-product accumulations
//              sum_rows Out' Out, so V, W, K_uf_bar never go through shared memory or HBM
//   reduction  per-CTA partials -> groups of 8 CTAs -> total, by the last CTA to arrive (tickets), in a
//              fixed order: deterministic, no separate reduce launches
//
// and the last CTA of pass 3 also runs the finishing step (Cholesky adjoint of L_A, all D + 2 + M D gradients)
// and writes the result straight into mapped host memory: one evaluation = 3 launches + 1 stream
// synchronisation, no memcpy.  The row-sharded multi-GPU evaluation runs the same three kernels with an
// all-reduce of the packed accumulator after each (NCCL on the context's stream, gps_comm.cu) and the
// finishing step as a fourth, one-CTA kernel.
//
// What a row costs.  Pass 1 computes k_i = K_uf[:, i] once (M exps) and stores it ([M][ldk], 8 M B/row); passes 2
// and 3 re-read it.  Pass 3 works in "k-space" (prototype + derivation: oracle/woodbury.py::fitc_obj_grad_kspace):
//   Kuf_bar_i = t_i a1 + (y_i/lam_i) a2 + [2 r_i E1 + (2/lam_i) E2 - 2 lam_bar_i E3] k_i
// with E1 = T2'T2, E2 = L_A^-T C_bar L_A^-1, E3 = A^-1 formed once per evaluation, so a row needs three
// M x M products instead of five and the only M x M row reduction of the pass is the symmetric
// Z = sum lam_bar_i k_i k_i'.  DMMA count per 8 rows at M = 20: 23 + 23 + 63.
//
// "Extra slot" trick: the operand matrices have 8 * MT >= M + 1 rows; row M carries a vector, so a
// dot product with every row rides along in the tile product for free (pass 1: y_i / sqrt(lam_i) -> v_y,
// pass 2: c2 = T2' beta -> W_i . beta, and t_i -> beta_bar in the accumulation).
#include <math.h>
#include <string.h>

#include "gps_common.cuh"
#include "gps_exp.cuh"
#include "gps_fitc_small.cuh"

namespace {

constexpr int FW = 4;          // warps per CTA
constexpr int FT = 32 * FW;    // threads per CTA
constexpr int TR = 32;         // rows per warp tile
constexpr int GRP = 8;         // CTAs per first-level reduction group
constexpr double F_INV_SQRT_PI = 0.56418958354775628695;
constexpr double F_INV_SQRT_2PI = 0.39894228040143267794;
constexpr double F_INV_SQRT2 = 0.70710678118654752440;
constexpr double F_HALF_LOG_2PI = 0.91893853320467274178;

__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, bool ok) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = ok ? 8 : 0;                                   // src-size 0: the 8 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit_group();
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
// K_uf tile of `rows` (16 or 32) data rows starting at r0 into smem [KP][rows + 4] (rows beyond N hold whatever the
// padded buffer holds: every use is masked).  16-byte chunks: r0 and ldk are multiples of 8.
template <int ROWS>
__device__ __forceinline__ void fetch_k_tile(double* kt, const double* __restrict__ Kst, int64_t ldk, int64_t r0, int64_t N,
                                             int M, int lane) {
  constexpr int CPR = ROWS / 2;                               // chunks per m-row
  if (r0 < N) {
    for (int c = lane; c < M * CPR; c += 32) {
      const int m = c / CPR, q = c - m * CPR;
      cp_async16(kt + m * (ROWS + 4) + 2 * q, Kst + (int64_t)m * ldk + r0 + 2 * q);
    }
  }
}
// `cnt` doubles (a multiple of 2) starting at src (16-byte aligned) into dst; chunks dealt to the lanes from `lane0`
__device__ __forceinline__ void fetch_span(double* dst, const double* __restrict__ src, int cnt, bool on, int lane, int lane0) {
  if (on)
    for (int c = (lane - lane0) & 31; 2 * c < cnt; c += 32) cp_async16(dst + 2 * c, src + 2 * c);
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPEND)); }

__host__ __device__ constexpr int lda_of(int kp) {   // operand stride: conflict-free half-warp fragment loads
  int l = kp;
  while ((l & 15) != 4 && (l & 15) != 12) ++l;
  return l;
}

template <int MT_, int KS_, int DT_>
struct FCfg {
  static constexpr int MT = MT_, KS = KS_, DT = DT_;
  static constexpr int MP = 8 * MT_, KP = 4 * KS_, LDA = lda_of(4 * KS_), LDU = DT_ + 1, NXT = DT_ / 8;
  static constexpr int NTRI = MT_ * (MT_ + 1) / 2;
  static_assert(MP <= 32, "one-warp Cholesky: at most 32 rows");
};

// replicated small matrices handed from kernel to kernel through global memory (doubles, MP-strided)
struct FL {
  int par, us, kuu, la, lainv, lc, lcinv, beta, t2, c2, bbar, vyb, cbar, total;
  __host__ __device__ FL(int MP, int D) {
    int o = 0;
    par = o;   o += 24;
    us = o;    o += MP * 16;
    kuu = o;   o += MP * MP;
    la = o;    o += MP * MP;
    lainv = o; o += MP * MP;
    lc = o;    o += MP * MP;
    lcinv = o; o += MP * MP;
    beta = o;  o += MP;
    t2 = o;    o += MP * MP;
    c2 = o;    o += MP;
    bbar = o;  o += MP;
    vyb = o;   o += MP;
    cbar = o;  o += MP * MP;
    total = o;
    (void)D;
  }
};

struct FusedArgs {
  const double* X;        // [N][D]
  const double* y;        // [N]
  const double* thU;      // theta[D+2] | U[M*D]  (device or mapped host memory)
  double* fs;             // replicated small matrices (FL)
  double* Kst;            // [M][ldk]  K_uf
  double* rowv;           // [6][ldk]  lam, lb0, rbar, tbar, alpha, d
  double* part;           // [G][len]        per-CTA partials
  double* gpart;          // [G/GRP + 1][len] group partials
  int* cnt;               // tickets of this kernel: [0] total, [1 + grp] groups
  double* acc1;           // [MP*MP]            C - I (lower tiles) with v_y in row M
  double* acc2;           // [MP*MP + 1]        R (lower tiles) with beta_bar in row M | objective share
  double* acc3;           // [MP*MP + MP*DT + MP + DT + 1]  Z | P | S0 | xq | sum lam_bar
  double* out_dev;        // [2 + D + 2 + M*D]  obj | g_theta | g_U | info
  double* out_host;       // same, mapped host memory (may be null)
  int* info;
  long long* prof;        // debug: globaltimer stamps of the phases (null = off), 16 slots per kernel
  double** peers;         // row-sharded runs: exchange areas of all ranks (peer memory), null = single GPU / NCCL
  int rank, world;
  unsigned long long seq; // sequence number of this kernel's exchange
  int64_t N, ldk;
  int D, M, score, finish;   // finish: pass 3 also runs the finishing step (single GPU)
  double jitter, invN, world_n;
};

// pass 1 takes theta | U BY VALUE in its launch parameters when they fit (D + 2 + M D <= THU_INLINE doubles; 170 at
// M = 20, D = 8): no staging copy, no extra stream operation in front of the evaluation
constexpr int THU_INLINE = 400;
struct FusedArgsP1 {
  FusedArgs a;
  int inline_thu;
  double thu[THU_INLINE];
};

__device__ __forceinline__ void stamp(const FusedArgs& a, int slot) {
  if (a.prof) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.prof[slot] = t;
  }
}

__host__ __device__ inline int len1_of(int MP) { return MP * MP; }
__host__ __device__ inline int len2_of(int MP) { return MP * MP + 1; }
__host__ __device__ inline int len3_of(int MP, int DT) { return MP * MP + MP * DT + MP + DT + 1; }

// ---- small dense products on DMMA: C[i][j] = alpha * sum_k opA(i,k) opB(k,j) + cbeta * C[i][j] --------------------
// A, B: MP x MP in shared memory, tight stride MP.  Output tiles are dealt round-robin to the CTA's warps.
// __noinline__: the preambles and the finishing step call it ~25 times; one shared copy stays in the instruction
// cache, whereas 25 inlined copies are code that runs once (those phases are instruction-fetch bound)
template <int MT, bool TA, bool TB>
__device__ __noinline__ void smm(const double* A, const double* B, double* C, int ldc, int ncols, double alpha,
                                 double cbeta, int warp, int lane) {
  constexpr int MP = 8 * MT;
  constexpr int TPW = (MT * MT + FW - 1) / FW;          // output tiles per warp: their DMMA chains are interleaved
  const int g = lane >> 2, t = lane & 3;
  double c[TPW][2];
  int ti[TPW], tj[TPW];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int idx = warp + q * FW;
    ti[q] = idx < MT * MT ? idx / MT : 0;
    tj[q] = idx < MT * MT ? idx - (idx / MT) * MT : 0;
    c[q][0] = c[q][1] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < 2 * MT; ++s) {
    const int k = 4 * s + t;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const double a = TA ? A[k * MP + 8 * ti[q] + g] : A[(8 * ti[q] + g) * MP + k];
      const double b = TB ? B[(8 * tj[q] + g) * MP + k] : B[k * MP + 8 * tj[q] + g];
      dmma(c[q][0], c[q][1], a, b);
    }
  }
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    if (warp + q * FW < MT * MT) {
      const int r = 8 * ti[q] + g, cc = 8 * tj[q] + 2 * t;
      if (cc < ncols) C[r * ldc + cc] = alpha * c[q][0] + (cbeta != 0.0 ? cbeta * C[r * ldc + cc] : 0.0);
      if (cc + 1 < ncols) C[r * ldc + cc + 1] = alpha * c[q][1] + (cbeta != 0.0 ? cbeta * C[r * ldc + cc + 1] : 0.0);
    }
  }
}

// y[i] = sum_k op(A)(i,k) x[k]   (threads < MP)
template <int MP, bool TA>
__device__ __noinline__ void smv(const double* A, const double* x, double* y, int tid) {
  if (tid < MP) {
    double s = 0.0;
#pragma unroll 8
    for (int k = 0; k < MP; ++k) s = fma(TA ? A[k * MP + tid] : A[tid * MP + k], x[k], s);
    y[tid] = s;
  }
}

// Cholesky adjoint with explicit inverse: Abar = 1/2 L^-T (P + P') L^-1, P = Phi(L' Lbar).  Lbar is overwritten (Z),
// T is scratch, result in Abar.
template <int MT>
__device__ __forceinline__ void chol_adjoint_smm(const double* L, const double* Linv, double* Lbar, double* T,
                                                 double* Abar, int tid, int warp, int lane) {
  constexpr int MP = 8 * MT;
  smm<MT, true, false>(L, Lbar, T, MP, MP, 1.0, 0.0, warp, lane);   // T = L' Lbar
  __syncthreads();
  for (int e = tid; e < MP * MP; e += FT) {
    const int r = e / MP, c = e - r * MP;
    const double lo = (r > c) ? T[e] : ((r == c) ? 0.5 * T[e] : 0.0);
    const double up = (c > r) ? T[c * MP + r] : ((r == c) ? 0.5 * T[e] : 0.0);
    Lbar[e] = lo + up;                                              // Z = P + P'
  }
  __syncthreads();
  smm<MT, false, false>(Lbar, Linv, T, MP, MP, 1.0, 0.0, warp, lane);   // T = Z L^-1
  __syncthreads();
  smm<MT, true, false>(Linv, T, Abar, MP, MP, 0.5, 0.0, warp, lane);    // Abar = 1/2 L^-T T
  __syncthreads();
}

// (A loop-over-shared-memory Cholesky / triangular inverse of ~60 instructions was tried in place of the
// register-resident, fully unrolled ones of gps_fitc_small.cuh, to cut the instruction-fetch cost of code that runs
// once per kernel: its dependent LDS -> DFMA -> STS chains made the preambles 2.6x SLOWER (7.5 -> 19.7 us) — dropped.)

// ---- reduce-scatter over the 8 lane groups (lane bits 2..4) ----------------------------------------------------
// v[2 gi + e] holds this lane's share for tile row 8 gi + 2 t + e; on return lane (g, t) holds the full sum of
// v[g], i.e. it OWNS tile row  8 (g >> 1) + 2 t + (g & 1).
__device__ __forceinline__ double rs8(const double (&v)[8], int lane) {
  double w[4], u[2];
  bool hi = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double send = hi ? v[k] : v[k + 4], keep = hi ? v[k + 4] : v[k];
    w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  hi = (lane & 8) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double send = hi ? w[k] : w[k + 2], keep = hi ? w[k + 2] : w[k];
    u[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  hi = (lane & 4) != 0;
  const double send = hi ? u[0] : u[1], keep = hi ? u[1] : u[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}
// 4 values (16-row tiles): lanes g and g ^ 1 both end with the full sum of v[g >> 1]; owned tile row
// 8 (g >> 2) + 2 t + ((g >> 1) & 1).
__device__ __forceinline__ double rs4(const double (&v)[4], int lane) {
  double u[2];
  bool hi = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double send = hi ? v[k] : v[k + 2], keep = hi ? v[k + 2] : v[k];
    u[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  hi = (lane & 8) != 0;
  const double send = hi ? u[0] : u[1], keep = hi ? u[1] : u[0];
  double r = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  return r;
}

// Out'[m][row] = sum_k Aop[m][k] K[row][k] for one 8-row group: kb[s] = K[row g][4 s + t] (B operand);
// out[i][e] = Out[row 2 t + e][m = 8 i + g].  TRI: Aop is lower triangular except its last 8-row tile.
template <class C, bool TRI>
__device__ __forceinline__ void tile_product(const double* __restrict__ Aop, const double (&kb)[C::KS], int g, int t,
                                             double (&out)[C::MT][2]) {
#pragma unroll
  for (int i = 0; i < C::MT; ++i) {
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int s = 0; s < C::KS; ++s) {
      if (!TRI || i == C::MT - 1 || s <= 2 * i + 1) dmma(c0, c1, Aop[(8 * i + g) * C::LDA + 4 * s + t], kb[s]);
    }
    out[i][0] = c0;
    out[i][1] = c1;
  }
}

// acc[tri(i,j)] += sum over the group's rows of a[i][.] b[j][.]   (lower tiles i >= j)
template <class C>
__device__ __forceinline__ void tri_accumulate(double (&acc)[C::NTRI][2], const double (&a)[C::MT][2],
                                               const double (&b)[C::MT][2]) {
  int idx = 0;
#pragma unroll
  for (int i = 0; i < C::MT; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      dmma(acc[idx][0], acc[idx][1], a[i][0], b[j][0]);
      dmma(acc[idx][0], acc[idx][1], a[i][1], b[j][1]);
      ++idx;
    }
}

// ---- CTA partial -> group partial -> total, by the last CTA to arrive at each level --------------------------------
// Returns true in the CTA that completed the total (its `acc` is final and visible to that CTA).  Partials are
// stored with an even stride and summed with 16-byte loads, all loads of a step in flight before the first add
// (the sums are latency-bound: each level is store -> fence -> ticket -> fence -> load).  Groups are 8 CTAs, 16 on
// grids beyond 128 CTAs; the order of every sum is fixed by the grid, so results are reproducible bit for bit.
__device__ __forceinline__ int group_size(int G) { return G > 128 ? 2 * GRP : GRP; }

__device__ __forceinline__ bool ticket_reduce(const double* __restrict__ cta_vals_smem, int len, double* part,
                                              double* gpart, int* cnt, double* acc, int tid) {
  __shared__ int s_flag;
  const int b = blockIdx.x, G = gridDim.x;
  const int gs = group_size(G);
  const int grp = b / gs, ngrp = (G + gs - 1) / gs;
  const int gsz = min(gs, G - grp * gs);
  const int lenp = (len + 1) & ~1, nv = lenp >> 1;          // even stride: 16-byte aligned rows
  for (int e = tid; e < lenp; e += FT) part[(int64_t)b * lenp + e] = (e < len) ? cta_vals_smem[e] : 0.0;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(&cnt[1 + grp], 1) == gsz - 1);
  __syncthreads();
  if (!s_flag) return false;
  __threadfence();
  {
    const double2* src = reinterpret_cast<const double2*>(part + (int64_t)grp * gs * lenp);
    double2* dst = reinterpret_cast<double2*>(gpart + (int64_t)grp * lenp);
    for (int e = tid; e < nv; e += FT) {
      double2 v[2 * GRP];
#pragma unroll
      for (int k = 0; k < 2 * GRP; ++k) v[k] = (k < gsz) ? __ldcg(src + (int64_t)k * nv + e) : make_double2(0.0, 0.0);
      double2 s = v[0];
#pragma unroll
      for (int k = 1; k < 2 * GRP; ++k) { s.x += v[k].x; s.y += v[k].y; }
      dst[e] = s;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    cnt[1 + grp] = 0;
    s_flag = (atomicAdd(&cnt[0], 1) == ngrp - 1);
  }
  __syncthreads();
  if (!s_flag) return false;
  __threadfence();
  {
    const double2* src = reinterpret_cast<const double2*>(gpart);
    for (int e = tid; e < nv; e += FT) {
      double2 s = make_double2(0.0, 0.0);
      for (int k0 = 0; k0 < ngrp; k0 += 16) {
        double2 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (k0 + k < ngrp) ? __ldcg(src + (int64_t)(k0 + k) * nv + e) : make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 16; ++k) { s.x += v[k].x; s.y += v[k].y; }
      }
      if (2 * e < len) acc[2 * e] = s.x;
      if (2 * e + 1 < len) acc[2 * e + 1] = s.y;
    }
  }
  if (tid == 0) cnt[0] = 0;
  __threadfence();
  __syncthreads();
  return true;
}

// ---- one-shot all-reduce over peer memory, run by the CTA that completed the rank's total -------------------------
// Every rank stores its packed accumulator into slot [parity][rank] of EVERY rank's exchange area (coalesced 16-byte
// stores over NVLink), publishes the sequence number in the peers' flag words, waits until all ranks' flags of its own
// area carry that number and sums the slots in rank order: the same bits on every rank, no collective launch between
// the pass kernels, ~1 NVLink round trip.  Two parities: a rank can be one exchange ahead of a peer that is still
// reading the previous one, never two (it needs that peer's flag to finish its own).  The wait is bounded (~2 s):
// a rank that never shows up raises the error flag instead of hanging the GPU.
__device__ __forceinline__ void p2p_allreduce(const FusedArgs& a, double* acc, int len, int tid) {
  const int W = a.world, par = (int)(a.seq & 1ull);
  const int nv = (len + 1) >> 1;                                   // 16-byte chunks (acc buffers have even capacity)
  const size_t slot_off = ((size_t)par * W + a.rank) * GPS_P2P_SLOT;
  const double2* src = reinterpret_cast<const double2*>(acc);
  for (int r = 0; r < W; ++r) {
    double2* dst = reinterpret_cast<double2*>(a.peers[r] + slot_off);
    for (int e = tid; e < nv; e += FT) dst[e] = __ldcg(src + e);
  }
  __threadfence_system();
  __syncthreads();
  unsigned long long* myflags = reinterpret_cast<unsigned long long*>(a.peers[a.rank] + (size_t)2 * W * GPS_P2P_SLOT) + (size_t)par * W;
  if (tid < W) {
    unsigned long long* pf = reinterpret_cast<unsigned long long*>(a.peers[tid] + (size_t)2 * W * GPS_P2P_SLOT) + (size_t)par * W + a.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pf), "l"(a.seq) : "memory");
    const long long t0 = clock64();
    unsigned long long seen = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(myflags + tid) : "memory");
      if (clock64() - t0 > 4000000000ll) {                        // ~2 s: give up, flag the evaluation as failed
        atomicCAS(a.info, 0, 3000000 + tid);
        break;
      }
    } while (seen < a.seq);
  }
  __syncthreads();
  const double* base = a.peers[a.rank] + (size_t)par * W * GPS_P2P_SLOT;
  for (int e = tid; e < nv; e += FT) {
    double2 s = make_double2(0.0, 0.0);
    for (int r = 0; r < W; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(base + (size_t)r * GPS_P2P_SLOT) + e);
      s.x += v.x;
      s.y += v.y;
    }
    if (2 * e < len) acc[2 * e] = s.x;
    if (2 * e + 1 < len) acc[2 * e + 1] = s.y;
  }
  __threadfence();
  __syncthreads();
}

// sum the warps' accumulator fragments into a dense [MP][ld] shared matrix (zeroed by the caller), warp by warp
template <class C>
__device__ __forceinline__ void tri_frags_to_smem(const double (&acc)[C::NTRI][2], double* S, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int w = 0; w < FW; ++w) {
    if (warp == w) {
      int idx = 0;
#pragma unroll
      for (int i = 0; i < C::MT; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          S[(8 * i + g) * C::MP + 8 * j + 2 * t] += acc[idx][0];
          S[(8 * i + g) * C::MP + 8 * j + 2 * t + 1] += acc[idx][1];
          ++idx;
        }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double sym_lower(const double* __restrict__ A, int MP, int r, int c) {
  return (r >= c) ? A[r * MP + c] : A[c * MP + r];
}

// =====================================================================================================================
// pass 1
// =====================================================================================================================
template <class C>
__global__ void __launch_bounds__(FT, 3) fused_p1_kernel(const __grid_constant__ FusedArgsP1 ap) {
  const FusedArgs& a = ap.a;
  constexpr int MP = C::MP, KS = C::KS, MT = C::MT, DT = C::DT, LDA = C::LDA, LDU = C::LDU;
  extern __shared__ __align__(16) double sh[];
  double* par = sh;                       // [24]   ea, sn2, invl[DT]
  double* Us = par + 24;                  // [KP][LDU]
  double* Aop = Us + C::KP * LDU;         // [MP][LDA]   L_A^-1
  double* B0 = Aop + MP * LDA;            // [MP][MP]
  double* B1 = B0 + MP * MP;              // [MP][MP]
  double* Li = B1 + MP * MP;              // [MP]
  double* scr = Li + MP;                  // [FW][2][TR]
  double* xtl = scr + FW * 2 * TR;        // [FW][2][TR][LDU]  double-buffered raw X tiles (cp.async)
  double* ytl = xtl + FW * 2 * TR * LDU;  // [FW][2][TR]       ... and their targets
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int D = a.D, M = a.M;
  const FL fl(MP, D);
  const bool st0 = blockIdx.x == 0 && tid == 0;
  if (st0) stamp(a, 0);
  // ---- preamble: kernel parameters, K_uu, L_A, L_A^-1 --------------------------------------------------------------
  if (st0) *a.info = 0;                   // only CTA 0 ever raises the flag (below): no memset in front of the launch
  const double* thU = ap.inline_thu ? ap.thu : a.thU;
  if (tid < 24) {
    double v = 0.0;
    if (tid == 0) v = exp(thU[0]);
    else if (tid == 1) v = exp(thU[D + 1]);
    else if (tid - 2 < D) v = exp(-thU[1 + (tid - 2)]);
    par[tid] = v;
  }
  __syncthreads();
  const double ea = par[0], sn2 = par[1];
  for (int e = tid; e < C::KP * LDU; e += FT) {
    const int m = e / LDU, d = e - m * LDU;
    Us[e] = (m < M && d < D) ? thU[D + 2 + m * D + d] * par[2 + d] : 0.0;
  }
  __syncthreads();
  if (st0) stamp(a, 5);
  for (int e = tid; e < MP * MP; e += FT) {
    const int i = e / MP, j = e - i * MP;
    double v = 0.0;
    if (i < M && j < M) {
      double r2 = 0.0;
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        const double df = Us[i * LDU + d] - Us[j * LDU + d];
        r2 = fma(df, df, r2);
      }
      v = ea * exp_neg(-0.5 * r2);
    }
    if (blockIdx.x == 0) a.fs[fl.kuu + e] = v;
    B0[e] = v + ((i == j) ? ((i < M) ? a.jitter : 1.0) : 0.0);
  }
  __syncthreads();
  if (st0) stamp(a, 6);
  if (warp == 0) {
    const int bad = warp_chol<MP>(B0, lane);
    if (bad && lane == 0 && blockIdx.x == 0) atomicCAS(a.info, 0, bad);
    if (lane < MP) Li[lane] = 1.0 / B0[lane * MP + lane];
    __syncwarp();
    if (st0) stamp(a, 7);
    warp_tri_inverse<MP>(B0, Li, B1, lane);
  }
  __syncthreads();
  if (st0) stamp(a, 8);
  for (int e = tid; e < MP * LDA; e += FT) {
    const int m = e / LDA, k = e - m * LDA;
    Aop[e] = (k < MP) ? B1[m * MP + k] : 0.0;
  }
  if (blockIdx.x == 0) {
    if (tid < 24) a.fs[fl.par + tid] = par[tid];
    for (int e = tid; e < MP * 16; e += FT) {
      const int m = e / 16, d = e - m * 16;
      a.fs[fl.us + e] = (m < C::KP && d < DT) ? Us[m * LDU + d] : 0.0;
    }
    for (int e = tid; e < MP * MP; e += FT) {
      a.fs[fl.la + e] = B0[e];
      a.fs[fl.lainv + e] = B1[e];
    }
  }
  __syncthreads();
  if (st0) stamp(a, 1);
  // ---- row pass ---------------------------------------------------------------------------------------------------------
  const int iE = MT - 1, gE = M & 7;
  double* sc = scr + warp * 2 * TR;
  double cacc[C::NTRI][2];
#pragma unroll
  for (int k = 0; k < C::NTRI; ++k) cacc[k][0] = cacc[k][1] = 0.0;
  double invl[DT];
#pragma unroll
  for (int d = 0; d < DT; ++d) invl[d] = par[2 + d];
  const int64_t N = a.N;
  double* lamg = a.rowv;
  // X tiles arrive through an 8-byte cp.async double buffer per warp: the loads of tile n + 1 are in flight while
  // tile n is processed (the row loop is otherwise exposed to the full HBM latency once per tile)
  double* xt = xtl + warp * 2 * TR * LDU;
  for (int e = lane; e < 2 * TR * LDU; e += 32) xt[e] = 0.0;            // pad columns d >= D stay zero
  __syncwarp();
  const int64_t rstep = (int64_t)gridDim.x * FW * TR;
  auto fetch_tile = [&](int64_t r0, int buf) {
    if (r0 < N) {
      const double* src = a.X + r0 * D;
      const int64_t left = (N - r0) * D;
      for (int c = lane; c < TR * D; c += 32) {
        const int row = c / D, d = c - row * D;
        cp_async8(xt + (buf * TR + row) * LDU + d, src + (c < left ? c : 0), c < left);
      }
      fetch_span(ytl + (warp * 2 + buf) * TR, a.y + r0, TR, true, lane, 0);   // y is zero-padded to the 128 tile
    }
    cp_async_commit_group();
  };
  int buf = 0;
  fetch_tile(((int64_t)blockIdx.x * FW + warp) * TR, 0);
  for (int64_t r0 = ((int64_t)blockIdx.x * FW + warp) * TR; r0 < N; r0 += rstep) {
    fetch_tile(r0 + rstep, buf ^ 1);
    cp_async_wait_group<1>();
    __syncwarp();
    const double* xb = xt + buf * TR * LDU;
    buf ^= 1;
    double vT[4][MT][2];
    double q8[8];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const int64_t row = r0 + 8 * gi + g;
      const bool live = row < N;
      double xs[DT];
#pragma unroll
      for (int d = 0; d < DT; ++d) xs[d] = xb[(8 * gi + g) * LDU + d] * invl[d];
      double kb[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int m = 4 * s + t;
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < DT; ++d) {
          const double df = Us[m * LDU + d] - xs[d];
          r2 = fma(df, df, r2);
        }
        const bool on = live && m < M;
        kb[s] = on ? ea * exp_neg(-0.5 * r2) : 0.0;
        if (on) a.Kst[(int64_t)m * a.ldk + row] = kb[s];
      }
      tile_product<C, true>(Aop, kb, g, t, vT[gi]);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        s0 = fma(vT[gi][i][0], vT[gi][i][0], s0);
        s1 = fma(vT[gi][i][1], vT[gi][i][1], s1);
      }
      q8[2 * gi] = s0;
      q8[2 * gi + 1] = s1;
    }
    const double q = rs8(q8, lane);
    const int rho = 8 * (g >> 1) + 2 * t + (g & 1);
    const int64_t orow = r0 + rho;
    const bool olive = orow < N;
    const double lam = ea + sn2 - q;
    const double rsq = olive ? 1.0 / sqrt(lam) : 0.0;
    const double yv = olive ? ytl[(warp * 2 + (buf ^ 1)) * TR + rho] : 0.0;    // (buf was flipped after the wait)
    if (olive) lamg[orow] = lam;
    sc[rho] = rsq;
    sc[TR + rho] = yv * rsq;
    __syncwarp();
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const double2 r2v = *reinterpret_cast<const double2*>(sc + 8 * gi + 2 * t);
      const double2 y2v = *reinterpret_cast<const double2*>(sc + TR + 8 * gi + 2 * t);
      double fa[MT][2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        fa[i][0] = vT[gi][i][0] * r2v.x;
        fa[i][1] = vT[gi][i][1] * r2v.y;
      }
      if (g == gE) {
        fa[iE][0] = y2v.x;
        fa[iE][1] = y2v.y;
      }
      tri_accumulate<C>(cacc, fa, fa);
    }
    __syncwarp();
  }
  // ---- reduction ------------------------------------------------------------------------------------------------------
  __syncthreads();
  if (st0) stamp(a, 2);
  for (int e = tid; e < MP * MP; e += FT) B0[e] = 0.0;
  __syncthreads();
  tri_frags_to_smem<C>(cacc, B0, warp, lane);
  if (st0) stamp(a, 3);
  if (ticket_reduce(B0, len1_of(MP), a.part, a.gpart, a.cnt, a.acc1, tid)) {
    if (a.peers) p2p_allreduce(a, a.acc1, len1_of(MP), tid);
    if (tid == 0) stamp(a, 4);
  }
}

// =====================================================================================================================
// pass 2
// =====================================================================================================================
template <class C>
__global__ void __launch_bounds__(FT, 3) fused_p2_kernel(FusedArgs a) {
  constexpr int MP = C::MP, KS = C::KS, MT = C::MT, LDA = C::LDA;
  extern __shared__ __align__(16) double sh[];
  double* Aop = sh;                       // [MP][LDA]   T2 with c2 in row M
  double* Li = Aop + MP * LDA;            // [MP]
  double* vy = Li + MP;                   // [MP]
  double* beta = vy + MP;                 // [MP]
  double* c2 = beta + MP;                 // [MP]
  double* scr = c2 + MP;                  // [FW][3][TR]
  double* red = scr + FW * 3 * TR;        // [32]
  double* rvl = red + 32;                 // [FW][2][2][TR]  lambda and y of the double-buffered tiles
  // one region, two lives: the preamble's four M x M matrices, then the double-buffered K_uf tiles of the row loop
  // (and the CTA partial after it) — keeps the kernel at three CTAs per SM
  double* B0 = rvl + FW * 4 * TR;         // C -> L_C
  double* B1 = B0 + MP * MP;              // L_C^-1
  double* B2 = B1 + MP * MP;              // L_A^-1
  double* B3 = B2 + MP * MP;              // T2
  double* ktl = B0;                       // [FW][2][KP][TR + 4]  K_uf tiles (cp.async)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int D = a.D, M = a.M;
  const FL fl(MP, D);
  const bool st0 = blockIdx.x == 0 && tid == 0;
  if (st0) stamp(a, 16);
  // ---- preamble: C = I + acc1, L_C, L_C^-1, beta, T2 = L_C^-1 L_A^-1, c2 = T2' beta ------------------------------------
  for (int e = tid; e < MP * MP; e += FT) {
    const int r = e / MP, c = e - r * MP;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < M && c < M) v += sym_lower(a.acc1, MP, r, c);
    B0[e] = v;
    B2[e] = a.fs[fl.lainv + e];
  }
  if (tid < MP) vy[tid] = (tid < M) ? a.acc1[M * MP + tid] : 0.0;
  __syncthreads();
  if (st0) stamp(a, 21);
  if (warp == 0) {
    const int bad = warp_chol<MP>(B0, lane);
    if (bad && lane == 0 && blockIdx.x == 0) atomicCAS(a.info, 0, 1000000 + bad);
    if (lane < MP) Li[lane] = 1.0 / B0[lane * MP + lane];
    __syncwarp();
    if (st0) stamp(a, 22);
    warp_tri_inverse<MP>(B0, Li, B1, lane);
  }
  __syncthreads();
  if (st0) stamp(a, 23);
  smv<MP, false>(B1, vy, beta, tid);                                  // beta = L_C^-1 v_y
  smm<MT, false, false>(B1, B2, B3, MP, MP, 1.0, 0.0, warp, lane);    // T2 = L_C^-1 L_A^-1
  __syncthreads();
  smv<MP, true>(B3, beta, c2, tid);                                   // c2 = T2' beta
  __syncthreads();
  for (int e = tid; e < MP * LDA; e += FT) {
    const int m = e / LDA, k = e - m * LDA;
    double v = (k < MP) ? B3[m * MP + k] : 0.0;
    if (m == M) v = (k < M) ? c2[k] : 0.0;
    Aop[e] = v;
  }
  if (blockIdx.x == 0) {
    for (int e = tid; e < MP * MP; e += FT) {
      a.fs[fl.lc + e] = B0[e];
      a.fs[fl.lcinv + e] = B1[e];
      a.fs[fl.t2 + e] = B3[e];
    }
    if (tid < MP) {
      a.fs[fl.beta + tid] = beta[tid];
      a.fs[fl.c2 + tid] = (tid < M) ? c2[tid] : 0.0;
    }
  }
  __syncthreads();
  if (st0) stamp(a, 17);
  // ---- row pass ---------------------------------------------------------------------------------------------------------
  const int iE = MT - 1, gE = M & 7;
  double* sc = scr + warp * 3 * TR;
  double racc[C::NTRI][2];
#pragma unroll
  for (int k = 0; k < C::NTRI; ++k) racc[k][0] = racc[k][1] = 0.0;
  double obj = 0.0;
  const int64_t N = a.N, ld = a.ldk;
  const double invN = a.invN;
  const int score = a.score;
  double* lamg = a.rowv;
  double* lb0g = a.rowv + ld;
  double* rbg = a.rowv + 2 * ld;
  double* tbg = a.rowv + 3 * ld;
  double* alg = a.rowv + 4 * ld;
  double* dg = a.rowv + 5 * ld;
  constexpr int LDK2 = TR + 4;
  double* ktw = ktl + warp * 2 * C::KP * LDK2;
  const int64_t rstep = (int64_t)gridDim.x * FW * TR;
  int buf = 0;
  double* rvw = rvl + warp * 4 * TR;
  auto fetch2 = [&](int64_t r0, int b) {
    fetch_k_tile<TR>(ktw + b * C::KP * LDK2, a.Kst, ld, r0, N, M, lane);
    fetch_span(rvw + b * 2 * TR, lamg + r0, TR, r0 < N, lane, 0);
    fetch_span(rvw + b * 2 * TR + TR, a.y + r0, TR, r0 < N, lane, 16);
    cp_async_commit_group();
  };
  fetch2(((int64_t)blockIdx.x * FW + warp) * TR, 0);
  for (int64_t r0 = ((int64_t)blockIdx.x * FW + warp) * TR; r0 < N; r0 += rstep) {
    fetch2(r0 + rstep, buf ^ 1);
    cp_async_wait_group<1>();
    __syncwarp();
    const double* kt = ktw + buf * C::KP * LDK2;
    const double* rv = rvw + buf * 2 * TR;
    buf ^= 1;
    double wT[4][MT][2];
    double r8[8];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const int64_t row = r0 + 8 * gi + g;
      const bool live = row < N;
      double kb[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int m = 4 * s + t;
        kb[s] = (live && m < M) ? kt[m * LDK2 + 8 * gi + g] : 0.0;
      }
      tile_product<C, true>(Aop, kb, g, t, wT[gi]);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const bool slot = (i == iE) && (g == gE);
        s0 = slot ? s0 : fma(wT[gi][i][0], wT[gi][i][0], s0);
        s1 = slot ? s1 : fma(wT[gi][i][1], wT[gi][i][1], s1);
      }
      r8[2 * gi] = s0;
      r8[2 * gi + 1] = s1;
      if (g == gE) {
        sc[2 * TR + 8 * gi + 2 * t] = wT[gi][iE][0];       // W_i . beta of rows 2t, 2t+1
        sc[2 * TR + 8 * gi + 2 * t + 1] = wT[gi][iE][1];
      }
    }
    const double r = rs8(r8, lane);
    __syncwarp();
    const int rho = 8 * (g >> 1) + 2 * t + (g & 1);
    const int64_t orow = r0 + rho;
    const bool olive = orow < N;
    double rbar = 0.0, tbar = 0.0;
    if (olive) {
      const double wb = sc[2 * TR + rho];
      const double lam = rv[rho], yi = rv[TR + rho];
      const double il = 1.0 / lam;
      const double d = il - r * il * il;
      const double alpha = (yi - wb) * il;
      double abar, dbar, lb0 = 0.0;
      if (score == GPS_CRPS) {
        const double s2 = 1.0 / d, s = sqrt(s2), z = alpha * s;
        const double tpm1 = erf(z * F_INV_SQRT2);
        const double gg = z * tpm1 + 2.0 * F_INV_SQRT_2PI * exp_neg(-0.5 * z * z) - F_INV_SQRT_PI;
        obj += s * gg * invN;
        abar = tpm1 * s2 * invN;
        dbar = -(0.5 * s2 * s * gg + 0.5 * tpm1 * alpha * s2 * s2) * invN;
      } else if (score == GPS_LOGS) {
        const double s2 = 1.0 / d;
        obj += (0.5 * alpha * alpha * s2 - 0.5 * log(d) + F_HALF_LOG_2PI) * invN;
        abar = alpha * s2 * invN;
        dbar = -(0.5 * alpha * alpha * s2 * s2 + 0.5 * s2) * invN;
      } else {  // NLML: 0.5 log lambda + 0.5 y alpha per row (K20:337-340 through Woodbury)
        obj += 0.5 * log(lam) + 0.5 * yi * alpha;
        abar = 0.5 * yi;
        dbar = 0.0;
        lb0 = 0.5 * il;
      }
      lb0 += dbar * (-il * il + 2.0 * r * il * il * il) - abar * alpha * il;
      rbar = -dbar * il * il;
      tbar = -abar * il;
      lb0g[orow] = lb0;
      rbg[orow] = rbar;
      tbg[orow] = tbar;
      alg[orow] = alpha;
      dg[orow] = d;
    }
    sc[rho] = rbar;
    sc[TR + rho] = tbar;
    __syncwarp();
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const double2 rb = *reinterpret_cast<const double2*>(sc + 8 * gi + 2 * t);
      const double2 tb = *reinterpret_cast<const double2*>(sc + TR + 8 * gi + 2 * t);
      double fa[MT][2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        fa[i][0] = wT[gi][i][0] * rb.x;
        fa[i][1] = wT[gi][i][1] * rb.y;
      }
      if (g == gE) {
        fa[iE][0] = tb.x;
        fa[iE][1] = tb.y;
      }
      tri_accumulate<C>(racc, fa, wT[gi]);
    }
    __syncwarp();
  }
  // ---- reduction ------------------------------------------------------------------------------------------------------
  __syncthreads();
  if (st0) stamp(a, 18);
  for (int e = tid; e < MP * MP + 1; e += FT) B0[e] = 0.0;     // B0 | B1 are contiguous: MP*MP + 1 doubles fit
  __syncthreads();
  tri_frags_to_smem<C>(racc, B0, warp, lane);
  const double o = block_sum(obj, red);
  if (tid == 0) B0[MP * MP] = o;
  __syncthreads();
  if (st0) stamp(a, 19);
  if (ticket_reduce(B0, len2_of(MP), a.part, a.gpart, a.cnt, a.acc2, tid)) {
    if (a.peers) p2p_allreduce(a, a.acc2, len2_of(MP), tid);
    if (tid == 0) stamp(a, 20);
  }
}

// =====================================================================================================================
// finishing step (one CTA): S, L_A_bar -> A_bar, all gradients.  `sh` needs 10 MP^2 + 8 MP + 64 doubles.
// =====================================================================================================================
template <class C>
__device__ void fused_finish(const FusedArgs& a, double* sh, int tid, int warp, int lane) {
  constexpr int MP = C::MP, MT = C::MT, DT = C::DT;
  const int D = a.D, M = a.M;
  const FL fl(MP, D);
  double* LAi = sh;                 // L_A^-1
  double* LC = LAi + MP * MP;
  double* LCi = LC + MP * MP;       // L_C^-1
  double* Cb = LCi + MP * MP;       // C_bar
  double* Rm = Cb + MP * MP;        // R
  double* Cm = Rm + MP * MP;        // C - I
  double* Zm = Cm + MP * MP;        // Z
  double* T1 = Zm + MP * MP;
  double* S = T1 + MP * MP;
  double* LA = S + MP * MP;
  double* beta = LA + MP * MP;      // [MP] each
  double* bbar = beta + MP;
  double* vyb = bbar + MP;
  double* vy = vyb + MP;
  double* b1 = vy + MP;
  double* lcb = b1 + MP;
  double* S0 = lcb + MP;
  double* red = S0 + MP;            // [32]
  const double* acc3 = a.acc3;
  const double* Pm = acc3 + MP * MP;          // [MP][DT]
  const double* S0g = Pm + MP * DT;           // [MP]
  const double* xq = S0g + MP;                // [DT]
  for (int e = tid; e < MP * MP; e += FT) {
    const int r = e / MP, c = e - r * MP;
    const bool in = r < M && c < M;
    LAi[e] = __ldcg(a.fs + fl.lainv + e);
    LA[e] = __ldcg(a.fs + fl.la + e);
    LC[e] = __ldcg(a.fs + fl.lc + e);
    LCi[e] = __ldcg(a.fs + fl.lcinv + e);
    Cb[e] = __ldcg(a.fs + fl.cbar + e);
    Rm[e] = in ? ((r >= c) ? __ldcg(a.acc2 + r * MP + c) : __ldcg(a.acc2 + c * MP + r)) : 0.0;
    Cm[e] = in ? ((r >= c) ? __ldcg(a.acc1 + r * MP + c) : __ldcg(a.acc1 + c * MP + r)) : 0.0;
    Zm[e] = in ? ((r >= c) ? __ldcg(acc3 + r * MP + c) : __ldcg(acc3 + c * MP + r)) : 0.0;
  }
  if (tid < MP) {
    const bool in = tid < M;
    beta[tid] = in ? __ldcg(a.fs + fl.beta + tid) : 0.0;
    bbar[tid] = in ? __ldcg(a.acc2 + M * MP + tid) : 0.0;
    vyb[tid] = in ? __ldcg(a.fs + fl.vyb + tid) : 0.0;
    vy[tid] = in ? __ldcg(a.acc1 + M * MP + tid) : 0.0;
    S0[tid] = in ? __ldcg(S0g + tid) : 0.0;
  }
  __syncthreads();
  if (tid == 0) stamp(a, 40);
  smv<MP, true>(LCi, beta, b1, tid);      // b1 = L_C^-T beta
  smv<MP, false>(LC, bbar, lcb, tid);     // L_C beta_bar
  smm<MT, false, true>(Rm, LC, T1, MP, MP, 1.0, 0.0, warp, lane);       // T1 = R L_C'
  __syncthreads();
  smm<MT, true, false>(LCi, T1, S, MP, MP, 2.0, 0.0, warp, lane);       // S = 2 L_C^-T R L_C'
  __syncthreads();
  smm<MT, false, false>(Cb, Cm, S, MP, MP, 2.0, 1.0, warp, lane);       // S += 2 C_bar (C - I)
  smm<MT, false, true>(Zm, LAi, T1, MP, MP, 1.0, 0.0, warp, lane);      // T1 = Z L_A^-T
  __syncthreads();
  smm<MT, false, false>(LAi, T1, S, MP, MP, -2.0, 1.0, warp, lane);     // S -= 2 L_A^-1 Z L_A^-T
  __syncthreads();
  for (int e = tid; e < MP * MP; e += FT) {
    const int r = e / MP, c = e - r * MP;
    S[e] += b1[r] * lcb[c] + vyb[r] * vy[c];
  }
  __syncthreads();
  if (tid == 0) stamp(a, 41);
  smm<MT, true, false>(LAi, S, T1, MP, MP, 1.0, 0.0, warp, lane);       // L_A^-T S
  __syncthreads();
  for (int e = tid; e < MP * MP; e += FT) {
    const int r = e / MP, c = e - r * MP;
    Zm[e] = (r >= c && r < M) ? -T1[e] : 0.0;                            // L_A_bar (Zm is free now)
  }
  __syncthreads();
  chol_adjoint_smm<MT>(LA, LAi, Zm, T1, S, tid, warp, lane);            // S = A_bar
  for (int e = tid; e < MP * MP; e += FT) {
    const int r = e / MP, c = e - r * MP;
    T1[e] = (r < M && c < M) ? S[e] * __ldcg(a.fs + fl.kuu + e) : 0.0;  // G2 = A_bar o K_uu
  }
  __syncthreads();
  if (tid == 0) stamp(a, 42);
  if (tid == 0) {
    double ld = 0.0;
    if (a.score == GPS_NLML)
      for (int m = 0; m < M; ++m) ld += log(LC[m * MP + m]);                 // sum log diag(L_C)
    red[0] = ld;
  }
  __syncthreads();
  const double logdet_c = red[0];
  // the gradient loops below touch us / P / par many times: stage them over the (now dead) matrices LAi .. Zm
  // (7 MP^2 >= 16 MP + MP DT + 24 + DT + 1 doubles for every supported MP)
  double* us = LAi;                           // [MP][16]
  double* Ps = us + MP * 16;                  // [MP][DT]
  double* par = Ps + MP * DT;                 // [24]
  double* xqs = par + 24;                     // [DT + 1]
  static_assert(7 * MP * MP >= MP * 16 + MP * DT + 24 + DT + 1, "finish staging must fit in the dead matrices");
  for (int e = tid; e < MP * 16; e += FT) us[e] = __ldcg(a.fs + fl.us + e);
  for (int e = tid; e < MP * DT; e += FT) Ps[e] = __ldcg(Pm + e);
  if (tid < 24) par[tid] = __ldcg(a.fs + fl.par + tid);
  if (tid <= DT) xqs[tid] = __ldcg(xq + tid);
  __syncthreads();
  const double ea = par[0], sn2 = par[1];
  const double sum_lb = xqs[DT];
  double* out = a.out_dev;
  // one sweep over G2 = A_bar o K_uu for the amplitude sum and all D length-scale sums (D + 1 accumulators per
  // thread, one block reduction), instead of D + 1 sweeps with two barriers each
  double acc[DT + 1];
#pragma unroll
  for (int d = 0; d <= DT; ++d) acc[d] = 0.0;
  for (int e = tid; e < MP * MP; e += FT) {
    const int i = e / MP, j = e - i * MP;
    const double g2 = T1[e];
    acc[DT] += g2;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      const double df = us[i * 16 + d] - us[j * 16 + d];
      acc[d] = fma(g2, df * df, acc[d]);
    }
  }
  for (int m = tid; m < M; m += FT) {
    acc[DT] += S0[m];
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      const double u = us[m * 16 + d];
      acc[d] += u * u * S0[m] - 2.0 * u * Ps[m * DT + d];
    }
  }
  double* wred = S;                            // [FW][DT + 1]  (A_bar is folded into G2: S is free)
#pragma unroll
  for (int d = 0; d <= DT; ++d) {
    const double v = warp_sum(acc[d]);
    if (lane == 0) wred[warp * (DT + 1) + d] = v;
  }
  __syncthreads();
  if (tid <= DT) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < FW; ++w) v += wred[w * (DT + 1) + tid];
    if (tid == DT) {
      double obj = __ldcg(a.acc2 + MP * MP);
      if (a.score == GPS_NLML) obj += 0.5 * a.world_n * 1.83787706640934548356 + logdet_c;   // N/2 log 2 pi + sum log diag(L_C)
      out[0] = obj;
      out[1] = ea * sum_lb + v;
      out[1 + D + 1] = sn2 * sum_lb;
    } else if (tid < D) {
      out[2 + tid] = v + xqs[tid];
    }
  }
  if (tid == 0) stamp(a, 43);
  for (int e = tid; e < M * D; e += FT) {
    const int m = e / D, d = e - m * D;
    const double u = us[m * 16 + d];
    double tt = 0.0;
    for (int j = 0; j < M; ++j) tt = fma(T1[m * MP + j], u - us[j * 16 + d], tt);
    out[1 + D + 2 + e] = -par[2 + d] * (u * S0[m] - Ps[m * DT + d]) - 2.0 * par[2 + d] * tt;
  }
  if (tid == 0) out[1 + D + 2 + M * D] = (double)atomicAdd(a.info, 0);
  __syncthreads();
  // results to mapped host memory in one coalesced sweep (single stores from one thread crawl over PCIe)
  if (a.out_host) {
    const int nout = 2 + D + 2 + M * D;
    for (int e = tid; e < nout; e += FT) a.out_host[e] = out[e];
  }
}

__host__ __device__ inline int finish_smem_doubles(int MP) { return 10 * MP * MP + 8 * MP + 64; }

// =====================================================================================================================
// pass 3
// =====================================================================================================================
template <class C>
__global__ void __launch_bounds__(FT, 2) fused_p3_kernel(FusedArgs a) {
  constexpr int MP = C::MP, KS = C::KS, MT = C::MT, DT = C::DT, LDA = C::LDA, NXT = C::NXT;
  constexpr int R3 = 16;                  // rows per warp tile in this pass
  extern __shared__ __align__(16) double sh[];
  double* E = sh;                         // [3][MP][LDA]   E1, E2, E3
  double* vec = E + 3 * MP * LDA;         // a1, a2, c1, invl : [4][MP]  (invl: DT <= MP entries)
  double* scr = vec + 4 * MP;             // [FW][R3]
  double* red = scr + FW * R3;            // [32]
  double* ktl = red + 32;                 // [FW][2][KP][R3 + 4]  double-buffered K_uf tiles (cp.async)
  double* rvl = ktl + FW * 2 * C::KP * (R3 + 4);   // [FW][2][5][R3]   ... lambda, lb0, rbar, tbar, y of the tile
  double* xrl = rvl + FW * 2 * 5 * R3;             // [FW][2][R3][DT]  ... and its raw X rows
  double* W0 = xrl + FW * 2 * R3 * DT;             // preamble / reduction / finish workspace
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int D = a.D, M = a.M;
  const FL fl(MP, D);
  const bool st0 = blockIdx.x == 0 && tid == 0;
  if (st0) stamp(a, 32);
  // ---- preamble ---------------------------------------------------------------------------------------------------------
  {
    double* LC = W0;
    double* LCi = LC + MP * MP;
    double* LAi = LCi + MP * MP;
    double* T2 = LAi + MP * MP;
    double* SW = T2 + MP * MP;
    double* X1 = SW + MP * MP;
    double* T1 = X1 + MP * MP;
    double* Cb = T1 + MP * MP;
    double* beta = Cb + MP * MP;
    double* bbar = beta + MP;
    double* vyb = bbar + MP;
    for (int e = tid; e < MP * MP; e += FT) {
      LC[e] = a.fs[fl.lc + e];
      LCi[e] = a.fs[fl.lcinv + e];
      LAi[e] = a.fs[fl.lainv + e];
      T2[e] = a.fs[fl.t2 + e];
    }
    if (tid < MP) {
      beta[tid] = (tid < M) ? a.fs[fl.beta + tid] : 0.0;
      bbar[tid] = (tid < M) ? a.acc2[M * MP + tid] : 0.0;
      vec[tid] = a.fs[fl.c2 + tid];                              // a1 = c2 = T2' beta
      vec[3 * MP + tid] = (tid < DT) ? a.fs[fl.par + 2 + tid] : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < MP * MP; e += FT) {
      const int r = e / MP, c = e - r * MP;
      SW[e] = (r < M && c < M) ? beta[r] * bbar[c] + 2.0 * sym_lower(a.acc2, MP, r, c) + bbar[r] * beta[c] : 0.0;
    }
    __syncthreads();
    if (st0) stamp(a, 44);
    smm<MT, true, false>(LCi, SW, X1, MP, MP, 1.0, 0.0, warp, lane);     // L_C^-T S_W
    smv<MP, true>(LCi, bbar, vyb, tid);                                  // vy_bar = L_C^-T beta_bar
    __syncthreads();
    for (int e = tid; e < MP * MP; e += FT) {
      const int r = e / MP, c = e - r * MP;
      double v = (r >= c && r < M) ? -X1[e] : 0.0;
      if (a.score == GPS_NLML && r == c && r < M) v += LCi[e];           // d/dL_C of sum log diag(L_C)
      X1[e] = v;                                                         // L_C_bar
    }
    __syncthreads();
    if (st0) stamp(a, 45);
    chol_adjoint_smm<MT>(LC, LCi, X1, T1, Cb, tid, warp, lane);          // C_bar
    if (st0) stamp(a, 46);
    smv<MP, true>(LAi, vyb, vec + MP, tid);                              // a2 = L_A^-T vy_bar
    smv<MP, true>(T2, bbar, vec + 2 * MP, tid);                          // c1 = T2' beta_bar
    smm<MT, true, false>(T2, T2, E, LDA, C::KP, 1.0, 0.0, warp, lane);                  // E1 = T2' T2
    smm<MT, true, false>(LAi, LAi, E + 2 * MP * LDA, LDA, C::KP, 1.0, 0.0, warp, lane); // E3 = A^-1
    smm<MT, false, false>(Cb, LAi, T1, MP, MP, 1.0, 0.0, warp, lane);                   // T1 = C_bar L_A^-1
    __syncthreads();
    smm<MT, true, false>(LAi, T1, E + MP * LDA, LDA, C::KP, 1.0, 0.0, warp, lane);      // E2 = L_A^-T C_bar L_A^-1
    if (blockIdx.x == 0) {
      for (int e = tid; e < MP * MP; e += FT) a.fs[fl.cbar + e] = Cb[e];
      if (tid < MP) {
        a.fs[fl.vyb + tid] = vyb[tid];
        a.fs[fl.bbar + tid] = bbar[tid];
      }
    }
    __syncthreads();
  }
  if (st0) stamp(a, 33);
  // ---- row pass ---------------------------------------------------------------------------------------------------------
  double* sc = scr + warp * R3;
  double zacc[C::NTRI][2], pacc[MT][NXT][2], s0acc[MT], xqacc[NXT];
#pragma unroll
  for (int k = 0; k < C::NTRI; ++k) zacc[k][0] = zacc[k][1] = 0.0;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    s0acc[i] = 0.0;
#pragma unroll
    for (int j = 0; j < NXT; ++j) pacc[i][j][0] = pacc[i][j][1] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < NXT; ++j) xqacc[j] = 0.0;
  double sum_lb = 0.0;
  double a1v[MT], a2v[MT], c1v[MT], invx[NXT];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    a1v[i] = vec[8 * i + g];
    a2v[i] = vec[MP + 8 * i + g];
    c1v[i] = vec[2 * MP + 8 * i + g];
  }
#pragma unroll
  for (int j = 0; j < NXT; ++j) invx[j] = vec[3 * MP + 8 * j + g];
  const int64_t N = a.N, ld = a.ldk;
  const double* lamg = a.rowv;
  const double* lb0g = a.rowv + ld;
  const double* rbg = a.rowv + 2 * ld;
  const double* tbg = a.rowv + 3 * ld;
  const double* E1 = E;
  const double* E2 = E + MP * LDA;
  const double* E3 = E + 2 * MP * LDA;
  constexpr int LDK3 = R3 + 4;
  double* ktw = ktl + warp * 2 * C::KP * LDK3;
  const int64_t rstep = (int64_t)gridDim.x * FW * R3;
  int buf = 0;
  double* rvw = rvl + warp * 2 * 5 * R3;
  double* xrw = xrl + warp * 2 * R3 * DT;
  auto fetch3 = [&](int64_t r0, int b) {
    const bool on = r0 < N;
    fetch_k_tile<R3>(ktw + b * C::KP * LDK3, a.Kst, ld, r0, N, M, lane);
    double* rvb = rvw + b * 5 * R3;
    fetch_span(rvb, lamg + r0, R3, on, lane, 0);
    fetch_span(rvb + R3, lb0g + r0, R3, on, lane, 8);
    fetch_span(rvb + 2 * R3, rbg + r0, R3, on, lane, 16);
    fetch_span(rvb + 3 * R3, tbg + r0, R3, on, lane, 24);
    fetch_span(rvb + 4 * R3, a.y + r0, R3, on, lane, 4);
    fetch_span(xrw + b * R3 * DT, a.X + r0 * D, R3 * D, on, lane, 12);       // X is zero-padded to the 128 tile
    cp_async_commit_group();
  };
  fetch3(((int64_t)blockIdx.x * FW + warp) * R3, 0);
  for (int64_t r0 = ((int64_t)blockIdx.x * FW + warp) * R3; r0 < N; r0 += rstep) {
    fetch3(r0 + rstep, buf ^ 1);
    cp_async_wait_group<1>();
    __syncwarp();
    const double* kt = ktw + buf * C::KP * LDK3;
    const double* rv = rvw + buf * 5 * R3;
    const double* xr = xrw + buf * R3 * DT;
    buf ^= 1;
    double kT[2][MT][2], pq[2][MT][2], p3[2][MT][2];
    double s4[4], b4[4];
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int64_t row = r0 + 8 * gi + g;
      const bool live = row < N;
      const int64_t rT = r0 + 8 * gi + 2 * t;         // layout-T rows rT, rT + 1 (rT even, ld a multiple of 8)
      const bool l0 = rT < N, l1 = rT + 1 < N;
      double kb[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int m = 4 * s + t;
        kb[s] = (live && m < M) ? kt[m * LDK3 + 8 * gi + g] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int m = 8 * i + g;
        double2 kk = make_double2(0.0, 0.0);
        if (m < M && l0) kk = *reinterpret_cast<const double2*>(kt + m * LDK3 + 8 * gi + 2 * t);
        kT[gi][i][0] = kk.x;
        kT[gi][i][1] = l1 ? kk.y : 0.0;
      }
      double2 lam2 = make_double2(1.0, 1.0), y2 = make_double2(0.0, 0.0), rb2 = y2, tb2 = y2;
      if (l0) {
        const int o = 8 * gi + 2 * t;
        lam2 = *reinterpret_cast<const double2*>(rv + o);
        rb2 = *reinterpret_cast<const double2*>(rv + 2 * R3 + o);
        tb2 = *reinterpret_cast<const double2*>(rv + 3 * R3 + o);
        y2 = *reinterpret_cast<const double2*>(rv + 4 * R3 + o);
        if (!l1) y2.y = 0.0;
      }
      if (!l1) { lam2.y = 1.0; rb2.y = 0.0; tb2.y = 0.0; }
      const double il0 = 1.0 / lam2.x, il1 = 1.0 / lam2.y;
      double p1[MT][2], p2[MT][2];
      tile_product<C, false>(E1, kb, g, t, p1);
      tile_product<C, false>(E2, kb, g, t, p2);
      tile_product<C, false>(E3, kb, g, t, p3[gi]);
      double s10 = 0.0, s11 = 0.0, bw0 = 0.0, bw1 = 0.0;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        s10 = fma(kT[gi][i][0], p2[i][0], s10);
        s11 = fma(kT[gi][i][1], p2[i][1], s11);
        bw0 = fma(c1v[i], kT[gi][i][0], bw0);
        bw1 = fma(c1v[i], kT[gi][i][1], bw1);
        pq[gi][i][0] = tb2.x * a1v[i] + (y2.x * il0) * a2v[i] + 2.0 * rb2.x * p1[i][0] + 2.0 * il0 * p2[i][0];
        pq[gi][i][1] = tb2.y * a1v[i] + (y2.y * il1) * a2v[i] + 2.0 * rb2.y * p1[i][1] + 2.0 * il1 * p2[i][1];
      }
      s4[2 * gi] = s10;
      s4[2 * gi + 1] = s11;
      b4[2 * gi] = bw0;
      b4[2 * gi + 1] = bw1;
    }
    const double s1 = rs4(s4, lane);
    const double bw = rs4(b4, lane);
    const int rho = 8 * (g >> 2) + 2 * t + ((g >> 1) & 1);
    const int64_t orow = r0 + rho;
    double lb = 0.0;
    if (orow < N) {
      const double il = 1.0 / rv[rho];
      lb = rv[R3 + rho] - bw * rv[4 * R3 + rho] * il * il - s1 * il * il;
    }
    if ((g & 1) == 0) {
      sum_lb += lb;
      sc[rho] = lb;
    }
    __syncwarp();
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int64_t rT = r0 + 8 * gi + 2 * t;
      const double2 lb2 = *reinterpret_cast<const double2*>(sc + 8 * gi + 2 * t);
      double gT[MT][2], za[MT][2];
      double gs0 = 0.0, gs1 = 0.0;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        gT[i][0] = (pq[gi][i][0] - 2.0 * lb2.x * p3[gi][i][0]) * kT[gi][i][0];    // G = Kuf_bar o Kuf
        gT[i][1] = (pq[gi][i][1] - 2.0 * lb2.y * p3[gi][i][1]) * kT[gi][i][1];
        za[i][0] = lb2.x * kT[gi][i][0];
        za[i][1] = lb2.y * kT[gi][i][1];
        s0acc[i] += gT[i][0] + gT[i][1];
        gs0 += gT[i][0];
        gs1 += gT[i][1];
      }
      tri_accumulate<C>(zacc, za, kT[gi]);                                        // Z += lam_bar k k'
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        gs0 += __shfl_xor_sync(0xffffffffu, gs0, o);
        gs1 += __shfl_xor_sync(0xffffffffu, gs1, o);
      }
#pragma unroll
      for (int j = 0; j < NXT; ++j) {
        const int c = 8 * j + g;
        const int o = (8 * gi + 2 * t) * D + c;
        const double x0 = (c < D && rT < N) ? xr[o] * invx[j] : 0.0;
        const double x1 = (c < D && rT + 1 < N) ? xr[o + D] * invx[j] : 0.0;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          dmma(pacc[i][j][0], pacc[i][j][1], gT[i][0], x0);                       // P += G' xs
          dmma(pacc[i][j][0], pacc[i][j][1], gT[i][1], x1);
        }
        xqacc[j] = fma(x0 * x0, gs0, fma(x1 * x1, gs1, xqacc[j]));
      }
    }
    __syncwarp();
  }
  // ---- reduction ------------------------------------------------------------------------------------------------------
  __syncthreads();
  if (st0) stamp(a, 34);
  const int len3 = len3_of(MP, DT);
  double* Zs = W0;                         // [MP][MP] | P [MP][DT] | S0 [MP] | xq [DT] | sum_lb
  double* Ps = Zs + MP * MP;
  double* S0s = Ps + MP * DT;
  double* xqs = S0s + MP;
  for (int e = tid; e < len3; e += FT) Zs[e] = 0.0;
  __syncthreads();
  tri_frags_to_smem<C>(zacc, Zs, warp, lane);
#pragma unroll
  for (int i = 0; i < MT; ++i) {           // S0: sum over the 4 threads of a group (their rows)
    s0acc[i] += __shfl_xor_sync(0xffffffffu, s0acc[i], 1);
    s0acc[i] += __shfl_xor_sync(0xffffffffu, s0acc[i], 2);
  }
#pragma unroll
  for (int j = 0; j < NXT; ++j) {
    xqacc[j] += __shfl_xor_sync(0xffffffffu, xqacc[j], 1);
    xqacc[j] += __shfl_xor_sync(0xffffffffu, xqacc[j], 2);
  }
  for (int w = 0; w < FW; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < MT; ++i) {
#pragma unroll
        for (int j = 0; j < NXT; ++j) {
          Ps[(8 * i + g) * DT + 8 * j + 2 * t] += pacc[i][j][0];
          Ps[(8 * i + g) * DT + 8 * j + 2 * t + 1] += pacc[i][j][1];
        }
        if (t == 0) S0s[8 * i + g] += s0acc[i];
      }
      if (t == 0) {
#pragma unroll
        for (int j = 0; j < NXT; ++j) xqs[8 * j + g] += xqacc[j];
      }
    }
    __syncthreads();
  }
  const double sl = block_sum(sum_lb, red);
  if (tid == 0) xqs[DT] = sl;
  __syncthreads();
  if (st0) stamp(a, 35);
  const bool last = ticket_reduce(Zs, len3, a.part, a.gpart, a.cnt, a.acc3, tid);
  if (last && a.peers) p2p_allreduce(a, a.acc3, len3, tid);
  if (last && tid == 0) stamp(a, 36);
  if (last && a.finish) {
    fused_finish<C>(a, W0, tid, warp, lane);
    if (tid == 0) stamp(a, 37);
  }
}

// finishing step as its own kernel (row-sharded runs: after the all-reduce of acc3)
template <class C>
__global__ void __launch_bounds__(FT) fused_finish_kernel(FusedArgs a) {
  extern __shared__ __align__(16) double sh[];
  fused_finish<C>(a, sh, threadIdx.x, threadIdx.x >> 5, threadIdx.x & 31);
}

// ---- prediction (K20:270-277 -> spgp_cal_mean_and_cov K20:76-83, diagonal only) ---------------------------------------
// mean* = c2 . k*,  var* = sn2 + e^a - |L_A^-1 k*|^2 + |T2 k*|^2
template <class C>
__global__ void __launch_bounds__(FT, 2) fused_predict_kernel(const double* __restrict__ Xs, int64_t T, int D, int M,
                                                              const double* __restrict__ fs, double* __restrict__ mean,
                                                              double* __restrict__ var) {
  constexpr int MP = C::MP, KS = C::KS, MT = C::MT, DT = C::DT, LDA = C::LDA, LDU = C::LDU;
  extern __shared__ __align__(16) double sh[];
  double* par = sh;                       // [24]
  double* Us = par + 24;                  // [KP][LDU]
  double* A1 = Us + C::KP * LDU;          // L_A^-1
  double* A2 = A1 + MP * LDA;             // T2 with c2 in row M
  double* scr = A2 + MP * LDA;            // [FW][TR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const FL fl(MP, D);
  if (tid < 24) par[tid] = fs[fl.par + tid];
  for (int e = tid; e < C::KP * LDU; e += FT) {
    const int m = e / LDU, d = e - m * LDU;
    Us[e] = (d < 16) ? fs[fl.us + m * 16 + d] : 0.0;
  }
  for (int e = tid; e < MP * LDA; e += FT) {
    const int m = e / LDA, k = e - m * LDA;
    A1[e] = (k < MP) ? fs[fl.lainv + m * MP + k] : 0.0;
    double v = (k < MP) ? fs[fl.t2 + m * MP + k] : 0.0;
    if (m == M) v = (k < M) ? fs[fl.c2 + k] : 0.0;
    A2[e] = v;
  }
  __syncthreads();
  const double ea = par[0], sn2 = par[1];
  const int iE = MT - 1, gE = M & 7;
  double* sc = scr + warp * TR;
  double invl[DT];
#pragma unroll
  for (int d = 0; d < DT; ++d) invl[d] = par[2 + d];
  for (int64_t r0 = ((int64_t)blockIdx.x * FW + warp) * TR; r0 < T; r0 += (int64_t)gridDim.x * FW * TR) {
    double v8[8];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const int64_t row = r0 + 8 * gi + g;
      const bool live = row < T;
      double xs[DT];
#pragma unroll
      for (int d = 0; d < DT; ++d) xs[d] = (live && d < D) ? Xs[row * D + d] * invl[d] : 0.0;
      double kb[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int m = 4 * s + t;
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < DT; ++d) {
          const double df = Us[m * LDU + d] - xs[d];
          r2 = fma(df, df, r2);
        }
        kb[s] = (live && m < M) ? ea * exp_neg(-0.5 * r2) : 0.0;
      }
      double vT[MT][2], wT[MT][2];
      tile_product<C, true>(A1, kb, g, t, vT);
      tile_product<C, true>(A2, kb, g, t, wT);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const bool slot = (i == iE) && (g == gE);
        s0 += (slot ? 0.0 : wT[i][0] * wT[i][0]) - vT[i][0] * vT[i][0];
        s1 += (slot ? 0.0 : wT[i][1] * wT[i][1]) - vT[i][1] * vT[i][1];
      }
      v8[2 * gi] = s0;
      v8[2 * gi + 1] = s1;
      if (g == gE) {
        sc[8 * gi + 2 * t] = wT[iE][0];
        sc[8 * gi + 2 * t + 1] = wT[iE][1];
      }
    }
    const double dv = rs8(v8, lane);
    __syncwarp();
    const int rho = 8 * (g >> 1) + 2 * t + (g & 1);
    if (r0 + rho < T) {
      mean[r0 + rho] = sc[rho];
      var[r0 + rho] = sn2 + ea + dv;
    }
    __syncwarp();
  }
}

// theta -= lr * g_theta, U -= lr_u * g_U on the device (K20:243-251); trace[it] = objective before the step.
// A failed factorisation (info != 0) freezes the parameters and records the iteration.
__global__ void fused_update_kernel(double* thU, const double* __restrict__ out, int D, int M, double lr, double lr_u,
                                    double* __restrict__ trace, int it, int* __restrict__ info, int* __restrict__ fail_it) {
  const int P = D + 2, Q = M * D;
  const int bad = *info;
  if (bad != 0) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && *fail_it < 0) *fail_it = it;
    return;
  }
  for (int e = threadIdx.x; e < P + Q; e += blockDim.x) thU[e] -= (e < P ? lr : lr_u) * out[1 + e];
  if (threadIdx.x == 0 && trace) trace[it] = out[0];
}

__global__ void fused_loo_kernel(const double* __restrict__ y, const double* __restrict__ rowv, int64_t N, int64_t ld,
                                 double* __restrict__ mean, double* __restrict__ var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double alpha = rowv[4 * ld + i], d = rowv[5 * ld + i];
  mean[i] = y[i] - alpha / d;   // K20:231
  var[i] = 1.0 / d;             // K20:232
}

__global__ void empty_kernel() {}

// ---- shared-memory sizes (doubles) ------------------------------------------------------------------------------------
template <class C> constexpr int smem_p1() { return 24 + C::KP * C::LDU + C::MP * C::LDA + 2 * C::MP * C::MP + C::MP + FW * 2 * TR + FW * 2 * TR * C::LDU + FW * 2 * TR; }
template <class C> constexpr int smem_p2() {
  return C::MP * C::LDA + 4 * C::MP + FW * 3 * TR + 32 + FW * 4 * TR +
         (4 * C::MP * C::MP > FW * 2 * C::KP * (TR + 4) ? 4 * C::MP * C::MP : FW * 2 * C::KP * (TR + 4));
}
template <class C> constexpr int smem_p3() {
  return 3 * C::MP * C::LDA + 4 * C::MP + FW * 16 + 32 + FW * 2 * C::KP * (16 + 4) + FW * 2 * 5 * 16 + FW * 2 * 16 * C::DT +
         (finish_smem_doubles(C::MP) > len3_of(C::MP, C::DT) ? finish_smem_doubles(C::MP) : len3_of(C::MP, C::DT));
}
template <class C> constexpr int smem_pred() { return 24 + C::KP * C::LDU + 2 * C::MP * C::LDA + FW * TR; }

template <typename K>
int set_smem_attr(gps_ctx* ctx, K kern, size_t bytes) {
  GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return GPS_OK;
}

}  // namespace

// =====================================================================================================================
// host side
// =====================================================================================================================
struct gps_fitc_fused {
  int M = 0, D = 0, MT = 0, KS = 0, DT = 0, MP = 0;
  int64_t N = 0, ldk = 0;
  int grid = 0;              // CTAs that cover the rows once (one 32-row tile per warp)
  int occ[4] = {3, 3, 2, 1}; // resident CTAs per SM of passes 1..3 (registers): persistent grids are capped at occ x SMs
  DevBuf Kst, rowv, fs, part, gpart, acc1, acc2, acc3, dthU, dout, trace;
  int* cnt = nullptr;       // 3 x (1 + ngrp_max) tickets + fail_it
  int cnt_stride = 0;
  double* h_in = nullptr;   // mapped pinned: theta | U
  double* h_out = nullptr;  // mapped pinned: obj | g_theta | g_U | info
  long long* prof = nullptr; // debug phase stamps (gps_dbg_fused_phases)
  size_t h_cap = 0;
  bool configured[64] = {};
  bool ready = false;       // passes 1 + 2 have run at the current (theta, U): loo / predict are valid
};

namespace {

struct Combo { int MT, KS; };
inline Combo combo_of(int M) {
  const int MT = M / 8 + 1;
  int KS = (M + 3) / 4;
  if (MT == 1) KS = 2;
  else if (MT == 2) KS = 4;
  else if (MT == 3) KS = KS <= 5 ? 5 : 6;
  else KS = 8;
  return {MT, KS};
}

#define FUSED_DISPATCH(fu, CALL)                                                                  \
  do {                                                                                            \
    const int key__ = (fu)->MT * 100 + (fu)->KS * 2 + ((fu)->DT == 16 ? 1 : 0);                   \
    switch (key__) {                                                                              \
      case 104: { using CF = FCfg<1, 2, 8>;  CALL; } break;                                       \
      case 105: { using CF = FCfg<1, 2, 16>; CALL; } break;                                       \
      case 208: { using CF = FCfg<2, 4, 8>;  CALL; } break;                                       \
      case 209: { using CF = FCfg<2, 4, 16>; CALL; } break;                                       \
      case 310: { using CF = FCfg<3, 5, 8>;  CALL; } break;                                       \
      case 311: { using CF = FCfg<3, 5, 16>; CALL; } break;                                       \
      case 312: { using CF = FCfg<3, 6, 8>;  CALL; } break;                                       \
      case 313: { using CF = FCfg<3, 6, 16>; CALL; } break;                                       \
      case 416: { using CF = FCfg<4, 8, 8>;  CALL; } break;                                       \
      case 417: { using CF = FCfg<4, 8, 16>; CALL; } break;                                       \
      default: return gps_fail(ctx, GPS_EINVAL, "fitc fused: unsupported M / D combination");     \
    }                                                                                             \
  } while (0)

template <class CF>
int configure_kernels(gps_ctx* ctx) {
  GPS_CHECK(set_smem_attr(ctx, fused_p1_kernel<CF>, smem_p1<CF>() * 8));
  GPS_CHECK(set_smem_attr(ctx, fused_p2_kernel<CF>, smem_p2<CF>() * 8));
  GPS_CHECK(set_smem_attr(ctx, fused_p3_kernel<CF>, smem_p3<CF>() * 8));
  GPS_CHECK(set_smem_attr(ctx, fused_finish_kernel<CF>, finish_smem_doubles(CF::MP) * 8));
  GPS_CHECK(set_smem_attr(ctx, fused_predict_kernel<CF>, smem_pred<CF>() * 8));
  return GPS_OK;
}

template <class CF>
int launch_pass(gps_ctx* ctx, gps_fitc_fused* fu, int pass, const FusedArgs& a, const double* host_thu = nullptr) {
  FusedArgs b = a;
  b.cnt = fu->cnt + (pass - 1) * fu->cnt_stride;
  const int cap = ctx->sm_count * fu->occ[pass <= 3 ? pass - 1 : 3];
  const int grid = fu->grid < cap ? fu->grid : cap;
  if (pass == 1) {
    FusedArgsP1 bp;
    bp.a = b;
    const int nthu = fu->D + 2 + fu->M * fu->D;
    bp.inline_thu = (host_thu && nthu <= THU_INLINE) ? 1 : 0;
    if (bp.inline_thu) memcpy(bp.thu, host_thu, (size_t)nthu * sizeof(double));
    fused_p1_kernel<CF><<<grid, FT, smem_p1<CF>() * 8, ctx->stream>>>(bp);
  }
  else if (pass == 2) fused_p2_kernel<CF><<<grid, FT, smem_p2<CF>() * 8, ctx->stream>>>(b);
  else if (pass == 3) fused_p3_kernel<CF><<<grid, FT, smem_p3<CF>() * 8, ctx->stream>>>(b);
  else fused_finish_kernel<CF><<<1, FT, finish_smem_doubles(CF::MP) * 8, ctx->stream>>>(b);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

template <class CF>
int launch_predict(gps_ctx* ctx, gps_fitc_fused* fu, const double* Xs, int64_t T, double* mean, double* var) {
  int64_t blocks = (T + FW * TR - 1) / (FW * TR);
  if (blocks > 2 * ctx->sm_count) blocks = 2 * ctx->sm_count;
  fused_predict_kernel<CF><<<(unsigned)blocks, FT, smem_pred<CF>() * 8, ctx->stream>>>(Xs, T, fu->D, fu->M, fu->fs.p, mean, var);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

}  // namespace

bool gps_fitc_fused_supports(const gps_ctx* ctx, int M, int score) {
  return M >= 1 && M <= 31 && ctx->D >= 1 && ctx->D <= 16 && (score == GPS_CRPS || score == GPS_LOGS || score == GPS_NLML);
}

void gps_fitc_fused_free(gps_ctx* ctx) {
  gps_fitc_fused* fu = ctx->fu;
  if (!fu) return;
  for (DevBuf* b : {&fu->Kst, &fu->rowv, &fu->fs, &fu->part, &fu->gpart, &fu->acc1, &fu->acc2, &fu->acc3, &fu->dthU,
                    &fu->dout, &fu->trace})
    if (b->p) cudaFree(b->p);
  if (fu->cnt) cudaFree(fu->cnt);
  if (fu->h_in) cudaFreeHost(fu->h_in);
  if (fu->h_out) cudaFreeHost(fu->h_out);
  if (fu->prof) cudaFree(fu->prof);
  delete fu;
  ctx->fu = nullptr;
}

// (re)size the workspaces for the context's data set and M inducing points
int gps_fitc_fused_prepare(gps_ctx* ctx, int M) {
  if (!ctx->fu) ctx->fu = new gps_fitc_fused();
  gps_fitc_fused* fu = ctx->fu;
  const int D = ctx->D;
  const int64_t N = ctx->N;
  const Combo cb = combo_of(M);
  fu->M = M; fu->D = D; fu->MT = cb.MT; fu->KS = cb.KS; fu->DT = D <= 8 ? 8 : 16; fu->MP = 8 * cb.MT;
  fu->N = N;
  fu->ldk = (N + TR - 1) / TR * TR;        // whole warp tiles: the cp.async tile fetches never leave a row of K_uf
  const int64_t blocks = (N + FW * TR - 1) / (FW * TR);
  const int64_t cap = (int64_t)ctx->sm_count * 3;
  fu->grid = (int)blocks;
  const int MP = fu->MP;
  const size_t maxlen = ((size_t)len3_of(MP, fu->DT) + 1) & ~(size_t)1;   // partial rows use an even stride
  const FL fl(MP, D);
  GPS_CHECK(gps_ensure(ctx, fu->Kst, (size_t)M * fu->ldk));
  GPS_CHECK(gps_ensure(ctx, fu->rowv, (size_t)6 * fu->ldk));
  GPS_CHECK(gps_ensure(ctx, fu->fs, (size_t)fl.total));
  const size_t gmax = (size_t)cap;
  GPS_CHECK(gps_ensure(ctx, fu->part, gmax * maxlen));
  GPS_CHECK(gps_ensure(ctx, fu->gpart, (gmax / GRP + 1) * maxlen));
  GPS_CHECK(gps_ensure(ctx, fu->acc1, maxlen));
  GPS_CHECK(gps_ensure(ctx, fu->acc2, maxlen));
  GPS_CHECK(gps_ensure(ctx, fu->acc3, maxlen));
  const size_t nio = (size_t)(2 + D + 2 + M * D);
  GPS_CHECK(gps_ensure(ctx, fu->dthU, nio));
  GPS_CHECK(gps_ensure(ctx, fu->dout, nio));
  if (!fu->cnt) {
    fu->cnt_stride = 1 + (int)(gmax / GRP + 1);
    const size_t nb = (size_t)(3 * fu->cnt_stride + 1) * sizeof(int);
    GPS_CUDA(cudaMalloc(&fu->cnt, nb));
    GPS_CUDA(cudaMemsetAsync(fu->cnt, 0, nb, ctx->stream));
  }
  if (fu->h_cap < nio) {
    if (fu->h_in) cudaFreeHost(fu->h_in);
    if (fu->h_out) cudaFreeHost(fu->h_out);
    fu->h_in = fu->h_out = nullptr;
    GPS_CUDA(cudaHostAlloc(&fu->h_in, nio * sizeof(double), cudaHostAllocMapped));
    GPS_CUDA(cudaHostAlloc(&fu->h_out, nio * sizeof(double), cudaHostAllocMapped));
    fu->h_cap = nio;
  }
  if (!ctx->d_info) GPS_CUDA(cudaMalloc(&ctx->d_info, sizeof(int)));
  const int key = fu->MT * 8 + (fu->KS & 3) * 2 + (fu->DT == 16 ? 1 : 0);
  if (!fu->configured[key & 63]) {
    FUSED_DISPATCH(fu, GPS_CHECK(configure_kernels<CF>(ctx)));
    fu->configured[key & 63] = true;
  }
  return GPS_OK;
}

static FusedArgs make_args(gps_ctx* ctx, gps_fitc_fused* fu, const double* thU, double jitter, int score, int64_t world_n,
                           int finish, bool host_out) {
  FusedArgs a = {};
  a.X = ctx->X.p; a.y = ctx->y.p; a.thU = thU; a.fs = fu->fs.p; a.Kst = fu->Kst.p; a.rowv = fu->rowv.p;
  a.part = fu->part.p; a.gpart = fu->gpart.p; a.cnt = fu->cnt; a.acc1 = fu->acc1.p; a.acc2 = fu->acc2.p;
  a.acc3 = fu->acc3.p; a.out_dev = fu->dout.p; a.out_host = host_out ? fu->h_out : nullptr; a.info = ctx->d_info;
  a.prof = fu->prof;
  a.N = fu->N; a.ldk = fu->ldk; a.D = fu->D; a.M = fu->M; a.score = score; a.finish = finish;
  a.jitter = jitter; a.invN = 1.0 / (double)world_n; a.world_n = (double)world_n;
  return a;
}

// hook for the all-reduce between the passes of a row-sharded evaluation (gps_comm.cu); null = single GPU
typedef int (*gps_allreduce_fn)(gps_ctx* ctx, double* buf, size_t n);

// Enqueue one evaluation at the parameters in `thU` (device or mapped host memory): 3 launches on one GPU,
// 4 + three all-reduces when `allreduce` is given.  No host synchronisation.
int gps_fitc_fused_enqueue(gps_ctx* ctx, const double* thU, double jitter, int score, int64_t world_n, bool want_grad,
                           bool host_out, gps_allreduce_fn allreduce, const double* host_thu = nullptr) {
  gps_fitc_fused* fu = ctx->fu;
  // row-sharded: the exchange runs inside the pass kernels over peer memory when the ranks could map each other's
  // exchange areas (then a sharded evaluation is the same three launches as a single-GPU one), else through the
  // `allreduce` hook (NCCL) between the kernels with the finishing step as a fourth launch
  gps_p2p_view pv;
  const bool p2p = allreduce && gps_comm_p2p_view(ctx, &pv, want_grad ? 3 : 2);
  if (p2p) allreduce = nullptr;
  FusedArgs a = make_args(ctx, fu, thU, jitter, score, world_n, allreduce ? 0 : 1, host_out);
  if (p2p) {
    a.peers = pv.peers; a.rank = pv.rank; a.world = pv.world; a.seq = pv.seq;
  }
  const int MP = fu->MP;
  FUSED_DISPATCH(fu, GPS_CHECK((launch_pass<CF>(ctx, fu, 1, a, host_thu))));
  if (allreduce) GPS_CHECK(allreduce(ctx, fu->acc1.p, (size_t)len1_of(MP)));
  a.seq++;
  FUSED_DISPATCH(fu, GPS_CHECK((launch_pass<CF>(ctx, fu, 2, a))));
  if (allreduce) GPS_CHECK(allreduce(ctx, fu->acc2.p, (size_t)len2_of(MP)));
  a.seq++;
  fu->ready = true;
  if (!want_grad) return GPS_OK;
  FUSED_DISPATCH(fu, GPS_CHECK((launch_pass<CF>(ctx, fu, 3, a))));
  if (allreduce) {
    GPS_CHECK(allreduce(ctx, fu->acc3.p, (size_t)len3_of(MP, fu->DT)));
    FUSED_DISPATCH(fu, GPS_CHECK((launch_pass<CF>(ctx, fu, 4, a))));
  }
  return GPS_OK;
}

static int fused_collect(gps_ctx* ctx, gps_fitc_fused* fu, bool want_grad, double* obj, double* grad_theta, double* grad_U) {
  const int D = fu->D, M = fu->M;
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  const double* h = fu->h_out;
  double objv = h[0];
  int info = 0;
  if (want_grad) {
    info = (int)h[1 + D + 2 + M * D];
  } else {
    // objective only: passes 1 and 2 ran; the objective share and the status are read back directly
    double o = 0.0;
    GPS_CUDA(cudaMemcpyAsync(&o, fu->acc2.p + fu->MP * fu->MP, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GPS_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    objv = o;
  }
  if (info >= 3000000) {
    fu->ready = false;
    return gps_fail(ctx, GPS_ECUDA, "fitc_eval_sharded: rank %d did not reach the peer-memory exchange within the time limit",
                    info - 3000000);
  }
  if (info != 0) {
    fu->ready = false;
    return gps_fail(ctx, GPS_ENOTPD, "fitc: %s not positive definite at pivot %d",
                    info >= 1000000 ? "I + V V'/lambda" : "K_uu + jitter I", info % 1000000);
  }
  if (obj) *obj = objv;
  if (want_grad) {
    if (grad_theta)
      for (int k = 0; k < D + 2; ++k) grad_theta[k] = h[1 + k];
    if (grad_U)
      for (int k = 0; k < M * D; ++k) grad_U[k] = h[1 + D + 2 + k];
  }
  return GPS_OK;
}

// objective-only evaluations of NLML need the log-determinant terms the finishing step adds
static int nlml_objective_terms(gps_ctx* ctx, gps_fitc_fused* fu, int64_t world_n, double* obj) {
  const int MP = fu->MP, M = fu->M;
  const FL fl(MP, fu->D);
  std::vector<double> lc((size_t)MP * MP);
  GPS_CUDA(cudaMemcpyAsync(lc.data(), fu->fs.p + fl.lc, lc.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  double o = *obj + 0.5 * (double)world_n * 1.83787706640934548356;
  for (int m = 0; m < M; ++m) o += log(lc[(size_t)m * MP + m]);
  *obj = o;
  return GPS_OK;
}

int gps_fitc_fused_eval(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                        int64_t world_n, gps_allreduce_fn allreduce, double* obj, double* grad_theta, double* grad_U) {
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_fitc_fused_prepare(ctx, M));
  gps_fitc_fused* fu = ctx->fu;
  const int D = ctx->D;
  for (int k = 0; k < D + 2; ++k) fu->h_in[k] = theta[k];
  for (int k = 0; k < M * D; ++k) fu->h_in[D + 2 + k] = U[k];
  // theta | U travel in pass 1's launch parameters when they fit, else pinned staging buffer -> device (a kernel
  // reading mapped host memory directly pays a PCIe round trip per dependent load: 25 us of pass 1 in the first version)
  const bool inl = D + 2 + M * D <= THU_INLINE;
  if (!inl)
    GPS_CUDA(cudaMemcpyAsync(fu->dthU.p, fu->h_in, (size_t)(D + 2 + M * D) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const bool want_grad = grad_theta || grad_U;
  GPS_CHECK(gps_fitc_fused_enqueue(ctx, fu->dthU.p, jitter, score, world_n, want_grad, true, allreduce, inl ? fu->h_in : nullptr));
  GPS_CHECK(fused_collect(ctx, fu, want_grad, obj, grad_theta, grad_U));
  if (!want_grad && score == GPS_NLML && obj) GPS_CHECK(nlml_objective_terms(ctx, fu, world_n, obj));
  ctx->fitc.pass2_done = true;
  ctx->fitc.loo_ok = true;
  ctx->fitc.large = false;
  ctx->fitc.fused = true;
  ctx->fitc.M = M;
  return GPS_OK;
}

// K20:219-251 in one call with theta and U resident on the device: `iters` x (3 launches + update), one
// synchronisation at the end.
int gps_fitc_fused_descend(gps_ctx* ctx, double* theta, double* U, int M, double jitter, int score, double lr_theta,
                           double lr_u, int iters, double* obj_trace, int64_t world_n, gps_allreduce_fn allreduce) {
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_fitc_fused_prepare(ctx, M));
  gps_fitc_fused* fu = ctx->fu;
  const int D = ctx->D, P = D + 2, Q = M * D;
  if (iters <= 0) return GPS_OK;
  GPS_CHECK(gps_ensure(ctx, fu->trace, (size_t)iters));
  for (int k = 0; k < P; ++k) fu->h_in[k] = theta[k];
  for (int k = 0; k < Q; ++k) fu->h_in[P + k] = U[k];
  int* fail_it = fu->cnt + 3 * fu->cnt_stride;
  const int minus1 = -1;
  GPS_CUDA(cudaMemcpyAsync(fu->dthU.p, fu->h_in, (size_t)(P + Q) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(fail_it, &minus1, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  GPS_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream));
  for (int it = 0; it < iters; ++it) {
    GPS_CHECK(gps_fitc_fused_enqueue(ctx, fu->dthU.p, jitter, score, world_n > 0 ? world_n : ctx->N, true, false, allreduce));
    fused_update_kernel<<<1, 256, 0, ctx->stream>>>(fu->dthU.p, fu->dout.p, D, M, lr_theta, lr_u, fu->trace.p, it,
                                                    ctx->d_info, fail_it);
    GPS_LAUNCH_CHECK();
    ctx->launches++;
  }
  int failed = -1, info = 0;
  GPS_CUDA(cudaMemcpyAsync(fu->h_out, fu->dthU.p, (size_t)(P + Q) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(&failed, fail_it, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (obj_trace)
    GPS_CUDA(cudaMemcpyAsync(obj_trace, fu->trace.p, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < P; ++k) theta[k] = fu->h_out[k];
  for (int k = 0; k < Q; ++k) U[k] = fu->h_out[P + k];
  fu->ready = false;
  ctx->fitc.pass2_done = false;
  if (info >= 3000000)
    return gps_fail(ctx, GPS_ECUDA, "fitc_descend: rank %d did not reach the peer-memory exchange within the time limit", info - 3000000);
  if (info != 0)
    return gps_fail(ctx, GPS_ENOTPD, "fitc_descend: %s not positive definite at pivot %d in iteration %d",
                    info >= 1000000 ? "I + V V'/lambda" : "K_uu + jitter I", info % 1000000, failed);
  return GPS_OK;
}

int gps_fitc_fused_loo(gps_ctx* ctx, double* dm, double* dv) {
  gps_fitc_fused* fu = ctx->fu;
  if (!fu || !fu->ready) return gps_fail(ctx, GPS_ESTATE, "fitc_loo: no evaluation to report");
  fused_loo_kernel<<<(unsigned)((fu->N + 255) / 256), 256, 0, ctx->stream>>>(ctx->y.p, fu->rowv.p, fu->N, fu->ldk, dm, dv);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

int gps_fitc_fused_predict(gps_ctx* ctx, const double* dXs, int64_t T, double* dm, double* dv) {
  gps_fitc_fused* fu = ctx->fu;
  if (!fu || !fu->ready) return gps_fail(ctx, GPS_ESTATE, "fitc_predict: run an evaluation at this theta, U first");
  FUSED_DISPATCH(fu, GPS_CHECK((launch_predict<CF>(ctx, fu, dXs, T, dm, dv))));
  return GPS_OK;
}

// debug: globaltimer stamps (ns) of the phases of the last evaluation; the first call arms the recording
int gps_fitc_fused_phases(gps_ctx* ctx, long long* out48) {
  gps_fitc_fused* fu = ctx->fu;
  if (!fu) return gps_fail(ctx, GPS_ESTATE, "fused phases: no fused evaluation yet");
  GPS_CUDA(cudaSetDevice(ctx->device));
  if (!fu->prof) {
    GPS_CUDA(cudaMalloc(&fu->prof, 48 * sizeof(long long)));
    GPS_CUDA(cudaMemset(fu->prof, 0, 48 * sizeof(long long)));
    for (int k = 0; k < 48; ++k) out48[k] = 0;
    return GPS_OK;
  }
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  GPS_CUDA(cudaMemcpy(out48, fu->prof, 48 * sizeof(long long), cudaMemcpyDeviceToHost));
  return GPS_OK;
}

// Floor of what one evaluation call must do: `launches` empty kernels on the context's stream and one stream
// synchronisation, wall-clock microseconds per repetition (best of `reps`).
int gps_launch_floor_us(gps_ctx* ctx, int launches, int reps, double* us) {
  GPS_CUDA(cudaSetDevice(ctx->device));
  double best = 1e30;
  for (int r = 0; r < reps + 3; ++r) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < launches; ++k) empty_kernel<<<1, 32, 0, ctx->stream>>>();
    GPS_CUDA(cudaStreamSynchronize(ctx->stream));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3;
    if (r >= 3 && dt < best) best = dt;
  }
  *us = best;
  return GPS_OK;
}
