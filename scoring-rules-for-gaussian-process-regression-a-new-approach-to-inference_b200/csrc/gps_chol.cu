// Blocked fp64 Cholesky / triangular inverse / Cholesky-inverse drivers for the full-GP path.
//
// Replaces chol_solve's potrf + two gesv on an N-column identity (KF:25-29, KF:242) by
//   POTRF   K = L L'            right-looking, 128-wide block columns
//   TRTRI   X = L^-1            recursive halving over block ranges, batched per level
//   LAUUM   K^-1 = X' X         one batched launch, lower tiles + mirror
//   SYMPROD S = K^-1 diag(dbar) K^-1   (the N^3 term of the analytic gradient, App. A.1)
// All N^3 work runs in the DMMA tile-GEMM (gps_gemm.cu); only the 128 x 128 diagonal blocks are
// factored and inverted by a dedicated one-CTA kernel below.
#include <algorithm>

#include "gps_common.cuh"

namespace {

constexpr int T = GPS_TILE;
constexpr int LSLD = T + 1;
constexpr size_t POTF2_SMEM = (size_t)T * LSLD * sizeof(double) + 3 * T * sizeof(double);

// One CTA (16 x 16 threads).  Thread (ty, tx) owns the 8 x 8 cyclic sub-block
// A[ty + 16 r][tx + 16 c].  Phase 1: right-looking Cholesky of the 128 x 128 diagonal block with
// the matrix in registers and the current column broadcast through shared memory.  Phase 2:
// forward substitution on the identity (all 128 right-hand sides at once) gives L^-1.
// Writes L (strict upper part zeroed) back to Kd and L^-1 (strict upper part zero) to Xd.
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ K, double* __restrict__ Xinv, int64_t ld, int blk, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* Ls = sm;                       // [128][129]
  double* buf = sm + (size_t)T * LSLD;   // [2][128]
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  double* Kd = K + (int64_t)blk * T * ld + (int64_t)blk * T;
  double* Xd = Xinv + (int64_t)blk * T * ld + (int64_t)blk * T;

  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) a[r][c] = Kd[(int64_t)(ty + 16 * r) * ld + tx + 16 * c];

  // ---- phase 1: Cholesky -----------------------------------------------------------------
#pragma unroll
  for (int kq = 0; kq < 8; ++kq) {
    for (int kc = 0; kc < 16; ++kc) {
      const int k = 16 * kq + kc;
      double* cb = buf + (k & 1) * T;
      if (tx == kc) {
#pragma unroll
        for (int r = 0; r < 8; ++r) cb[ty + 16 * r] = a[r][kq];
      }
      __syncthreads();
      double piv = cb[k];
      if (!(piv > 0.0)) {
        if (tid == 0) atomicCAS(info, 0, blk * T + k + 1);
        piv = 1.0;
      }
      // one reciprocal square root per step keeps the serial chain short: L_kk = piv * rs
      const double rs = rsqrt(piv);
      const double sq = piv * rs;
      double li[8], lc[8];
#pragma unroll
      for (int r = kq; r < 8; ++r) {
        const int i = ty + 16 * r;
        li[r] = (i > k) ? cb[i] * rs : 0.0;
      }
#pragma unroll
      for (int c = kq; c < 8; ++c) {
        const int j = tx + 16 * c;
        lc[c] = (j > k) ? cb[j] * rs : 0.0;
      }
#pragma unroll
      for (int r = kq; r < 8; ++r)
#pragma unroll
        for (int c = kq; c < 8; ++c) a[r][c] -= li[r] * lc[c];
      if (tx == kc) {
#pragma unroll
        for (int r = kq; r < 8; ++r) {
          const int i = ty + 16 * r;
          if (i > k) a[r][kq] = li[r];
          else if (i == k) a[r][kq] = sq;
        }
      }
    }
  }
  // L -> shared + global (zero the strict upper triangle)
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int i = ty + 16 * r, j = tx + 16 * c;
      const double v = (j <= i) ? a[r][c] : 0.0;
      Ls[i * LSLD + j] = v;
      Kd[(int64_t)i * ld + j] = v;
    }
  __syncthreads();
  double* dinv = buf + 2 * T;            // [128] 1 / L_kk, all computed up front (off the serial chain)
  if (tid < T) dinv[tid] = 1.0 / Ls[tid * LSLD + tid];
  __syncthreads();

  // ---- phase 2: X = L^-1 by forward substitution on I ------------------------------------
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) a[r][c] = (ty + 16 * r == tx + 16 * c) ? 1.0 : 0.0;

#pragma unroll
  for (int kq = 0; kq < 8; ++kq) {
    for (int kc = 0; kc < 16; ++kc) {
      const int k = 16 * kq + kc;
      double* rb = buf + (k & 1) * T;
      if (ty == kc) {
        const double inv = dinv[k];
#pragma unroll
        for (int c = 0; c <= kq; ++c) {
          const double x = a[kq][c] * inv;
          a[kq][c] = x;
          rb[tx + 16 * c] = x;
        }
      }
      __syncthreads();
      double li[8], xr[8];
#pragma unroll
      for (int r = kq; r < 8; ++r) {
        const int i = ty + 16 * r;
        li[r] = (i > k) ? Ls[i * LSLD + k] : 0.0;
      }
#pragma unroll
      for (int c = 0; c <= kq; ++c) {
        const int j = tx + 16 * c;
        xr[c] = (j <= k) ? rb[j] : 0.0;
      }
#pragma unroll
      for (int r = kq; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= kq; ++c) a[r][c] -= li[r] * xr[c];
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int i = ty + 16 * r, j = tx + 16 * c;
      Xd[(int64_t)i * ld + j] = (j <= i) ? a[r][c] : 0.0;
    }
}

// ---- diagonal-block kernel, second generation ------------------------------------------------------
// Same contract as potf2_inv_kernel (L and L^-1 of one 128 x 128 diagonal block, one CTA), but
// blocked by 32: the 32 x 32 diagonal sub-blocks are factored and inverted by ONE WARP with the
// matrix rows in registers (pivot broadcast by shuffle: no block barrier on the serial chain), and
// every 32 x 32 x 32 product — panel solve, trailing update, off-diagonal blocks of the inverse —
// runs on DMMA from shared memory.  The lower 4 x 4 block triangle lives in shared memory as ten
// [32][36] blocks (leading dimension 36: conflict-free m8n8k4 fragment loads).
constexpr int SB = 32, SLD = 36, SBLK = SB * SLD;
constexpr size_t POTF2V2_SMEM = (size_t)(10 + 10 + 3) * SBLK * sizeof(double) + 128 * sizeof(double);

__device__ __forceinline__ void dmma4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ int blk_idx(int i, int j) { return i * (i + 1) / 2 + j; }
// c[nf] += sgn * A[strip mf] * B'   (A, B blocks with k contiguous)
__device__ __forceinline__ void strip_nt(const double* __restrict__ A, const double* __restrict__ B, int mf, int lane,
                                         double sgn, double (&c)[4][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const double a = sgn * A[(mf * 8 + g) * SLD + kk * 4 + t];
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) dmma4(c[nf][0], c[nf][1], a, B[(nf * 8 + g) * SLD + kk * 4 + t]);
  }
}
// c[nf] += sgn * A[strip mf] * B    (B block with rows = k, n contiguous)
__device__ __forceinline__ void strip_nn(const double* __restrict__ A, const double* __restrict__ B, int mf, int lane,
                                         double sgn, double (&c)[4][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const double a = sgn * A[(mf * 8 + g) * SLD + kk * 4 + t];
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) dmma4(c[nf][0], c[nf][1], a, B[(kk * 4 + t) * SLD + nf * 8 + g]);
  }
}
__device__ __forceinline__ void strip_store(double* __restrict__ C, int mf, int lane, const double (&c)[4][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nf = 0; nf < 4; ++nf) {
    C[(mf * 8 + g) * SLD + nf * 8 + 2 * t] = c[nf][0];
    C[(mf * 8 + g) * SLD + nf * 8 + 2 * t + 1] = c[nf][1];
  }
}
__device__ __forceinline__ void strip_load(const double* __restrict__ C, int mf, int lane, double (&c)[4][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nf = 0; nf < 4; ++nf) {
    c[nf][0] = C[(mf * 8 + g) * SLD + nf * 8 + 2 * t];
    c[nf][1] = C[(mf * 8 + g) * SLD + nf * 8 + 2 * t + 1];
  }
}

__global__ void __launch_bounds__(256, 1)
potf2_inv_dmma_kernel(double* __restrict__ K, double* __restrict__ Xinv, int64_t ld, int blk, int* __restrict__ info,
                      long long* __restrict__ prof) {
  extern __shared__ __align__(16) double sm[];
  int pslot = 0;
#define GPS_PROF() do { if (prof && threadIdx.x == 0) prof[pslot] = clock64(); ++pslot; } while (0)
  GPS_PROF();
  double* Lb = sm;                    // 10 lower blocks of A -> L
  double* Xb = Lb + 10 * SBLK;        // 10 lower blocks of L^-1
  double* Tb = Xb + 10 * SBLK;        // 3 scratch blocks
  double* Li = Tb + 3 * SBLK;         // [128] reciprocal diagonal of L
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Kd = K + (int64_t)blk * T * ld + (int64_t)blk * T;
  double* Xd = Xinv + (int64_t)blk * T * ld + (int64_t)blk * T;

  // load the lower block triangle: 40 independent coalesced loads per thread (rows of 32 doubles)
#pragma unroll
  for (int b = 0; b < 10; ++b) {
    const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0;
    const int bj = b - bi * (bi + 1) / 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * 256, r = e >> 5, c = e & 31;
      Lb[b * SBLK + r * SLD + c] = Kd[(int64_t)(bi * SB + r) * ld + bj * SB + c];
    }
  }
  __syncthreads();
  GPS_PROF();   // 1: loaded

  for (int kb = 0; kb < 4; ++kb) {
    double* D = Lb + blk_idx(kb, kb) * SBLK;
    // (1) one warp: Cholesky of the 32 x 32 diagonal sub-block, rows in registers
    if (warp == 0) {
      double r[SB];
      double* colbuf = Tb;   // 2 x 32 doubles of the scratch blocks (idle during the factorisation)
#pragma unroll
      for (int j = 0; j < SB; ++j) r[j] = D[lane * SLD + j];
      // The pivot chain runs on a private copy of the row's diagonal entry: dg -= l_j^2 needs only the lane's
      // own l_j, so the next pivot is broadcast (and its rsqrt started) before column j has gone through shared
      // memory for the bulk update — the serial path per pivot is shuffle + rsqrt + two multiplies.
      double dg = D[lane * SLD + lane];
      double li = 0.0;
      int first_bad = -1;
      double piv = __shfl_sync(0xffffffffu, dg, 0);
#pragma unroll
      for (int j = 0; j < SB; ++j) {
        const bool bad = !(piv > 0.0);
        first_bad = (bad && first_bad < 0) ? j : first_bad;
        const double rinv = rsqrt(bad ? 1.0 : piv);   // (an fp32-seeded Newton rsqrt measured slower)
        const double lj = r[j] * rinv;
        li = (lane == j) ? rinv : li;
        dg = fma(-lj, lj, dg);
        if (j + 1 < SB) piv = __shfl_sync(0xffffffffu, dg, j + 1);
        // column j of L goes through shared memory (double-buffered: one __syncwarp per pivot) and comes back as
        // broadcast 16-byte loads: 16 LDS instead of 62 SHFL per pivot
        double* cb = colbuf + (j & 1) * SB;
        cb[lane] = lj;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < SB; c += 2) {          // pairs (c, c + 1); j and c are compile-time after unrolling
          if (c + 1 > j) {
            const double2 l2 = *reinterpret_cast<const double2*>(cb + c);
            if (c > j) r[c] = fma(-lj, l2.x, r[c]);
            r[c + 1] = fma(-lj, l2.y, r[c + 1]);
          }
        }
        r[j] = lj;
      }
      if (first_bad >= 0 && lane == 0) atomicCAS(info, 0, blk * T + kb * SB + first_bad + 1);
#pragma unroll
      for (int j = 0; j < SB; ++j) D[lane * SLD + j] = (j <= lane) ? r[j] : 0.0;
      Li[kb * SB + lane] = li;
    }
    __syncthreads();
    GPS_PROF();   // 2 + 3 kb: sub-block factored
    // (2) panel: solve X L_D' = A(i,kb) row by row (one thread per row, column-oriented
    //     substitution with the row in registers; L_D broadcast from shared memory)
    const int nbelow = 3 - kb;
    if (tid < nbelow * SB) {
      double* P = Lb + blk_idx(kb + 1 + tid / SB, kb) * SBLK + (tid & 31) * SLD;
      double x[SB];
#pragma unroll
      for (int j = 0; j < SB; ++j) x[j] = P[j];
      const volatile double* Dv = D;
      const double* Lik = Li + kb * SB;
#pragma unroll
      for (int j = 0; j < SB; ++j) {
        x[j] *= Lik[j];
#pragma unroll
        for (int c = j + 1; c < SB; ++c) x[c] = fma(-Dv[c * SLD + j], x[j], x[c]);
      }
#pragma unroll
      for (int j = 0; j < SB; ++j) P[j] = x[j];
    }
    __syncthreads();
    GPS_PROF();   // 3 + 3 kb: panel solved
    // (3) trailing update inside the 128 block: A(i,j) -= L(i,kb) L(j,kb)'
    const int ntr = nbelow * (nbelow + 1) / 2;
    for (int item = warp; item < ntr * 4; item += 8) {
      const int bidx = item / 4, mf = item & 3;
      const int ii = (bidx >= 3) ? 2 : (bidx >= 1) ? 1 : 0;
      const int jj = bidx - ii * (ii + 1) / 2;
      const int i = kb + 1 + ii, j = kb + 1 + jj;
      double* Cb = Lb + blk_idx(i, j) * SBLK;
      double c[4][2];
      strip_load(Cb, mf, lane, c);
      strip_nt(Lb + blk_idx(i, kb) * SBLK, Lb + blk_idx(j, kb) * SBLK, mf, lane, -1.0, c);
      strip_store(Cb, mf, lane, c);
    }
    __syncthreads();
    GPS_PROF();   // 4 + 3 kb: trailing blocks updated
  }
  // inverses of the four diagonal sub-blocks, one warp each: lane c solves L x = e_c
  if (warp < 4) {
    const double* D = Lb + blk_idx(warp, warp) * SBLK;
    double* XD = Xb + blk_idx(warp, warp) * SBLK;
    const double* Lik = Li + warp * SB;
    double x[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
    const volatile double* Dv = D;
#pragma unroll
    for (int i = 0; i < SB; ++i) {
      x[i] *= Lik[i];
#pragma unroll
      for (int k2 = i + 1; k2 < SB; ++k2) x[k2] = fma(-Dv[k2 * SLD + i], x[i], x[k2]);
    }
#pragma unroll
    for (int i = 0; i < SB; ++i) XD[i * SLD + lane] = (i >= lane) ? x[i] : 0.0;
  }
  __syncthreads();
  GPS_PROF();   // 14: sub-block inverses
  // (4) off-diagonal blocks of X = L^-1 by block diagonals: X(i,j) = -X(i,i) sum_k L(i,k) X(k,j)
  for (int d = 1; d < 4; ++d) {
    const int nblk = 4 - d;
    for (int item = warp; item < nblk * 4; item += 8) {
      const int i = d + item / 4, j = i - d, mf = item & 3;
      double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      for (int k2 = j; k2 < i; ++k2) strip_nn(Lb + blk_idx(i, k2) * SBLK, Xb + blk_idx(k2, j) * SBLK, mf, lane, 1.0, c);
      strip_store(Tb + (i - d) * SBLK, mf, lane, c);
    }
    __syncthreads();
    for (int item = warp; item < nblk * 4; item += 8) {
      const int i = d + item / 4, j = i - d, mf = item & 3;
      double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      strip_nn(Xb + blk_idx(i, i) * SBLK, Tb + (i - d) * SBLK, mf, lane, -1.0, c);
      strip_store(Xb + blk_idx(i, j) * SBLK, mf, lane, c);
    }
    __syncthreads();
  }
  GPS_PROF();   // 15: off-diagonal blocks of the inverse
  // write back: full 128 x 128 tiles of L and L^-1 with explicit zeros above the diagonal
#pragma unroll 8
  for (int q = 0; q < T * T / 512; ++q) {          // 16-byte stores: two columns per thread
    const int e = (tid + q * 256) * 2;
    const int r = e >> 7, c = e & 127;
    const int bi = r >> 5, bj = c >> 5;
    double2 lv = make_double2(0.0, 0.0), xv = make_double2(0.0, 0.0);
    if (bj <= bi) {
      const int o = blk_idx(bi, bj) * SBLK + (r & 31) * SLD + (c & 31);
      lv = *reinterpret_cast<const double2*>(Lb + o);
      xv = *reinterpret_cast<const double2*>(Xb + o);
    }
    *reinterpret_cast<double2*>(Kd + (int64_t)r * ld + c) = lv;
    *reinterpret_cast<double2*>(Xd + (int64_t)r * ld + c) = xv;
  }
  __syncthreads();
  GPS_PROF();   // 16: written back
#undef GPS_PROF
}

struct Node { int lo, mid, hi, depth; };

// Split point of a TRTRI node.  Halving puts 39 % of all inversion flops into the top node's X phase
// (X21 = -X22 P), which cannot start before the whole right half is factored AND inverted: with the merges
// overlapped behind POTRF that phase is a 4.8 ms tail after the factorisation has finished (lane trace, round 2).
// Large nodes are therefore split off-centre (pct % of the tiles to the left, rounded to the POTRF outer block so the
// early P phase is released at an outer-step boundary): the late X phase shrinks with the square of the right part,
// the work moves into the P phase, which runs while POTRF still leaves SMs idle.
int split_point(int lo, int hi, int pct, int ob) {
  const int n = hi - lo;
  if (n < 16 || pct == 50) return lo + n / 2;
  int mid = lo + (n * pct + 50) / 100;
  const int al = (mid + ob / 2) / ob * ob;
  if (al > lo && al < hi) mid = al;
  if (mid <= lo) mid = lo + 1;
  if (mid >= hi) mid = hi - 1;
  return mid;
}

void collect(int lo, int hi, int depth, int pct, int ob, std::vector<Node>& out) {
  if (hi - lo <= 1) return;
  const int mid = split_point(lo, hi, pct, ob);
  out.push_back({lo, mid, hi, depth});
  collect(lo, mid, depth + 1, pct, ob, out);
  collect(mid, hi, depth + 1, pct, ob, out);
}

}  // namespace

int gps_build_tasks(gps_ctx* ctx, int64_t Np) {
  if (ctx->ws_Np == Np && ctx->d_tasks) return GPS_OK;
  const int nb = (int)(Np / T);
  std::vector<GemmTask>& h = ctx->h_tasks;
  h.clear();
  auto push = [&](int a_row, int b_row, int k0, int k1, int ci, int cj) {
    GemmTask t;
    t.a_row = a_row; t.b_row = b_row; t.k0 = k0; t.k1 = k1; t.c_row = ci * T; t.c_col = cj * T;
    t.flags = 0; t.pad = 0;
    h.push_back(t);
  };
  // POTRF, two-level: outer block columns of OB tiles.  Inside a block column the 128-wide
  // steps (diagonal kernel, panel solve, update of the remaining inner columns) only touch that
  // block column; the rest of the matrix receives ONE update per outer step with k = OB*128,
  // split into the part the next block column needs (A) and the remainder (B) so that the next
  // block column can be factored on a second stream while B runs (look-ahead).
  const int OB = ctx->potrf_ob;
  const int no = (nb + OB - 1) / OB;
  ctx->potrf_panel.assign(nb, {});
  ctx->potrf_inner.assign(nb, {});
  ctx->potrf_innerB.assign(nb, {});
  ctx->potrf_innerL.assign(nb, {});
  ctx->potrf_trailA.assign(no, {});
  ctx->potrf_trailA1.assign(no, {});
  ctx->potrf_trailB.assign(no, {});
  for (int o = 0; o < no; ++o) {
    const int c0 = o * OB, c1 = std::min(nb, c0 + OB);
    for (int k = c0; k < c1; ++k) {
      ctx->potrf_panel[k].off = h.size();
      for (int i = k + 1; i < nb; ++i) push(i * T, k * T, k * T, (k + 1) * T, i, k);
      ctx->potrf_panel[k].cnt = h.size() - ctx->potrf_panel[k].off;
      // inner update, rows of the diagonal block (on the chain) and rows below it (second panel stream)
      ctx->potrf_inner[k].off = h.size();
      for (int j = k + 1; j < c1; ++j)
        for (int i = j; i < c1; ++i) push(i * T, j * T, k * T, (k + 1) * T, i, j);
      ctx->potrf_inner[k].cnt = h.size() - ctx->potrf_inner[k].off;
      ctx->potrf_innerB[k].off = h.size();
      for (int j = k + 1; j < c1; ++j)
        for (int i = c1; i < nb; ++i) push(i * T, j * T, k * T, (k + 1) * T, i, j);
      ctx->potrf_innerB[k].cnt = h.size() - ctx->potrf_innerB[k].off;
      // the same update of the rows below, left-looking: tile column k receives its block-internal updates in one
      // launch with k-range [c0, k) tiles right before its panel solve (A/B knob 13)
      ctx->potrf_innerL[k].off = h.size();
      if (k > c0)
        for (int i = c1; i < nb; ++i) push(i * T, k * T, c0 * T, k * T, i, k);
      ctx->potrf_innerL[k].cnt = h.size() - ctx->potrf_innerL[k].off;
    }
    const int n1 = std::min(nb, c1 + OB);
    // part A of the trailing update = the next block column, its diagonal 512 x 512 block (what the next chain
    // waits for: <= 10 tiles) ahead of the rows below it (what the next below-lane waits for)
    ctx->potrf_trailA[o].off = h.size();
    for (int j = c1; j < n1; ++j)
      for (int i = j; i < n1; ++i) push(i * T, j * T, c0 * T, c1 * T, i, j);
    ctx->potrf_trailA[o].cnt = h.size() - ctx->potrf_trailA[o].off;
    ctx->potrf_trailA1[o].off = h.size();
    for (int j = c1; j < n1; ++j)
      for (int i = n1; i < nb; ++i) push(i * T, j * T, c0 * T, c1 * T, i, j);
    ctx->potrf_trailA1[o].cnt = h.size() - ctx->potrf_trailA1[o].off;
    // (super-tile order for this list and for LAUUM was measured in round 2: DRAM traffic of the (KC, KC) launches
    // 14.3 -> 13.2 GB and LAUUM 4.0 -> 3.0 GB per evaluation, but LAUUM 10.41 -> 10.65 ms and no gain here — the
    // launches are DMMA-bound at ~5 % of HBM bandwidth, so the plain orders stay)
    ctx->potrf_trailB[o].off = h.size();
    for (int j = n1; j < nb; ++j)
      for (int i = j; i < nb; ++i) push(i * T, j * T, c0 * T, c1 * T, i, j);
    ctx->potrf_trailB[o].cnt = h.size() - ctx->potrf_trailB[o].off;
  }
  // TRTRI: nodes by depth, deepest first
  std::vector<Node> nodes;
  collect(0, nb, 0, ctx->trtri_split_pct, OB, nodes);
  int maxd = -1;
  for (auto& n : nodes) maxd = std::max(maxd, n.depth);
  ctx->trtri_p.clear();
  ctx->trtri_x.clear();
  for (int d = maxd; d >= 0; --d) {
    gps_ctx::Range rp, rx;
    rp.off = h.size();
    for (auto& n : nodes)
      if (n.depth == d)
        for (int j = n.lo; j < n.mid; ++j)          // longest k-range first within a node
          for (int i = n.mid; i < n.hi; ++i) push(i * T, j * T, j * T, n.mid * T, i, j);
    rp.cnt = h.size() - rp.off;
    rx.off = h.size();
    for (auto& n : nodes)
      if (n.depth == d)
        for (int i = n.hi - 1; i >= n.mid; --i)
          for (int j = n.lo; j < n.mid; ++j) push(i * T, j * T, n.mid * T, (i + 1) * T, i, j);
    rx.cnt = h.size() - rx.off;
    ctx->trtri_p.push_back(rp);
    ctx->trtri_x.push_back(rx);
  }
  // TRTRI again, in the order the overlapped driver issues it.  Work is released in few, large
  // launches so that it can fill the idle SMs of the POTRF tail (eagerly issuing every small merge
  // serialises hundreds of few-CTA launches on the low-priority stream and fills nothing): for a node
  // [lo, mid, hi) the whole left subtree, batched by depth, and the node's P phase (L21 * X11) are
  // issued after the outer step that factors tile column mid-1; the right subtree is scheduled the
  // same way recursively; the X phases (-X22 * P) follow at the step that finishes the right half.
  ctx->trtri_sched.clear();
  {
    auto p_tasks = [&](const Node& n) {
      for (int j = n.lo; j < n.mid; ++j)
        for (int i = n.mid; i < n.hi; ++i) push(i * T, j * T, j * T, n.mid * T, i, j);
    };
    auto x_rows = [&](const Node& n, int r0, int r1) {       // rows [r0, r1) of the node's X21 = -X22 * P
      for (int i = r1 - 1; i >= r0; --i)
        for (int j = n.lo; j < n.mid; ++j) push(i * T, j * T, n.mid * T, (i + 1) * T, i, j);
    };
    auto x_tasks = [&](const Node& n) { x_rows(n, n.mid, n.hi); };
    auto emit = [&](int step, int phase, size_t off) {
      gps_ctx::TriLaunch tl{step, phase, {}};
      tl.r.off = off;
      tl.r.cnt = h.size() - off;
      if (tl.r.cnt) ctx->trtri_sched.push_back(tl);
    };
    auto subtree = [&](int lo, int hi, int step) {   // every node inside [lo, hi), deepest level first
      for (int d = maxd; d >= 0; --d) {
        size_t off = h.size();
        for (auto& n : nodes)
          if (n.depth == d && n.lo >= lo && n.hi <= hi) p_tasks(n);
        emit(step, 0, off);
        off = h.size();
        for (auto& n : nodes)
          if (n.depth == d && n.lo >= lo && n.hi <= hi) x_tasks(n);
        emit(step, 1, off);
      }
    };
    // walk down the right spine of the tree
    std::vector<Node> spine;
    int lo = 0, hi = nb;
    // Row-wise release of the spine's X phases (knob 15): rows [lo, mid) of EVERY ancestor's X21 need only the
    // ancestor's P (released long ago), the inverse of the diagonal block [lo, mid) (the sub-tree just emitted) and
    // the same rows of the deeper ancestors' X21 — so they go out here, deepest ancestor first, instead of after
    // the whole right half is inverted.  The late work after POTRF's last step shrinks from the top node's full X
    // phase (39 % of all inversion flops) to the last row group's share.
    const bool rowwise = ctx->trtri_rowwise != 0;
    while (hi - lo > 1) {
      const int mid = split_point(lo, hi, ctx->trtri_split_pct, OB);
      Node n{lo, mid, hi, 0};
      const int step = (mid - 1) / OB;
      subtree(lo, mid, step);
      size_t off = h.size();
      p_tasks(n);
      emit(step, 0, off);
      if (rowwise)
        for (int s = (int)spine.size() - 1; s >= 0; --s) {
          off = h.size();
          x_rows(spine[s], lo, mid);
          emit(step, 1, off);
        }
      spine.push_back(n);
      lo = mid;
    }
    for (int s = (int)spine.size() - 1; s >= 0; --s) {
      size_t off = h.size();
      if (rowwise) x_rows(spine[s], lo, hi);          // the last leaf's rows
      else x_tasks(spine[s]);
      emit((spine[s].hi - 1) / OB, 1, off);
    }
  }
  // LAUUM: lower tiles, k in [i, nb); longest first
  ctx->lauum.off = h.size();
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j <= i; ++j) push(i * T, j * T, i * T, nb * T, i, j);
  ctx->lauum.cnt = h.size() - ctx->lauum.off;
  // SYMPROD: lower tiles, full k.  Tiles are issued super-tile by super-tile (SUPER x SUPER tiles
  // ~ one wave of CTAs), so that the CTAs resident at the same time share 2 * SUPER operand
  // panels through L2 instead of streaming ~150 different ones from HBM.
  ctx->symprod.off = h.size();
  {
    const int SUPER = 12;
    for (int I = 0; I < nb; I += SUPER)
      for (int J = 0; J <= I; J += SUPER)
        for (int i = I; i < std::min(nb, I + SUPER); ++i)
          for (int j = J; j < std::min(nb, J + SUPER) && j <= i; ++j) push(i * T, j * T, 0, nb * T, i, j);
  }
  ctx->symprod.cnt = h.size() - ctx->symprod.off;

  if (h.size() > ctx->tasks_cap) {
    if (ctx->d_tasks) cudaFree(ctx->d_tasks);
    ctx->d_tasks = nullptr;
    GPS_CUDA(cudaMalloc(&ctx->d_tasks, h.size() * sizeof(GemmTask)));
    ctx->tasks_cap = h.size();
  }
  GPS_CUDA(cudaMemcpyAsync(ctx->d_tasks, h.data(), h.size() * sizeof(GemmTask), cudaMemcpyHostToDevice,
                           ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPS_OK;
}

namespace {

int ensure_potrf_streams(gps_ctx* ctx, int nb, int no) {
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM));
    GPS_CUDA(cudaFuncSetAttribute(potf2_inv_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)POTF2V2_SMEM));
    configured = true;
  }
  if (!ctx->panel_stream) {
    int lo = 0, hi = 0;
    GPS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least, hi = greatest (numerically smaller)
    GPS_CUDA(cudaStreamCreateWithPriority(&ctx->panel_stream, cudaStreamNonBlocking, hi));
    GPS_CUDA(cudaStreamCreateWithPriority(&ctx->panel2_stream, cudaStreamNonBlocking, hi));
    const int mid = (lo - 1 >= hi) ? lo - 1 : lo;
    GPS_CUDA(cudaStreamCreateWithPriority(&ctx->trail_stream, cudaStreamNonBlocking, mid));
    GPS_CUDA(cudaStreamCreateWithPriority(&ctx->tri_stream, cudaStreamNonBlocking, lo));
    GPS_CUDA(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    GPS_CUDA(cudaEventCreateWithFlags(&ctx->join_trail_ev, cudaEventDisableTiming));
    GPS_CUDA(cudaEventCreateWithFlags(&ctx->join_tri_ev, cudaEventDisableTiming));
  }
  auto grow = [&](std::vector<cudaEvent_t>& v, size_t n) -> int {
    while (v.size() < n) {
      cudaEvent_t e;
      GPS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      v.push_back(e);
    }
    return GPS_OK;
  };
  GPS_CHECK(grow(ctx->potrf_events, (size_t)2 * no + 2));
  GPS_CHECK(grow(ctx->tile_events, (size_t)nb));
  GPS_CHECK(grow(ctx->below_events, (size_t)no));
  GPS_CHECK(grow(ctx->trailA1_events, (size_t)no));
  return GPS_OK;
}

// Blocked right-looking POTRF, two levels (outer block columns of OB tiles), three concurrent lanes:
//   chain   (s_pan,  highest priority): for each tile column k of the block column — diagonal kernel,
//           panel solve and inner update of the rows INSIDE the diagonal 512 x 512 block only
//           (<= 3 + 6 tiles): the serial path is as short as it can be with 128-wide steps;
//   below   (s_pan2, highest priority): the same panel solve / inner update for the rows below the
//           diagonal block, one step behind the chain (tile event k);
//   trail   (s_trail): the k = 512 update of the rest of the matrix, part A (what the next block column
//           needs) first so that the next chain starts while part B still runs (look-ahead).
// with_trtri: the inversion merges are released on a fourth, lowest-priority stream as their operands
// become final (see gps_build_tasks); s_trail is then an internal mid-priority stream, otherwise the
// caller's stream.
int trace_mark(gps_ctx* ctx, int code, cudaStream_t s) {
  if (!ctx->trace_on) return GPS_OK;
  if (ctx->trace_used == ctx->trace.size()) {
    cudaEvent_t e;
    GPS_CUDA(cudaEventCreate(&e));
    ctx->trace.push_back({code, e});
  }
  ctx->trace[ctx->trace_used].first = code;
  GPS_CUDA(cudaEventRecord(ctx->trace[ctx->trace_used].second, s));
  ctx->trace_used++;
  return GPS_OK;
}

// Launch a list of EQUAL-length tasks so that a partial last wave does not cost a whole wave time: the tiles
// beyond the last full wave of sm_count tasks go into a second launch with the 32-row strip policy (four CTAs of
// half the duration per tile), which slot in as the last full wave drains.
int gemm_equal_tasks(gps_ctx* ctx, const double* A, const double* B, double* C, int64_t Np, double alpha, double beta,
                     const GemmTask* tasks, size_t cnt) {
  const size_t wave = (size_t)ctx->sm_count;
  size_t rest = cnt % wave;
  if (cnt < 4 * wave || rest * 3 > wave * 2) rest = 0;   // small launches / a well-filled last wave: one launch
  GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_KC, A, Np, B, Np, C, Np, alpha, beta, nullptr, false, tasks, cnt - rest));
  if (rest) {
    const int saved = ctx->gemm_strip_policy;
    ctx->gemm_strip_policy = 32;
    const int r = gps_gemm_tasks(ctx, GEMM_KC_KC, A, Np, B, Np, C, Np, alpha, beta, nullptr, false, tasks + (cnt - rest), rest);
    ctx->gemm_strip_policy = saved;
    GPS_CHECK(r);
  }
  return GPS_OK;
}

// restores the context's stream and strip policy on every exit path of the factorisation driver (its GPS_CUDA /
// GPS_CHECK macros return from the middle of the lane loop)
struct LaneGuard {
  gps_ctx* c;
  cudaStream_t s;
  int policy;
  explicit LaneGuard(gps_ctx* ctx) : c(ctx), s(ctx->stream), policy(ctx->gemm_strip_policy) {}
  ~LaneGuard() { c->stream = s; c->gemm_strip_policy = policy; c->gemm_grid_cap = 0; }
};

int potrf_lanes(gps_ctx* ctx, double* K, double* Xinv, double* scratch, int64_t Np, bool with_trtri) {
  const int nb = (int)(Np / T);
  const int OB = ctx->potrf_ob;
  const int no = (nb + OB - 1) / OB;
  GPS_CHECK(ensure_potrf_streams(ctx, nb, no));
  // overlap_trtri == 2 (debug): every lane on the caller's stream, so that launches run one at a time and
  // CUDA events around a launch bracket only that launch (the kernel-timing pass of bench.py)
  const bool serial = ctx->overlap_trtri == 2;
  cudaStream_t s_user = ctx->stream, s_tri = ctx->tri_stream;
  cudaStream_t s_pan = serial ? s_user : ctx->panel_stream, s_pan2 = serial ? s_user : ctx->panel2_stream;
  cudaStream_t s_trail = with_trtri ? ctx->trail_stream : s_user;
  GPS_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), s_user));
  if (with_trtri) {
    GPS_CUDA(cudaEventRecord(ctx->fork_ev, s_user));
    GPS_CUDA(cudaStreamWaitEvent(s_trail, ctx->fork_ev, 0));
    GPS_CUDA(cudaStreamWaitEvent(s_tri, ctx->fork_ev, 0));
  }
  // event 2*o: block column o has received every update (s_trail); 2*o + 1: its diagonal block chain is done
  // (s_pan); tile_events[k]: tile column k factored inside the diagonal block; below_events[o]: rows below done
  GPS_CUDA(cudaEventRecord(ctx->potrf_events[0], s_trail));
  ctx->trace_used = 0;
  GPS_CHECK(trace_mark(ctx, 0, s_trail));
  int rc = GPS_OK;
  size_t next_tri = 0;
  auto gemm_nt = [&](const double* B, double alpha, double beta, const gps_ctx::Range& r, size_t skip, size_t cnt) {
    return gps_gemm_tasks(ctx, GEMM_KC_KC, K, Np, B, Np, K, Np, alpha, beta, nullptr, false,
                          ctx->d_tasks + r.off + skip, cnt);
  };
  for (int o = 0; o < no && rc == GPS_OK; ++o) {
    const int c0 = o * OB, c1 = std::min(nb, c0 + OB);
    GPS_CUDA(cudaStreamWaitEvent(s_pan, ctx->potrf_events[2 * o], 0));
    GPS_CUDA(cudaStreamWaitEvent(s_pan2, o ? ctx->trailA1_events[o - 1] : ctx->potrf_events[0], 0));
    for (int k = c0; k < c1 && rc == GPS_OK; ++k) {
      const size_t n_in = (size_t)(c1 - k - 1);   // panel tiles inside the diagonal block come first in the list
      ctx->stream = s_pan;
      if (ctx->potf2_variant == 0)
        potf2_inv_kernel<<<1, 256, POTF2_SMEM, s_pan>>>(K, Xinv, Np, k, ctx->d_info);
      else
        potf2_inv_dmma_kernel<<<1, 256, POTF2V2_SMEM, s_pan>>>(K, Xinv, Np, k, ctx->d_info, ctx->potf2_prof);
      if (cudaGetLastError() != cudaSuccess) rc = gps_fail(ctx, GPS_ECUDA, "potf2 launch failed");
      ctx->launches++;
      // panel: L_ik = A_ik * inv(L_kk)';  inner: A_ij -= L_ik L_jk' for the remaining columns of the block column
      ctx->gemm_strip_policy = ctx->chain_strip;
      if (rc == GPS_OK) rc = gemm_nt(Xinv, 1.0, 0.0, ctx->potrf_panel[k], 0, n_in);
      if (rc == GPS_OK) rc = gemm_nt(K, -1.0, 1.0, ctx->potrf_inner[k], 0, ctx->potrf_inner[k].cnt);
      ctx->gemm_strip_policy = 0;
      if (rc != GPS_OK) break;
      GPS_CUDA(cudaEventRecord(ctx->tile_events[k], s_pan));
      GPS_CHECK(trace_mark(ctx, 5000 + k, s_pan));
      ctx->stream = s_pan2;
      if (ctx->potrf_left && ctx->potrf_innerL[k].cnt)    // needs only rows / columns that were final one tile step ago
        rc = gemm_nt(K, -1.0, 1.0, ctx->potrf_innerL[k], 0, ctx->potrf_innerL[k].cnt);
      GPS_CUDA(cudaStreamWaitEvent(s_pan2, ctx->tile_events[k], 0));
      if (rc == GPS_OK) rc = gemm_nt(Xinv, 1.0, 0.0, ctx->potrf_panel[k], n_in, ctx->potrf_panel[k].cnt - n_in);
      if (rc == GPS_OK && !ctx->potrf_left) rc = gemm_nt(K, -1.0, 1.0, ctx->potrf_innerB[k], 0, ctx->potrf_innerB[k].cnt);
      if (rc == GPS_OK) rc = trace_mark(ctx, 6000 + k, s_pan2);
    }
    if (rc != GPS_OK) break;
    GPS_CUDA(cudaEventRecord(ctx->potrf_events[2 * o + 1], s_pan));
    GPS_CUDA(cudaEventRecord(ctx->below_events[o], s_pan2));
    GPS_CHECK(trace_mark(ctx, 1000 + o, s_pan));
    GPS_CHECK(trace_mark(ctx, 2000 + o, s_pan2));
    // trailing update from block column o
    ctx->stream = s_trail;
    GPS_CUDA(cudaStreamWaitEvent(s_trail, ctx->potrf_events[2 * o + 1], 0));
    GPS_CUDA(cudaStreamWaitEvent(s_trail, ctx->below_events[o], 0));
    rc = gemm_nt(K, -1.0, 1.0, ctx->potrf_trailA[o], 0, ctx->potrf_trailA[o].cnt);
    if (rc != GPS_OK) break;
    GPS_CUDA(cudaEventRecord(ctx->potrf_events[2 * o + 2], s_trail));
    GPS_CHECK(trace_mark(ctx, 7000 + o, s_trail));
    ctx->gemm_grid_cap = ctx->cap_trail;
    rc = gemm_nt(K, -1.0, 1.0, ctx->potrf_trailA1[o], 0, ctx->potrf_trailA1[o].cnt);
    if (rc != GPS_OK) { ctx->gemm_grid_cap = 0; break; }
    GPS_CUDA(cudaEventRecord(ctx->trailA1_events[o], s_trail));
    rc = gemm_nt(K, -1.0, 1.0, ctx->potrf_trailB[o], 0, ctx->potrf_trailB[o].cnt);
    ctx->gemm_grid_cap = 0;   // (gemm_equal_tasks here: no gain, the other lanes fill the tail)
    if (rc == GPS_OK) rc = trace_mark(ctx, 3000 + o, s_trail);
    if (rc != GPS_OK || !with_trtri) continue;
    // inversion merges whose operands are final after this step
    ctx->stream = s_tri;
    GPS_CUDA(cudaStreamWaitEvent(s_tri, ctx->potrf_events[2 * o + 1], 0));
    GPS_CUDA(cudaStreamWaitEvent(s_tri, ctx->below_events[o], 0));
    ctx->gemm_strip_policy = ctx->tri_strip;
    ctx->gemm_grid_cap = ctx->cap_trtri;
    while (next_tri < ctx->trtri_sched.size() && ctx->trtri_sched[next_tri].step == o && rc == GPS_OK) {
      const auto& tl = ctx->trtri_sched[next_tri++];
      if (tl.phase == 0)   // P = L21 * X11
        rc = gps_gemm_tasks(ctx, GEMM_KC_MC, K, Np, Xinv, Np, scratch, Np, 1.0, 0.0, nullptr, false,
                            ctx->d_tasks + tl.r.off, tl.r.cnt);
      else                 // X21 = -X22 * P
        rc = gps_gemm_tasks(ctx, GEMM_KC_MC, Xinv, Np, scratch, Np, Xinv, Np, -1.0, 0.0, nullptr, false,
                            ctx->d_tasks + tl.r.off, tl.r.cnt);
    }
    ctx->gemm_strip_policy = 0;
    ctx->gemm_grid_cap = 0;
    if (rc == GPS_OK) rc = trace_mark(ctx, 4000 + o, s_tri);
  }
  return rc;
}

int potrf_driver(gps_ctx* ctx, double* K, double* Xinv, double* scratch, int64_t Np, bool with_trtri) {
  cudaStream_t s_user = ctx->stream;
  int rc;
  {
    LaneGuard guard(ctx);
    rc = potrf_lanes(ctx, K, Xinv, scratch, Np, with_trtri);
  }
  // join every forked lane into the caller's stream, also after an error: later calls on this context must not
  // overtake work still queued on the internal streams
  int jrc = GPS_OK;
  for (cudaStream_t s : {ctx->panel_stream, ctx->panel2_stream, ctx->trail_stream, ctx->tri_stream}) {
    if (!s || s == s_user) continue;
    cudaEvent_t ev = (s == ctx->tri_stream) ? ctx->join_tri_ev : ctx->join_trail_ev;
    if (!ev) continue;
    if (cudaEventRecord(ev, s) != cudaSuccess || cudaStreamWaitEvent(s_user, ev, 0) != cudaSuccess) {
      cudaGetLastError();
      if (jrc == GPS_OK) jrc = gps_fail(ctx, GPS_ECUDA, "potrf: joining the factorisation lanes failed");
    }
  }
  return rc != GPS_OK ? rc : jrc;
}

}  // namespace

int gps_potrf(gps_ctx* ctx, double* K, double* Xinv, int64_t Np) {
  return potrf_driver(ctx, K, Xinv, nullptr, Np, false);
}

// POTRF with the TRTRI merges issued on a lowest-priority stream right after the outer step that
// finalises their operands: the tail of POTRF (serial chain of diagonal blocks, trailing updates too
// small to fill the GPU) is filled with inversion work that would otherwise start after it.
int gps_potrf_trtri(gps_ctx* ctx, double* K, double* Xinv, double* scratch, int64_t Np) {
  const int no = (int)((Np / T + ctx->potrf_ob - 1) / ctx->potrf_ob);
  if (ctx->overlap_trtri != 1 || no < 3) {
    GPS_CHECK(gps_potrf(ctx, K, Xinv, Np));
    return gps_trtri(ctx, K, Xinv, scratch, Np);
  }
  return potrf_driver(ctx, K, Xinv, scratch, Np, true);
}

int gps_check_info(gps_ctx* ctx) {
  int info = 0;
  GPS_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info != 0)
    return gps_fail(ctx, GPS_ENOTPD, "matrix not positive definite: leading minor of order %d", info);
  return GPS_OK;
}

int gps_trtri(gps_ctx* ctx, const double* L, double* Xinv, double* scratch, int64_t Np) {
  for (size_t lv = 0; lv < ctx->trtri_p.size(); ++lv) {
    // P = L21 * X11
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, L, Np, Xinv, Np, scratch, Np, 1.0, 0.0, nullptr, false,
                             ctx->d_tasks + ctx->trtri_p[lv].off, ctx->trtri_p[lv].cnt));
    // X21 = -X22 * P
    GPS_CHECK(gps_gemm_tasks(ctx, GEMM_KC_MC, Xinv, Np, scratch, Np, Xinv, Np, -1.0, 0.0, nullptr, false,
                             ctx->d_tasks + ctx->trtri_x[lv].off, ctx->trtri_x[lv].cnt));
  }
  return GPS_OK;
}

int gps_lauum(gps_ctx* ctx, const double* Xinv, double* Kinv, int64_t Np) {
  return gps_gemm_tasks(ctx, GEMM_MC_MC, Xinv, Np, Xinv, Np, Kinv, Np, 1.0, 0.0, nullptr, true,
                        ctx->d_tasks + ctx->lauum.off, ctx->lauum.cnt);
}

namespace {
// T = Kinv * diag(dvec): 16-byte loads/stores over the full matrix
__global__ void __launch_bounds__(256)
scale_cols_kernel(const double* __restrict__ A, const double* __restrict__ dvec, double* __restrict__ Tm, int64_t Np) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (e >= Np * Np) return;
  const int64_t c = e % Np;
  const double2 a = *reinterpret_cast<const double2*>(A + e);
  const double2 d = *reinterpret_cast<const double2*>(dvec + c);
  double2 o;
  o.x = a.x * d.x;
  o.y = a.y * d.y;
  *reinterpret_cast<double2*>(Tm + e) = o;
}
}  // namespace

// S = Kinv diag(dvec) Kinv (lower tiles).  The k-scaling can ride in the GEMM's fragment loop (scratch ==
// nullptr), but that costs 5 % of the DMMA rate (32.4 vs 34.2 TFLOP/s at n = 10112, tools/symprod_probe.py);
// one HBM-bound pass that writes T = Kinv diag(dvec) (0.3 ms at N = 10^4) and a plain T * Kinv' is faster.
int gps_symprod(gps_ctx* ctx, const double* Kinv, const double* dvec, double* scratch, double* S, int64_t Np) {
  if (!scratch)
    return gps_gemm_tasks(ctx, GEMM_KC_KC, Kinv, Np, Kinv, Np, S, Np, 1.0, 0.0, dvec, false,
                          ctx->d_tasks + ctx->symprod.off, ctx->symprod.cnt);
  const int64_t pairs = Np * Np / 2;
  scale_cols_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, ctx->stream>>>(Kinv, dvec, scratch, Np);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  // equal-length tasks (full k): 3160 tiles = 21.35 lock-step waves at N = 10^4 — see gemm_equal_tasks
  return gemm_equal_tasks(ctx, scratch, Kinv, S, Np, 1.0, 0.0, ctx->d_tasks + ctx->symprod.off, ctx->symprod.cnt);
}
