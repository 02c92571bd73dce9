// Task-list FP64 tile GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64) for sm_100a.
//
// Every dense N^3 stage of the full-GP evaluation is expressed as a list of independent
// 128 x 128 output tiles ("tasks"), each a contraction over a k-range of two operand panels:
//   POTRF panel solve / trailing update   C (-)= A B'        both operands k-contiguous  (KC,KC)
//   TRTRI recursion                       C  =  A B          A k-contiguous, B n-contiguous (KC,MC)
//   LAUUM  K^-1 = L^-T L^-1               C  =  A' B         both m/n-contiguous          (MC,MC)
//   S = K^-1 diag(dbar) K^-1              C  =  A diag B'    (KC,KC) with a k-scaling vector
// Block-level triangular structure lives in the task's k-range; diagonal blocks of triangular
// operands hold explicit zeros, so the kernel itself is a plain dense tile GEMM.
//
// A CTA computes a TM x 128 slice of a task (TM = 128 or 64) with WM x WN warps; each warp owns
// MF x NF m8n8k4 fragments.  Operands are staged with 16-byte cp.async into a ring of padded
// shared-memory tiles whose leading dimensions (BK + 4, rows + 4 doubles) make every half-warp
// fragment load hit 16 distinct 8-byte bank pairs.  The tile policy is a template parameter
// (GemmCfg); the policies that were measured against each other are kept selectable
// (gps_dbg_set_variant) and the numbers are in profiles/.
//
// tcgen05 has no f64 kind, so on B200 the FP64 tensor path IS mma.sync DMMA.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <map>
#include <mutex>
#include <tuple>

#include "gps_common.cuh"

namespace {

// Tile policy.  A task is always a 128 x 128 output tile; a CTA computes a TM x 128 slice of it
// (TM = 128: one CTA per task, TM = 64: two).  BK = k-depth of a pipeline stage, STAGES = ring
// depth, WM x WN = warp grid, PIPE = explicit register double-buffering of the fragments,
// MINB = CTAs per SM the register budget is sized for.
template <int TM_, int BK_, int STAGES_, int WM_, int WN_, bool PIPE_, int MINB_>
struct GemmCfg {
  static constexpr int TM = TM_;                                 // x 128 columns
  static constexpr int BK = BK_;
  static constexpr int STAGES = STAGES_;
  static constexpr bool PIPE = PIPE_;
  static constexpr int WN = WN_, MINB = MINB_;
  static constexpr int THREADS = 32 * WM_ * WN_;
  static constexpr int MF = TM_ / WM_ / 8, NF = 128 / WN_ / 8;   // m8n8k4 fragments per warp tile
  static constexpr int SPLIT = 128 / TM_;
  static constexpr int LD_KC = BK_ + 4;                          // [rows][BK+4]: (BK+4) % 16 == 4
  static constexpr int LD_MC_A = TM_ + 4, LD_MC_B = 128 + 4;     // [BK][rows+4]: (rows+4) % 16 == 4
  static constexpr int OPER_A = (TM_ * LD_KC > BK_ * LD_MC_A) ? TM_ * LD_KC : BK_ * LD_MC_A;
  static constexpr int OPER_B = (128 * LD_KC > BK_ * LD_MC_B) ? 128 * LD_KC : BK_ * LD_MC_B;
  static constexpr int STAGE = OPER_A + OPER_B + BK_;            // doubles per stage (+ the k-scaling slice)
  static constexpr size_t SMEM = (size_t)STAGES_ * STAGE * sizeof(double);
};

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// stage one ROWS x BK operand panel.  KC: rows are matrix rows, k contiguous in memory.
// MC: the panel is read from a [k][rows] matrix (rows contiguous in memory).
template <class Cfg, int ROWS, bool MC>
__device__ __forceinline__ void load_panel(double* s, const double* __restrict__ g, int64_t ld, int row0,
                                           int k, int tid) {
  constexpr int CHUNKS = ROWS * Cfg::BK / 2;     // 16-byte chunks per panel
  constexpr int LDM = ROWS + 4;
#pragma unroll
  for (int i = 0; i < CHUNKS / Cfg::THREADS; ++i) {
    const int c = tid + i * Cfg::THREADS;
    if (!MC) {
      constexpr int CPR = Cfg::BK / 2;          // chunks per row
      const int r = c / CPR, kc = c % CPR;
      cp_async16(s + r * Cfg::LD_KC + kc * 2, g + (int64_t)(row0 + r) * ld + k + kc * 2);
    } else {
      constexpr int CPK = ROWS / 2;             // chunks per k-row
      const int kr = c / CPK, mc = c % CPK;
      cp_async16(s + kr * LDM + mc * 2, g + (int64_t)(k + kr) * ld + row0 + mc * 2);
    }
  }
}

template <class Cfg, int ROWS, bool MC>
__device__ __forceinline__ double frag(const double* s, int row, int k) {
  return MC ? s[k * (ROWS + 4) + row] : s[row * Cfg::LD_KC + k];
}

template <class Cfg, bool A_MC, bool B_MC, bool DVEC, bool MIRROR>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
gemm_tile_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                 double* __restrict__ C, int64_t ldc, double alpha, double beta,
                 const double* __restrict__ dvec, const GemmTask* __restrict__ tasks) {
  extern __shared__ __align__(16) double smem[];
  constexpr int BK = Cfg::BK, STAGES = Cfg::STAGES, KSTEPS = Cfg::BK / 4, TM = Cfg::TM;
  constexpr int MF = Cfg::MF, NF = Cfg::NF, WROWS = 8 * Cfg::MF, WCOLS = 8 * Cfg::NF;
  constexpr int STAGE = Cfg::STAGE, OPER_A = Cfg::OPER_A;
  GemmTask t = tasks[blockIdx.x / Cfg::SPLIT];
  const bool diag_tile = t.c_row == t.c_col;
  const int half = (int)(blockIdx.x % Cfg::SPLIT) * TM;
  t.a_row += half;
  t.c_row += half;
  if (t.flags & GEMM_TRI_END) t.k1 -= 128 - half - TM;
  if (t.flags & GEMM_TRI_BEGIN) t.k0 += half;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
  const int g = lane >> 2, tq = lane & 3;
  const int nk = (t.k1 - t.k0) / BK;

  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      load_panel<Cfg, TM, A_MC>(smem + (size_t)s * STAGE, A, lda, t.a_row, t.k0 + s * BK, tid);
      load_panel<Cfg, 128, B_MC>(smem + (size_t)s * STAGE + OPER_A, B, ldb, t.b_row, t.k0 + s * BK, tid);
      if (DVEC && tid < BK / 2)   // the stage's slice of the k-scaling vector rides in the same cp.async group
        cp_async16(smem + (size_t)s * STAGE + OPER_A + Cfg::OPER_B + tid * 2, dvec + t.k0 + s * BK + tid * 2);
    }
    cp_async_commit();
  }

  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = kb + STAGES - 1;
      if (nx < nk) {
        double* sl = smem + (size_t)(nx % STAGES) * STAGE;
        load_panel<Cfg, TM, A_MC>(sl, A, lda, t.a_row, t.k0 + nx * BK, tid);
        load_panel<Cfg, 128, B_MC>(sl + OPER_A, B, ldb, t.b_row, t.k0 + nx * BK, tid);
        if (DVEC && tid < BK / 2) cp_async16(sl + OPER_A + Cfg::OPER_B + tid * 2, dvec + t.k0 + nx * BK + tid * 2);
      }
      cp_async_commit();
    }
    const double* As = smem + (size_t)(kb % STAGES) * STAGE;
    const double* Bs = As + OPER_A;
    const double* Ds = Bs + Cfg::OPER_B;
    if (Cfg::PIPE) {
      double a[2][MF], b[2][NF];
#pragma unroll
      for (int i = 0; i < MF; ++i) a[0][i] = frag<Cfg, TM, A_MC>(As, wm * WROWS + i * 8 + g, tq);
#pragma unroll
      for (int j = 0; j < NF; ++j) b[0][j] = frag<Cfg, 128, B_MC>(Bs, wn * WCOLS + j * 8 + g, tq);
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const int cur = kk & 1, nxt = cur ^ 1;
        if (kk + 1 < KSTEPS) {
          const int k = (kk + 1) * 4 + tq;
#pragma unroll
          for (int j = 0; j < NF; ++j) b[nxt][j] = frag<Cfg, 128, B_MC>(Bs, wn * WCOLS + j * 8 + g, k);
#pragma unroll
          for (int i = 0; i < MF; ++i) a[nxt][i] = frag<Cfg, TM, A_MC>(As, wm * WROWS + i * 8 + g, k);
        }
        if (DVEC) {
          const double dv = Ds[kk * 4 + tq];
#pragma unroll
          for (int i = 0; i < MF; ++i) a[cur][i] *= dv;
        }
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const int k = kk * 4 + tq;
        double a[MF], b[NF];
#pragma unroll
        for (int i = 0; i < MF; ++i) a[i] = frag<Cfg, TM, A_MC>(As, wm * WROWS + i * 8 + g, k);
#pragma unroll
        for (int j = 0; j < NF; ++j) b[j] = frag<Cfg, 128, B_MC>(Bs, wn * WCOLS + j * 8 + g, k);
        if (DVEC) {
          const double dv = Ds[k];
          if (MF <= NF) {
#pragma unroll
            for (int i = 0; i < MF; ++i) a[i] *= dv;
          } else {
#pragma unroll
            for (int j = 0; j < NF; ++j) b[j] *= dv;
          }
        }
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
  }
  cp_async_wait<0>();
  static_assert(!MIRROR || (size_t)128 * (TM + 1) * sizeof(double) <= Cfg::SMEM, "transpose buffer must fit in the operand ring");
  if (MIRROR) __syncthreads();   // the operand ring is reused as the transpose buffer below

  // epilogue: thread holds C[row][col..col+1] of each fragment
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const int row = t.c_row + wm * WROWS + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int col = t.c_col + wn * WCOLS + j * 8 + 2 * tq;
      double2* p = reinterpret_cast<double2*>(C + (int64_t)row * ldc + col);
      double2 v;
      v.x = alpha * acc[i][j][0];
      v.y = alpha * acc[i][j][1];
      if (beta != 0.0) {
        const double2 o = *p;
        v.x += beta * o.x;
        v.y += beta * o.y;
      }
      *p = v;
      if (MIRROR && !diag_tile) {
        // stage the transposed slice in the (now idle) operand ring: Ts[col][row], leading dimension TM + 1
        const int lr = wm * WROWS + i * 8 + g, lc = wn * WCOLS + j * 8 + 2 * tq;
        smem[lc * (TM + 1) + lr] = v.x;
        smem[(lc + 1) * (TM + 1) + lr] = v.y;
      }
    }
  }
  if (MIRROR && !diag_tile) {
    // mirrored tile C[c_col + n][c_row + m]: each warp writes rows of TM contiguous doubles (coalesced) instead
    // of every thread scattering 8-byte stores with stride ldc
    __syncthreads();
    for (int n = warp; n < 128; n += Cfg::THREADS / 32) {
      double* dst = C + (int64_t)(t.c_col + n) * ldc + t.c_row;
      for (int m = lane; m < TM; m += 32) dst[m] = smem[n * (TM + 1) + m];
    }
  }
}

// =====================================================================================================================
// TMA-fed variant of the (KC, KC) tile GEMM (A/B policy 9; round-1 verdict item 4c).
// Same CTA tile (64 x 128 slice of a 128 x 128 task, 4 warps, BK = 16), same k-order per output element -> bit-identical
// results; only the operand movement differs: one elected thread issues two cp.async.bulk.tensor.2d per k-block
// (A: 64 x 16, B: 128 x 16 doubles) that complete on an mbarrier, instead of 12 16-byte cp.async per thread.  Tiles
// land dense (128-byte rows) with the hardware 128-byte swizzle; fragment loads apply the same XOR (16-byte chunk
// index ^ (row & 7)).  For 8-byte elements that pattern is 2-way bank conflicted across the half-warp's four rows
// pairs (the padded cp.async layout is conflict-free) — the kernel is DMMA-bound, so this does not show.
// =====================================================================================================================
constexpr int TMA_STAGES = 4;
constexpr int TMA_A_BYTES = 64 * 16 * 8, TMA_B_BYTES = 128 * 16 * 8, TMA_STAGE_BYTES = TMA_A_BYTES + TMA_B_BYTES;
constexpr size_t TMA_SMEM = (size_t)TMA_STAGES * TMA_STAGE_BYTES + 1024;   // + slack for the 1024-byte alignment

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, int parity) {
  // no PTX labels: the function is inlined at several sites of one kernel
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// dense [rows][16 doubles] tile with the 128-byte swizzle: address of element (row, k)
__device__ __forceinline__ double swz_ld(const unsigned char* tile, int row, int k) {
  const int off = row * 128 + ((((k >> 1) ^ (row & 7)) << 4) | ((k & 1) << 3));
  return *reinterpret_cast<const double*>(tile + off);
}

// operand tile (row `row` of the panel, contraction index k) in either layout
template <int ROWS, bool MC>
__device__ __forceinline__ double tma_frag(const unsigned char* tile, int row, int k) {
  if (MC) return reinterpret_cast<const double*>(tile)[k * ROWS + row];   // dense [16][ROWS], no swizzle
  return swz_ld(tile, row, k);                                            // dense [ROWS][16], 128-byte swizzle
}

template <bool A_MC, bool B_MC, bool MIRROR>
__global__ void __launch_bounds__(128, 2)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C,
                int64_t ldc, double alpha, double beta, const GemmTask* __restrict__ tasks, int mode) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ unsigned long long full[TMA_STAGES], empty[TMA_STAGES];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int TM = 64, MF = 4, NF = 8, WROWS = 32, WCOLS = 64, BK = 16, S = TMA_STAGES;   // 2 x 2 warps, 32 x 64 warp tiles
  GemmTask t = tasks[blockIdx.x >> 1];
  const bool diag_tile = t.c_row == t.c_col;
  const int half = (int)(blockIdx.x & 1) * TM;
  t.a_row += half;
  t.c_row += half;
  if (t.flags & GEMM_TRI_END) t.k1 -= 128 - half - TM;
  if (t.flags & GEMM_TRI_BEGIN) t.k0 += half;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, tq = lane & 3;
  const int nk = (t.k1 - t.k0) / BK;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], mode == 2 ? 128 : 4);   // arrivals per phase: one per warp (mode 1) or per thread (mode 2)
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int kb) {
    unsigned char* sl = smem + (size_t)(kb % S) * TMA_STAGE_BYTES;
    const int k = t.k0 + kb * BK;
    mbar_expect_tx(&full[kb % S], TMA_STAGE_BYTES);
    if (A_MC) tma_load_2d(sl, &tmA, &full[kb % S], t.a_row, k);
    else tma_load_2d(sl, &tmA, &full[kb % S], k, t.a_row);
    if (B_MC) tma_load_2d(sl + TMA_A_BYTES, &tmB, &full[kb % S], t.b_row, k);
    else tma_load_2d(sl + TMA_A_BYTES, &tmB, &full[kb % S], k, t.b_row);
  };
  if (tid == 0)
    for (int s = 0; s < S - 1 && s < nk; ++s) issue(s);
  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int kb = 0; kb < nk; ++kb) {
    // No CTA-wide barrier in the main loop: the producer thread alone waits until all four warps have released the
    // slot it is about to refill (the one read in iteration kb - 1), the other warps run ahead up to S - 1 stages.
    if (mode == 0) {
      mbar_wait(&full[kb % S], (kb / S) & 1);
      __syncthreads();                     // everyone is done with the slot the next load overwrites
      if (tid == 0 && kb + S - 1 < nk) issue(kb + S - 1);
    } else {
      if (tid == 0 && kb + S - 1 < nk) {
        if (kb >= 1) mbar_wait(&empty[(kb - 1) % S], ((kb - 1) / S) & 1);
        issue(kb + S - 1);
      }
      // the producer lane spins alone inside the branch above: bring warp 0 back together before the warp-wide
      // mma.sync below (with independent thread scheduling the other 31 lanes are otherwise free to run ahead of
      // it — a first build without this line produced wrong tiles now and then)
      __syncwarp();
      mbar_wait(&full[kb % S], (kb / S) & 1);
    }
    const unsigned char* As = smem + (size_t)(kb % S) * TMA_STAGE_BYTES;
    const unsigned char* Bs = As + TMA_A_BYTES;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      const int k = kk * 4 + tq;
      double a[MF], b[NF];
#pragma unroll
      for (int i = 0; i < MF; ++i) a[i] = tma_frag<TM, A_MC>(As, wm * WROWS + i * 8 + g, k);
#pragma unroll
      for (int j = 0; j < NF; ++j) b[j] = tma_frag<128, B_MC>(Bs, wn * WCOLS + j * 8 + g, k);
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (mode == 1) {
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[kb % S])) : "memory");
    } else if (mode == 2) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[kb % S])) : "memory");
    }
  }
  double* tbuf = reinterpret_cast<double*>(smem);
  static_assert(!MIRROR || (size_t)128 * (TM + 1) * sizeof(double) <= (size_t)TMA_STAGES * TMA_STAGE_BYTES, "transpose buffer must fit");
  if (MIRROR) __syncthreads();             // the operand ring is reused as the transpose buffer below
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const int row = t.c_row + wm * WROWS + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int col = t.c_col + wn * WCOLS + j * 8 + 2 * tq;
      double2* p = reinterpret_cast<double2*>(C + (int64_t)row * ldc + col);
      double2 v;
      v.x = alpha * acc[i][j][0];
      v.y = alpha * acc[i][j][1];
      if (beta != 0.0) {
        const double2 o = *p;
        v.x += beta * o.x;
        v.y += beta * o.y;
      }
      *p = v;
      if (MIRROR && !diag_tile) {
        const int lr = wm * WROWS + i * 8 + g, lc = wn * WCOLS + j * 8 + 2 * tq;
        tbuf[lc * (TM + 1) + lr] = v.x;
        tbuf[(lc + 1) * (TM + 1) + lr] = v.y;
      }
    }
  }
  if (MIRROR && !diag_tile) {
    __syncthreads();
    for (int n = warp; n < 128; n += 4) {
      double* dst = C + (int64_t)(t.c_col + n) * ldc + t.c_row;
      for (int m = lane; m < TM; m += 32) dst[m] = tbuf[n * (TM + 1) + m];
    }
  }
}

// tensor maps keyed by (base, leading dimension, box), created on first use (driver entry point: no -lcuda)
struct TmaMaps {
  PFN_cuTensorMapEncodeTiled encode = nullptr;
  std::map<std::tuple<const void*, int64_t, int, int>, CUtensorMap> cache;
  std::mutex mu;   // contexts may be driven from different host threads; the cache is process-wide
};
TmaMaps g_tma;

// panel_rows x 16 operand tile of a row-major matrix with leading dimension ld.  KC (k contiguous): box {16, rows},
// 128-byte swizzle; MC (panel index contiguous): box {rows, 16}, dense.  The row extent is left unbounded (the map is
// only used with in-range coordinates; the buffer's true height is not known here).
int tma_map(gps_ctx* ctx, const double* base, int64_t ld, int panel_rows, bool mc, CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(g_tma.mu);
  if (!g_tma.encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return gps_fail(ctx, GPS_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    g_tma.encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  auto key = std::make_tuple((const void*)base, ld, panel_rows, mc ? 1 : 0);
  auto it = g_tma.cache.find(key);
  if (it == g_tma.cache.end()) {
    if (g_tma.cache.size() > 512) g_tma.cache.clear();
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)1 << 30};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {mc ? (cuuint32_t)panel_rows : 16u, mc ? 16u : (cuuint32_t)panel_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = g_tma.encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, mc ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gps_fail(ctx, GPS_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    it = g_tma.cache.emplace(key, m).first;
  }
  *out = it->second;
  return GPS_OK;
}

// Persistent variant of gemm_tma_kernel for launches that share the GPU with a latency-critical lane (the diagonal-block
// chain of POTRF): the grid is capped below the number of resident CTA slots, every CTA takes 64 x 128 slices from a
// ticket counter until the list is drained, so slots stay free for the chain's kernels at all times (a CTA of a deep
// merge holds its slot for up to 0.7 ms; without free slots a chain launch of a few dozen CTAs waits for that many
// to retire).  Same tile, same k-order per output element: bit-identical results.
template <bool A_MC, bool B_MC>
__global__ void __launch_bounds__(128, 2)
gemm_tma_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C,
                        int64_t ldc, double alpha, double beta, const GemmTask* __restrict__ tasks, int nslices,
                        int* __restrict__ ticket) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ unsigned long long full[TMA_STAGES];
  __shared__ int next_slice;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int TM = 64, MF = 4, NF = 8, WROWS = 32, WCOLS = 64, BK = 16, S = TMA_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, tq = lane & 3;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  unsigned base = 0;                       // k-blocks this CTA has consumed so far: position and phase of the ring
  int slice = blockIdx.x;
  while (slice < nslices) {
    GemmTask t = tasks[slice >> 1];
    const int half = (slice & 1) * TM;
    t.a_row += half;
    t.c_row += half;
    if (t.flags & GEMM_TRI_END) t.k1 -= 128 - half - TM;
    if (t.flags & GEMM_TRI_BEGIN) t.k0 += half;
    const int nk = (t.k1 - t.k0) / BK;
    auto issue = [&](int kb) {
      const unsigned q = base + kb;
      unsigned char* sl = smem + (size_t)(q % S) * TMA_STAGE_BYTES;
      const int k = t.k0 + kb * BK;
      mbar_expect_tx(&full[q % S], TMA_STAGE_BYTES);
      if (A_MC) tma_load_2d(sl, &tmA, &full[q % S], t.a_row, k);
      else tma_load_2d(sl, &tmA, &full[q % S], k, t.a_row);
      if (B_MC) tma_load_2d(sl + TMA_A_BYTES, &tmB, &full[q % S], t.b_row, k);
      else tma_load_2d(sl + TMA_A_BYTES, &tmB, &full[q % S], k, t.b_row);
    };
    if (tid == 0)
      for (int s = 0; s < S - 1 && s < nk; ++s) issue(s);
    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kb = 0; kb < nk; ++kb) {
      const unsigned q = base + kb;
      mbar_wait(&full[q % S], (q / S) & 1);
      __syncthreads();                     // everyone is done with the slot the next load overwrites
      if (tid == 0 && kb + S - 1 < nk) issue(kb + S - 1);
      const unsigned char* As = smem + (size_t)(q % S) * TMA_STAGE_BYTES;
      const unsigned char* Bs = As + TMA_A_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        const int k = kk * 4 + tq;
        double a[MF], b[NF];
#pragma unroll
        for (int i = 0; i < MF; ++i) a[i] = tma_frag<TM, A_MC>(As, wm * WROWS + i * 8 + g, k);
#pragma unroll
        for (int j = 0; j < NF; ++j) b[j] = tma_frag<128, B_MC>(Bs, wn * WCOLS + j * 8 + g, k);
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      const int row = t.c_row + wm * WROWS + i * 8 + g;
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int col = t.c_col + wn * WCOLS + j * 8 + 2 * tq;
        double2* p = reinterpret_cast<double2*>(C + (int64_t)row * ldc + col);
        double2 v;
        v.x = alpha * acc[i][j][0];
        v.y = alpha * acc[i][j][1];
        if (beta != 0.0) {
          const double2 o = *p;
          v.x += beta * o.x;
          v.y += beta * o.y;
        }
        *p = v;
      }
    }
    base += (unsigned)nk;
    if (tid == 0) next_slice = (int)gridDim.x + atomicAdd(ticket, 1);
    __syncthreads();                       // the ring is free for the next slice's first loads; next_slice is visible
    slice = next_slice;
  }
}

template <bool A_MC, bool B_MC>
int launch_tma_persist(gps_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                       double alpha, double beta, const GemmTask* tasks, size_t ntasks) {
  CUtensorMap ma, mb;
  GPS_CHECK(tma_map(ctx, A, lda, 64, A_MC, &ma));
  GPS_CHECK(tma_map(ctx, B, ldb, 128, B_MC, &mb));
  auto kern = gemm_tma_persist_kernel<A_MC, B_MC>;
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
    configured = true;
  }
  if (!ctx->d_tickets) GPS_CUDA(cudaMalloc(&ctx->d_tickets, 256 * sizeof(int)));
  int* ticket = ctx->d_tickets + (ctx->ticket_seq++ & 255u);
  GPS_CUDA(cudaMemsetAsync(ticket, 0, sizeof(int), ctx->stream));
  kern<<<(unsigned)ctx->gemm_grid_cap, 128, TMA_SMEM, ctx->stream>>>(ma, mb, C, ldc, alpha, beta, tasks, (int)(ntasks * 2), ticket);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <bool A_MC, bool B_MC, bool MIRROR>
int launch_tma_k(gps_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, double alpha,
                 double beta, const GemmTask* tasks, size_t ntasks) {
  CUtensorMap ma, mb;
  GPS_CHECK(tma_map(ctx, A, lda, 64, A_MC, &ma));
  GPS_CHECK(tma_map(ctx, B, ldb, 128, B_MC, &mb));
  auto kern = gemm_tma_kernel<A_MC, B_MC, MIRROR>;
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
    configured = true;
  }
  kern<<<(unsigned)(ntasks * 2), 128, TMA_SMEM, ctx->stream>>>(ma, mb, C, ldc, alpha, beta, tasks, ctx->gemm_variant - 9);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

int launch_tma(gps_ctx* ctx, int kind, bool mirror, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
               int64_t ldc, double alpha, double beta, const GemmTask* tasks, size_t ntasks) {
  if (ctx->gemm_grid_cap > 0 && ntasks * 2 > (size_t)ctx->gemm_grid_cap && ctx->gemm_variant == 9) {
    if (kind == GEMM_KC_KC) return launch_tma_persist<false, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, tasks, ntasks);
    if (kind == GEMM_KC_MC) return launch_tma_persist<false, true>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, tasks, ntasks);
  }
  if (kind == GEMM_KC_KC) return launch_tma_k<false, false, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, tasks, ntasks);
  if (kind == GEMM_KC_MC) return launch_tma_k<false, true, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, tasks, ntasks);
  if (mirror) return launch_tma_k<true, true, true>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, tasks, ntasks);
  return launch_tma_k<true, true, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, tasks, ntasks);
}

template <class Cfg, bool A_MC, bool B_MC, bool DVEC, bool MIRROR>
int launch(gps_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
           double alpha, double beta, const double* dvec, const GemmTask* tasks, size_t ntasks) {
  auto kern = gemm_tile_kernel<Cfg, A_MC, B_MC, DVEC, MIRROR>;
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    configured = true;
  }
  kern<<<(unsigned)(ntasks * Cfg::SPLIT), Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(A, lda, B, ldb, C, ldc, alpha, beta, dvec,
                                                                  tasks);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

template <class Cfg>
int dispatch(gps_ctx* ctx, int kind, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
             int64_t ldc, double alpha, double beta, const double* dvec, bool mirror, const GemmTask* d_tasks,
             size_t ntasks) {
  if (kind == GEMM_KC_KC) {
    if (dvec) return launch<Cfg, false, false, true, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, dvec, d_tasks, ntasks);
    return launch<Cfg, false, false, false, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, dvec, d_tasks, ntasks);
  }
  if (kind == GEMM_KC_MC)
    return launch<Cfg, false, true, false, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, dvec, d_tasks, ntasks);
  if (kind == GEMM_MC_MC) {
    if (mirror) return launch<Cfg, true, true, false, true>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, dvec, d_tasks, ntasks);
    return launch<Cfg, true, true, false, false>(ctx, A, lda, B, ldb, C, ldc, alpha, beta, dvec, d_tasks, ntasks);
  }
  return gps_fail(ctx, GPS_EINVAL, "gemm kind %d", kind);
}

}  // namespace

int gps_gemm_tasks(gps_ctx* ctx, int kind, const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* C, int64_t ldc, double alpha, double beta, const double* dvec, bool mirror,
                   const GemmTask* d_tasks, size_t ntasks) {
  if (ntasks == 0) return GPS_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->time_gemm) {
    if (ctx->gemm_events_used == ctx->gemm_events.size()) {
      cudaEvent_t a, b;
      GPS_CUDA(cudaEventCreate(&a));
      GPS_CUDA(cudaEventCreate(&b));
      ctx->gemm_events.push_back({a, b});
    }
    e0 = ctx->gemm_events[ctx->gemm_events_used].first;
    e1 = ctx->gemm_events[ctx->gemm_events_used].second;
    ctx->gemm_events_used++;
    GPS_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  int r;
#define GPS_GEMM_ARGS ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks
  // strip policies for the few-tile, k = 128 launches on POTRF's serial chain: a 128 x 128 task is split over
  // 4 (or 8) CTAs of 32 (16) rows, so the eight dependent k-steps each carry a quarter (eighth) of the DMMA work
  // ... and every other launch too small to fill the SMs with the shipped policy (deep TRTRI merges, the last
  // POTRF block columns, M x M products of the FITC matrix form, mid-size grid sweeps): 8 CTAs per task up to
  // sm_count / 8 tasks, 4 CTAs per task up to sm_count / 4.  The k-order per output element is the same in
  // every policy, so results do not depend on the choice.
  int strip = ctx->gemm_strip_policy;
  if (strip == 0 && ctx->gemm_auto_strip) {
    if (ntasks * 8 <= (size_t)ctx->sm_count) strip = 16;
    else if (ntasks * 4 <= (size_t)ctx->sm_count) strip = 32;
  }
  if (ctx->gemm_variant >= 9 && ctx->gemm_variant <= 11 && strip == 0 && !dvec && (kind == GEMM_KC_KC || kind == GEMM_KC_MC || kind == GEMM_MC_MC)) {
    r = launch_tma(ctx, kind, mirror, A, lda, B, ldb, C, ldc, alpha, beta, d_tasks, ntasks);   // TMA-fed operand ring
  } else if (strip == 32) {
    r = dispatch<GemmCfg<32, 16, 3, 1, 4, false, 2>>(GPS_GEMM_ARGS);
  } else if (strip == 16) {
    r = dispatch<GemmCfg<16, 16, 3, 1, 4, false, 2>>(GPS_GEMM_ARGS);
  } else
  switch (ctx->gemm_variant) {            //            TM  BK  ST WM WN PIPE  MINB
    case 0: r = dispatch<GemmCfg<128, 16, 4, 2, 4, false, 1>>(GPS_GEMM_ARGS); break;
    case 2: r = dispatch<GemmCfg<128, 32, 3, 2, 4, false, 1>>(GPS_GEMM_ARGS); break;
    case 4: r = dispatch<GemmCfg<128, 16, 4, 4, 4, false, 1>>(GPS_GEMM_ARGS); break;
    case 7: r = dispatch<GemmCfg<64, 16, 3, 2, 2, true, 2>>(GPS_GEMM_ARGS); break;
    case 8: r = dispatch<GemmCfg<64, 16, 3, 1, 4, false, 2>>(GPS_GEMM_ARGS); break;   // 64 x 32 warp tiles
    case 5: r = dispatch<GemmCfg<128, 32, 3, 4, 4, false, 1>>(GPS_GEMM_ARGS); break;
    default: r = dispatch<GemmCfg<64, 16, 3, 2, 2, false, 2>>(GPS_GEMM_ARGS); break;   // 6: two 4-warp CTAs per SM
  }
#undef GPS_GEMM_ARGS
  if (r != GPS_OK) return r;
  ctx->launches++;
  if (ctx->time_gemm) GPS_CUDA(cudaEventRecord(e1, ctx->stream));
  return GPS_OK;
}
