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
// Tile: 128 x 128 x 16, 256 threads = 8 warps as 2 (m) x 4 (n), warp tile 64 x 32 =
// 8 x 4 m8n8k4 fragments (64 fp64 accumulators per thread).  Operands are staged with
// 16-byte cp.async into a 4-deep ring of padded shared-memory tiles whose leading dimensions
// (20 / 132 doubles) make every half-warp fragment load hit 16 distinct 8-byte bank pairs.
//
// tcgen05 has no f64 kind, so on B200 the FP64 tensor path IS mma.sync DMMA.
#include "gps_common.cuh"

namespace {

constexpr int LD_MC = 132;    // [BK][132]

// Tile-shape policy: BK = k-depth of one pipeline stage, STAGES = ring depth, PIPE = explicit
// register double-buffering of the m8n8k4 fragments across k4-steps.
template <int BK_, int STAGES_, bool PIPE_, int WM_ = 2>
struct GemmCfg {
  static constexpr int BK = BK_;
  static constexpr int STAGES = STAGES_;
  static constexpr bool PIPE = PIPE_;
  static constexpr int WM = WM_, WN = 4;                   // warp grid: WM x 4 warps
  static constexpr int THREADS = 32 * WM_ * 4;
  static constexpr int MF = 128 / WM_ / 8, NF = 4;         // m8n8k4 fragments per warp tile
  static constexpr int LD_KC = BK_ + 4;                    // [128][BK+4]: (BK+4) % 16 == 4
  static constexpr int OPER = 128 * LD_KC;                 // doubles per operand per stage (>= BK*132)
  static constexpr size_t SMEM = (size_t)STAGES_ * 2 * OPER * sizeof(double);
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

// stage one 128 x BK operand panel. KC: rows are matrix rows, k contiguous in memory.
// MC: the panel is read from a [k][m] matrix (m contiguous in memory).
template <class Cfg, bool MC>
__device__ __forceinline__ void load_panel(double* s, const double* __restrict__ g, int64_t ld, int row0,
                                           int k, int tid) {
  constexpr int CHUNKS = 128 * Cfg::BK / 2;     // 16-byte chunks per panel
#pragma unroll
  for (int i = 0; i < CHUNKS / Cfg::THREADS; ++i) {
    const int c = tid + i * Cfg::THREADS;
    if (!MC) {
      constexpr int CPR = Cfg::BK / 2;          // chunks per row
      const int r = c / CPR, kc = c % CPR;
      cp_async16(s + r * Cfg::LD_KC + kc * 2, g + (int64_t)(row0 + r) * ld + k + kc * 2);
    } else {
      const int kr = c >> 6, mc = c & 63;
      cp_async16(s + kr * LD_MC + mc * 2, g + (int64_t)(k + kr) * ld + row0 + mc * 2);
    }
  }
}

template <class Cfg, bool MC>
__device__ __forceinline__ double frag(const double* s, int row, int k) {
  return MC ? s[k * LD_MC + row] : s[row * Cfg::LD_KC + k];
}

template <class Cfg, bool A_MC, bool B_MC, bool DVEC, bool MIRROR>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_tile_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                 double* __restrict__ C, int64_t ldc, double alpha, double beta,
                 const double* __restrict__ dvec, const GemmTask* __restrict__ tasks) {
  extern __shared__ __align__(16) double smem[];
  constexpr int BK = Cfg::BK, STAGES = Cfg::STAGES, OPER = Cfg::OPER, KSTEPS = Cfg::BK / 4;
  constexpr int MF = Cfg::MF, NF = Cfg::NF, WROWS = 8 * Cfg::MF;
  const GemmTask t = tasks[blockIdx.x];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;   // WM x 4 warps
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
      load_panel<Cfg, A_MC>(smem + (size_t)s * 2 * OPER, A, lda, t.a_row, t.k0 + s * BK, tid);
      load_panel<Cfg, B_MC>(smem + (size_t)s * 2 * OPER + OPER, B, ldb, t.b_row, t.k0 + s * BK, tid);
    }
    cp_async_commit();
  }

  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = kb + STAGES - 1;
      if (nx < nk) {
        double* sl = smem + (size_t)(nx % STAGES) * 2 * OPER;
        load_panel<Cfg, A_MC>(sl, A, lda, t.a_row, t.k0 + nx * BK, tid);
        load_panel<Cfg, B_MC>(sl + OPER, B, ldb, t.b_row, t.k0 + nx * BK, tid);
      }
      cp_async_commit();
    }
    const double* As = smem + (size_t)(kb % STAGES) * 2 * OPER;
    const double* Bs = As + OPER;
    if (Cfg::PIPE) {
      double a[2][MF], b[2][NF];
#pragma unroll
      for (int i = 0; i < MF; ++i) a[0][i] = frag<Cfg, A_MC>(As, wm * WROWS + i * 8 + g, tq);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[0][j] = frag<Cfg, B_MC>(Bs, wn * 32 + j * 8 + g, tq);
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const int cur = kk & 1, nxt = cur ^ 1;
        if (kk + 1 < KSTEPS) {
          const int k = (kk + 1) * 4 + tq;
#pragma unroll
          for (int j = 0; j < 4; ++j) b[nxt][j] = frag<Cfg, B_MC>(Bs, wn * 32 + j * 8 + g, k);
#pragma unroll
          for (int i = 0; i < MF; ++i) a[nxt][i] = frag<Cfg, A_MC>(As, wm * WROWS + i * 8 + g, k);
        }
        if (DVEC) {
          const double dv = __ldg(dvec + t.k0 + kb * BK + kk * 4 + tq);
#pragma unroll
          for (int j = 0; j < 4; ++j) b[cur][j] *= dv;
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
        for (int i = 0; i < MF; ++i) a[i] = frag<Cfg, A_MC>(As, wm * WROWS + i * 8 + g, k);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = frag<Cfg, B_MC>(Bs, wn * 32 + j * 8 + g, k);
        if (DVEC) {
          const double dv = __ldg(dvec + t.k0 + kb * BK + k);
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] *= dv;
        }
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: thread holds C[row][col..col+1] of each fragment
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const int row = t.c_row + wm * WROWS + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int col = t.c_col + wn * 32 + j * 8 + 2 * tq;
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
      if (MIRROR && t.c_row != t.c_col) {
        C[(int64_t)col * ldc + row] = v.x;
        C[(int64_t)(col + 1) * ldc + row] = v.y;
      }
    }
  }
}

template <class Cfg, bool A_MC, bool B_MC, bool DVEC, bool MIRROR>
int launch(gps_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
           double alpha, double beta, const double* dvec, const GemmTask* tasks, size_t ntasks) {
  auto kern = gemm_tile_kernel<Cfg, A_MC, B_MC, DVEC, MIRROR>;
  static bool configured = false;
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    configured = true;
  }
  kern<<<(unsigned)ntasks, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(A, lda, B, ldb, C, ldc, alpha, beta, dvec,
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
  switch (ctx->gemm_variant) {
    case 1: r = dispatch<GemmCfg<16, 4, true>>(ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks); break;
    case 2: r = dispatch<GemmCfg<32, 3, false>>(ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks); break;
    case 4: r = dispatch<GemmCfg<16, 4, false, 4>>(ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks); break;
    case 5: r = dispatch<GemmCfg<32, 3, false, 4>>(ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks); break;
    case 3: r = dispatch<GemmCfg<32, 3, true>>(ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks); break;
    default: r = dispatch<GemmCfg<16, 4, false>>(ctx, kind, A, lda, B, ldb, C, ldc, alpha, beta, dvec, mirror, d_tasks, ntasks); break;
  }
  if (r != GPS_OK) return r;
  ctx->launches++;
  if (ctx->time_gemm) GPS_CUDA(cudaEventRecord(e1, ctx->stream));
  return GPS_OK;
}
