// ARD squared-exponential Gram construction (replaces ARD, KF:7-23, and the `+ sn2 * eye` of
// KF:241) — one fused pass: scaled-difference distance, exp, signal variance and the noise
// diagonal are all applied in the tile that is written, nothing but K itself touches HBM.
//
// The reference forms the distance through the norm expansion 2xx' - |x|^2 - |x'|^2 (KF:15-20);
// here it is the direct sum of squared scaled differences, which is the same number to rounding
// and exactly e^a on the diagonal.
//
// par (device): [0] = e^a, [1] = e^c (noise variance), [2 + d] = 1 / l_d.
#include "gps_common.cuh"
#include "gps_exp.cuh"

namespace {

constexpr int TS = 128;       // output tile
constexpr int MAXD = 64;

// Symmetric training Gram into the tile-padded [Np, Np] buffer: lower tiles only (diagonal
// tiles complete).  Rows/cols >= N are the identity so that the padded matrix stays SPD and
// the padding never mixes with the data.  256 threads; thread (ty, tx) writes rows
// ty + 16 r and, per row, four coalesced 16-byte pairs at columns 32 c + 2 tx.
__global__ void __launch_bounds__(256, 2)
gram_sym_kernel(const double* __restrict__ X, int64_t N, int64_t Np, int D, const double* __restrict__ par,
                double* __restrict__ K) {
  extern __shared__ double sh[];
  double* xi = sh;                 // [D][128]  scaled rows of the tile's row block
  double* xj = sh + (size_t)D * TS;
  // decode lower-triangular tile index
  const int nb = (int)(Np / TS);
  int bi = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((int64_t)(bi + 1) * (bi + 2) / 2 <= (int64_t)blockIdx.x) ++bi;
  while ((int64_t)bi * (bi + 1) / 2 > (int64_t)blockIdx.x) --bi;
  const int bj = (int)(blockIdx.x - (int64_t)bi * (bi + 1) / 2);
  (void)nb;
  const int tid = threadIdx.x;
  const double ea = par[0], sn2 = par[1];
  for (int e = tid; e < TS * D; e += 256) {
    const int r = e / D, d = e - r * D;
    const double il = par[2 + d];
    xi[d * TS + r] = X[((int64_t)bi * TS + r) * D + d] * il;
    xj[d * TS + r] = X[((int64_t)bj * TS + r) * D + d] * il;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  // two column halves in turn (8 x 4 accumulators instead of 8 x 8): 94 registers instead of 172, so two CTAs share
  // an SM and one CTA's exp chains / stores overlap the other's distance loop (ncu: FP64 pipe 42 % at one CTA per SM)
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    double acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int d = 0; d < D; ++d) {
      double a[8], b[4];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = xi[d * TS + ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        b[2 * c] = xj[d * TS + 32 * (2 * h + c) + 2 * tx];
        b[2 * c + 1] = xj[d * TS + 32 * (2 * h + c) + 2 * tx + 1];
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double df = a[r] - b[c];
          acc[r][c] = fma(df, df, acc[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int64_t i = (int64_t)bi * TS + ty + 16 * r;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t j = (int64_t)bj * TS + 32 * (2 * h + c) + 2 * tx;
        double2 v;
        v.x = ea * exp_neg(-0.5 * acc[r][2 * c]);
        v.y = ea * exp_neg(-0.5 * acc[r][2 * c + 1]);
        if (i >= N || j >= N) v.x = 0.0;
        if (i >= N || j + 1 >= N) v.y = 0.0;
        if (i == j) v.x = (i < N) ? v.x + sn2 : 1.0;
        if (i == j + 1) v.y = (i < N) ? v.y + sn2 : 1.0;
        *reinterpret_cast<double2*>(K + i * Np + j) = v;
      }
    }
  }
}

// Rectangular cross-Gram out[n, m] (row stride ldo) = ARD(x, xp): bounds-checked, no noise.
__global__ void __launch_bounds__(256)
gram_rect_kernel(const double* __restrict__ x, int64_t n, const double* __restrict__ xp, int64_t m, int D,
                 const double* __restrict__ par, double* __restrict__ out, int64_t ldo) {
  extern __shared__ double sh[];
  double* xi = sh;
  double* xj = sh + (size_t)D * TS;
  const int64_t i0 = (int64_t)blockIdx.y * TS, j0 = (int64_t)blockIdx.x * TS;
  const int tid = threadIdx.x;
  const double ea = par[0];
  for (int e = tid; e < TS * D; e += 256) {
    const int r = e / D, d = e - r * D;
    const double il = par[2 + d];
    xi[d * TS + r] = (i0 + r < n) ? x[(i0 + r) * D + d] * il : 0.0;
    xj[d * TS + r] = (j0 + r < m) ? xp[(j0 + r) * D + d] * il : 0.0;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  double acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;
  for (int d = 0; d < D; ++d) {
    double a[8], b[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = xi[d * TS + ty + 16 * r];
#pragma unroll
    for (int c = 0; c < 8; ++c) b[c] = xj[d * TS + tx + 16 * c];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double df = a[r] - b[c];
        acc[r][c] = fma(df, df, acc[r][c]);
      }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t i = i0 + ty + 16 * r;
    if (i >= n) continue;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int64_t j = j0 + tx + 16 * c;
      if (j < m) out[i * ldo + j] = ea * exp_neg(-0.5 * acc[r][c]);
    }
  }
}

}  // namespace

int gps_gram_sym(gps_ctx* ctx, const double* X, int64_t N, int64_t Np, int D, const double* d_par, double* K) {
  if (D > MAXD) return gps_fail(ctx, GPS_EINVAL, "D=%d exceeds %d", D, MAXD);
  const int64_t nb = Np / TS;
  const int64_t tiles = nb * (nb + 1) / 2;
  const size_t smem = (size_t)2 * D * TS * sizeof(double);
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(gram_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * MAXD * TS * sizeof(double))));
    GPS_CUDA(cudaFuncSetAttribute(gram_rect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * MAXD * TS * sizeof(double))));
    configured = true;
  }
  gram_sym_kernel<<<(unsigned)tiles, 256, smem, ctx->stream>>>(X, N, Np, D, d_par, K);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}

int gps_gram_rect(gps_ctx* ctx, const double* x, int64_t n, const double* xp, int64_t m, int D,
                  const double* d_par, double* out, int64_t ldo) {
  if (D > MAXD) return gps_fail(ctx, GPS_EINVAL, "D=%d exceeds %d", D, MAXD);
  if (n == 0 || m == 0) return GPS_OK;
  const size_t smem = (size_t)2 * D * TS * sizeof(double);
  GPS_ONCE_PER_DEVICE(ctx);
  if (!configured) {
    GPS_CUDA(cudaFuncSetAttribute(gram_rect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * MAXD * TS * sizeof(double))));
    configured = true;
  }
  dim3 grid((unsigned)((m + TS - 1) / TS), (unsigned)((n + TS - 1) / TS));
  gram_rect_kernel<<<grid, 256, smem, ctx->stream>>>(x, n, xp, m, D, d_par, out, ldo);
  GPS_LAUNCH_CHECK();
  ctx->launches++;
  return GPS_OK;
}
