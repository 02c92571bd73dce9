// The collective of the row-sharded FITC evaluation, inside the library: one NCCL communicator per context
// (one process per GPU), all-reduces of the packed accumulators enqueued on the context's stream between the
// row passes — no host synchronisation, no Python callback on the data path.
//
// libnccl is bound at run time (dlopen): libgpscore.so has no link-time dependency on it, so the library still
// loads on a box without NCCL and single-GPU use never touches it.  Lookup order: the copy the process already
// has (torch's bundled libnccl.so.2, found with RTLD_NOLOAD), a path given with gps_comm_set_library, then the
// system's libnccl.so.2.  Only the C API that has been stable since NCCL 2.0 is used.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <vector>

#include "gps_common.cuh"

struct gps_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory exchange area (see gps_common.cuh)
  double* xbuf = nullptr;                 // this rank's area (cudaMalloc)
  std::vector<void*> peer_map;            // cudaIpcOpenMemHandle mappings of the other ranks' areas
  double** d_peers = nullptr;             // device copy of the base pointers
  unsigned long long seq = 1;
  int transport = 1;                      // 1: peer memory inside the kernels when available, 0: NCCL only
  double* d_rows = nullptr;               // [world] row counts of the ranks (block objectives: fold geometry)
};

namespace {

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string path_hint, err;
};

NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.lib) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h && !g_nccl.path_hint.empty()) h = dlopen(g_nccl.path_hint.c_str(), RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    const char* e = dlerror();
    g_nccl.err = e ? e : "dlopen(libnccl.so.2) failed";
    return false;
  }
#define GPS_NCCL_SYM(name)                                                     \
  g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(h, "nccl" #name)); \
  if (!g_nccl.name) {                                                          \
    g_nccl.err = "libnccl lacks nccl" #name;                                   \
    return false;                                                              \
  }
  GPS_NCCL_SYM(GetUniqueId)
  GPS_NCCL_SYM(CommInitRank)
  GPS_NCCL_SYM(AllReduce)
  GPS_NCCL_SYM(CommDestroy)
  GPS_NCCL_SYM(GetErrorString)
  GPS_NCCL_SYM(GetVersion)
#undef GPS_NCCL_SYM
  g_nccl.lib = h;
  return true;
}

}  // namespace

static size_t xbuf_doubles(int world) { return (size_t)2 * world * GPS_P2P_SLOT + (size_t)2 * world + 16; }

// Exchange area + peer mappings.  The 64-byte IPC handles travel through the communicator itself (ncclAllGather):
// no other channel between the ranks is needed.  Any failure leaves peers unset and the NCCL path in use.
static void p2p_setup(gps_ctx* ctx) {
  gps_comm* cm = ctx->comm;
  const int W = cm->world;
  if (W < 2 || W > 16) return;
  decltype(&ncclAllGather) AllGather = reinterpret_cast<decltype(&ncclAllGather)>(dlsym(g_nccl.lib, "ncclAllGather"));
  if (!AllGather) return;
  const size_t nd = xbuf_doubles(W);
  if (cudaMalloc(&cm->xbuf, nd * sizeof(double)) != cudaSuccess) { cudaGetLastError(); cm->xbuf = nullptr; return; }
  cudaMemset(cm->xbuf, 0, nd * sizeof(double));
  cudaIpcMemHandle_t mine;
  char* d_h = nullptr;
  std::vector<cudaIpcMemHandle_t> all(W);
  bool ok = cudaIpcGetMemHandle(&mine, cm->xbuf) == cudaSuccess &&
            cudaMalloc(&d_h, (size_t)(W + 1) * sizeof mine) == cudaSuccess &&
            cudaMemcpy(d_h + (size_t)W * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice) == cudaSuccess &&
            AllGather(d_h + (size_t)W * sizeof mine, d_h, sizeof mine, ncclChar, cm->comm, ctx->stream) == ncclSuccess &&
            cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
            cudaMemcpy(all.data(), d_h, (size_t)W * sizeof mine, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (d_h) cudaFree(d_h);
  std::vector<double*> bases(W, nullptr);
  for (int r = 0; ok && r < W; ++r) {
    if (r == cm->rank) { bases[r] = cm->xbuf; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
    cm->peer_map.push_back(p);
    bases[r] = static_cast<double*>(p);
  }
  // every rank must agree on the outcome: a single failure switches all of them to NCCL
  int* d_ok = nullptr;
  int flag = ok ? 1 : 0, agreed = 0;
  if (cudaMalloc(&d_ok, sizeof(int)) == cudaSuccess) {
    cudaMemcpy(d_ok, &flag, sizeof(int), cudaMemcpyHostToDevice);
    if (g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, cm->comm, ctx->stream) == ncclSuccess &&
        cudaStreamSynchronize(ctx->stream) == cudaSuccess)
      cudaMemcpy(&agreed, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_ok);
  }
  cudaGetLastError();
  if (agreed == 1 && cudaMalloc(&cm->d_peers, (size_t)W * sizeof(double*)) == cudaSuccess) {
    cudaMemcpy(cm->d_peers, bases.data(), (size_t)W * sizeof(double*), cudaMemcpyHostToDevice);
  } else {
    cm->d_peers = nullptr;
  }
}

bool gps_comm_p2p_view(gps_ctx* ctx, gps_p2p_view* v, int exchanges) {
  gps_comm* cm = ctx->comm;
  if (!cm || !cm->d_peers || cm->transport != 1) return false;
  v->peers = cm->d_peers;
  v->rank = cm->rank;
  v->world = cm->world;
  v->seq = cm->seq;
  cm->seq += (unsigned long long)exchanges;
  return true;
}

void gps_comm_free(gps_ctx* ctx) {
  if (!ctx->comm) return;
  cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->comm->peer_map) cudaIpcCloseMemHandle(p);
  if (ctx->comm->d_peers) cudaFree(ctx->comm->d_peers);
  if (ctx->comm->xbuf) cudaFree(ctx->comm->xbuf);
  if (ctx->comm->d_rows) cudaFree(ctx->comm->d_rows);
  if (ctx->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
}

// in-place sum of n doubles over the communicator's ranks, enqueued on the context's stream
int gps_comm_allreduce(gps_ctx* ctx, double* buf, size_t n) {
  if (!ctx->comm || !ctx->comm->comm) return gps_fail(ctx, GPS_ESTATE, "no communicator: call gps_comm_init first");
  const ncclResult_t r = g_nccl.AllReduce(buf, buf, n, ncclDouble, ncclSum, ctx->comm->comm, ctx->stream);
  if (r != ncclSuccess) return gps_fail(ctx, GPS_ECUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
  return GPS_OK;
}

// First global row of this rank's block: the ranks hold consecutive blocks in rank order, their row counts are
// exchanged through the communicator (exact in doubles) and must add up to world_n.
static int row_offset_of(gps_ctx* ctx, int64_t world_n, int64_t* off) {
  gps_comm* cm = ctx->comm;
  const int W = cm->world;
  if (!cm->d_rows) GPS_CUDA(cudaMalloc(&cm->d_rows, (size_t)W * sizeof(double)));
  std::vector<double> h(W, 0.0);
  h[cm->rank] = (double)ctx->N;
  GPS_CUDA(cudaMemcpyAsync(cm->d_rows, h.data(), (size_t)W * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GPS_CHECK(gps_comm_allreduce(ctx, cm->d_rows, (size_t)W));
  GPS_CUDA(cudaMemcpyAsync(h.data(), cm->d_rows, (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPS_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t before = 0, total = 0;
  for (int r = 0; r < W; ++r) {
    if (r < cm->rank) before += (int64_t)h[r];
    total += (int64_t)h[r];
  }
  if (total != world_n)
    return gps_fail(ctx, GPS_EINVAL, "fitc_eval_sharded: the ranks hold %lld rows, world_n = %lld", (long long)total,
                    (long long)world_n);
  *off = before;
  return GPS_OK;
}

extern "C" {

int gps_comm_set_library(const char* path) {
  g_nccl.path_hint = path ? path : "";
  return GPS_OK;
}

int gps_comm_unique_id(void* out128) {
  if (!out128) return GPS_EINVAL;
  if (!nccl_load()) return GPS_ESTATE;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return GPS_ECUDA;
  memcpy(out128, &id, sizeof id);
  return GPS_OK;
}

int gps_comm_init(gps_ctx* ctx, const void* uid128, int rank, int world) {
  if (!ctx) return GPS_EINVAL;
  if (!uid128 || world < 1 || rank < 0 || rank >= world) return gps_fail(ctx, GPS_EINVAL, "comm_init: bad arguments");
  if (!nccl_load()) return gps_fail(ctx, GPS_ESTATE, "comm_init: %s", g_nccl.err.c_str());
  GPS_CUDA(cudaSetDevice(ctx->device));
  gps_comm_free(ctx);
  ctx->comm = new gps_comm();
  ctx->comm->rank = rank;
  ctx->comm->world = world;
  ncclUniqueId id;
  memcpy(&id, uid128, sizeof id);
  const ncclResult_t r = g_nccl.CommInitRank(&ctx->comm->comm, world, id, rank);
  if (r != ncclSuccess) {
    const int rc = gps_fail(ctx, GPS_ECUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    delete ctx->comm;
    ctx->comm = nullptr;
    return rc;
  }
  p2p_setup(ctx);
  return GPS_OK;
}

/* transport of the small-M row-sharded evaluation: 1 = one-shot all-reduce over peer memory inside the pass kernels
 * (default when the IPC mappings could be set up), 0 = ncclAllReduce between the kernels.  Returns the transport in
 * effect through *active (may be NULL). */
int gps_comm_set_transport(gps_ctx* ctx, int transport, int* active) {
  if (!ctx) return GPS_EINVAL;
  if (!ctx->comm) return gps_fail(ctx, GPS_ESTATE, "no communicator: call gps_comm_init first");
  ctx->comm->transport = transport ? 1 : 0;
  if (active) *active = (ctx->comm->transport == 1 && ctx->comm->d_peers) ? 1 : 0;
  return GPS_OK;
}

int gps_comm_info(gps_ctx* ctx, int* rank, int* world, int* nccl_version) {
  if (!ctx) return GPS_EINVAL;
  if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
  if (world) *world = ctx->comm ? ctx->comm->world : 1;
  if (nccl_version) {
    *nccl_version = 0;
    if (g_nccl.lib) g_nccl.GetVersion(nccl_version);
  }
  return GPS_OK;
}

int gps_comm_destroy(gps_ctx* ctx) {
  if (!ctx) return GPS_EINVAL;
  gps_comm_free(ctx);
  return GPS_OK;
}

// in-place sum of a device buffer over the ranks (test hook and building block of the sharded helpers)
int gps_comm_allreduce_sum(gps_ctx* ctx, double* buf, int64_t n) {
  if (!ctx) return GPS_EINVAL;
  if (!buf || n <= 0) return gps_fail(ctx, GPS_EINVAL, "allreduce: bad arguments");
  GPS_CUDA(cudaSetDevice(ctx->device));
  GPS_CHECK(gps_comm_allreduce(ctx, buf, (size_t)n));
  return GPS_OK;
}

// Row-sharded FITC objective + gradient: this context holds a contiguous block of the world_n rows; the packed
// accumulators of the three passes are all-reduced on the context's stream (replaces the three
// all-reduces a host would otherwise issue between gps_fitc_pass1 / pass2 / pass3).
int gps_fitc_eval_sharded(gps_ctx* ctx, const double* theta, const double* U, int M, double jitter, int score,
                          int64_t world_n, double* obj, double* grad_theta, double* grad_U) {
  if (!ctx) return GPS_EINVAL;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "fitc: call gps_set_data first");
  if (!theta || !U || M <= 0 || world_n < ctx->N) return gps_fail(ctx, GPS_EINVAL, "fitc_eval_sharded: bad arguments");
  if (score < GPS_CRPS || score > GPS_KC) return gps_fail(ctx, GPS_EINVAL, "fitc_eval_sharded: unknown score %d", score);
  if (!ctx->comm) return gps_fail(ctx, GPS_ESTATE, "fitc_eval_sharded: call gps_comm_init first");
  if (M < ctx->fitc_large_min_m && gps_fitc_fused_supports(ctx, M, score))
    return gps_fitc_fused_eval(ctx, theta, U, M, jitter, score, world_n, gps_comm_allreduce, obj, grad_theta, grad_U);
  // the block objectives (4-fold DSS, kc) shard through the matrix form for every M: their folds are ranges of the
  // GLOBAL row order, so the ranks' blocks must be consecutive in rank order
  int64_t row_offset = 0;
  if (score == GPS_DSS || score == GPS_KC) {
    GPS_CUDA(cudaSetDevice(ctx->device));
    GPS_CHECK(row_offset_of(ctx, world_n, &row_offset));
  }
  return gps_fitc_large_eval_sharded(ctx, theta, U, M, jitter, score, world_n, row_offset, gps_comm_allreduce, obj, grad_theta,
                                     grad_U);
}

// The optimiser loop K20:219-251 on a row-sharded problem: theta and U stay on every rank's device (all ranks apply the
// same update to the same all-reduced gradient), an iteration is the three pass kernels (exchange inside them, or
// ncclAllReduce between them) and the update kernel; one synchronisation and read-back after the last iteration.
int gps_fitc_descend_sharded(gps_ctx* ctx, double* theta, double* U, int M, double jitter, int score, int64_t world_n,
                             double lr_theta, double lr_u, int iters, double* obj_trace) {
  if (!ctx) return GPS_EINVAL;
  if (ctx->N == 0) return gps_fail(ctx, GPS_ESTATE, "fitc: call gps_set_data first");
  if (!theta || !U || M <= 0 || iters < 0 || world_n < ctx->N) return gps_fail(ctx, GPS_EINVAL, "fitc_descend_sharded: bad arguments");
  if (!ctx->comm) return gps_fail(ctx, GPS_ESTATE, "fitc_descend_sharded: call gps_comm_init first");
  if (!gps_fitc_fused_supports(ctx, M, score))
    return gps_fail(ctx, GPS_EINVAL, "fitc_descend_sharded: M <= 31 and crps / logs / nlml only");
  return gps_fitc_fused_descend(ctx, theta, U, M, jitter, score, lr_theta, lr_u, iters, obj_trace, world_n, gps_comm_allreduce);
}

}  // extern "C"
