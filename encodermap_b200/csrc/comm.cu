// libemk's own NCCL communicator (emk_comm_*): the exchange steps of the two paths that shard.
//
//   * full-set cost (tile ranges, inputs replicated):  ONE fused launch sums the float64 loss and the (n, l) float32
//     gradient of all ranks (ncclGroupStart / two ncclAllReduce / ncclGroupEnd -- NCCL merges the group into one kernel,
//     where two torch.distributed.all_reduce calls are two launches with two stream hand-overs);
//   * per-batch cost inside data-parallel training (every rank owns n/G rows): one fused all-gather of the high-d rows
//     and the latent rows, and afterwards one fused {all-reduce loss, reduce-scatter dL/dz}.
//
// NCCL is bound at run time with dlopen (the process that calls in here -- torch -- has libnccl.so.2 loaded already), so
// libemk.so keeps linking against libc/libm only and single-GPU users never touch NCCL.  Types and enum values below are
// the stable NCCL 2.x ABI (nccl.h: ncclUniqueId is 128 bytes, ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "emk_common.cuh"

namespace emk {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

static NcclApi g_api;
static ncclComm_t g_comm = nullptr;
static int g_rank = -1, g_world = 0;
static std::mutex g_mu;

static int load_nccl() {
  if (g_api.handle) return EMK_OK;
  const char* names[] = {getenv("EMK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  EMK_REQUIRE(h != nullptr, EMK_E_UNSUPPORTED,
              "emk_comm: cannot dlopen libnccl.so.2 (%s); import torch first or set EMK_NCCL_LIB to the library path", dlerror());
#define EMK_SYM(field, name)                                                                       \
  do {                                                                                             \
    *reinterpret_cast<void**>(&g_api.field) = dlsym(h, name);                                      \
    EMK_REQUIRE(g_api.field != nullptr, EMK_E_UNSUPPORTED, "emk_comm: %s not found in NCCL", name); \
  } while (0)
  EMK_SYM(GetUniqueId, "ncclGetUniqueId");
  EMK_SYM(CommInitRank, "ncclCommInitRank");
  EMK_SYM(CommDestroy, "ncclCommDestroy");
  EMK_SYM(GetErrorString, "ncclGetErrorString");
  EMK_SYM(AllReduce, "ncclAllReduce");
  EMK_SYM(ReduceScatter, "ncclReduceScatter");
  EMK_SYM(AllGather, "ncclAllGather");
  EMK_SYM(GroupStart, "ncclGroupStart");
  EMK_SYM(GroupEnd, "ncclGroupEnd");
#undef EMK_SYM
  g_api.handle = h;
  return EMK_OK;
}

#define EMK_NCCL(call)                                                                                              \
  do {                                                                                                              \
    ncclResult_t r_ = (call);                                                                                       \
    if (r_ != 0) return fail(1000 + (int)r_, "%s failed: %s", #call, g_api.GetErrorString ? g_api.GetErrorString(r_) : "?"); \
  } while (0)

static int require_comm(const char* who) {
  EMK_REQUIRE(g_comm != nullptr, EMK_E_ARG, "%s: no communicator (call emk_comm_init on every rank first)", who);
  return EMK_OK;
}

}  // namespace emk

using namespace emk;

extern "C" {

int emk_comm_unique_id(void* id_out) {
  EMK_REQUIRE(id_out, EMK_E_NULL, "emk_comm_unique_id: NULL output");
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = load_nccl();
  if (rc) return rc;
  EMK_NCCL(g_api.GetUniqueId(static_cast<ncclUniqueId*>(id_out)));
  return EMK_OK;
}

int emk_comm_init(int rank, int world, const void* nccl_unique_id) {
  EMK_REQUIRE(nccl_unique_id, EMK_E_NULL, "emk_comm_init: NULL unique id");
  EMK_REQUIRE(world >= 1 && rank >= 0 && rank < world, EMK_E_ARG, "emk_comm_init: bad rank %d / world %d", rank, world);
  std::lock_guard<std::mutex> lock(g_mu);
  EMK_REQUIRE(g_comm == nullptr, EMK_E_ARG, "emk_comm_init: a communicator already exists (one per process; emk_comm_destroy first)");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  EMK_NCCL(g_api.CommInitRank(&g_comm, world, id, rank));
  g_rank = rank;
  g_world = world;
  return EMK_OK;
}

int emk_comm_info(int* rank, int* world) {
  if (rank) *rank = g_rank;
  if (world) *world = g_world;
  return g_comm ? EMK_OK : EMK_E_ARG;
}

int emk_comm_allreduce(double* loss, float* grad, int64_t grad_count, void* stream) {
  int rc = require_comm("emk_comm_allreduce");
  if (rc) return rc;
  EMK_REQUIRE(loss || grad, EMK_E_NULL, "emk_comm_allreduce: nothing to reduce");
  EMK_REQUIRE(grad_count >= 0, EMK_E_ARG, "emk_comm_allreduce: negative count");
  cudaStream_t st = as_stream(stream);
  EMK_NCCL(g_api.GroupStart());
  ncclResult_t r1 = 0, r2 = 0;
  if (loss) r1 = g_api.AllReduce(loss, loss, 1, kNcclFloat64, kNcclSum, g_comm, st);
  if (grad && grad_count > 0) r2 = g_api.AllReduce(grad, grad, (size_t)grad_count, kNcclFloat32, kNcclSum, g_comm, st);
  EMK_NCCL(g_api.GroupEnd());
  EMK_NCCL(r1);
  EMK_NCCL(r2);
  return EMK_OK;
}

int emk_comm_allgather2(const float* a_send, float* a_recv, int64_t a_count, const float* b_send, float* b_recv, int64_t b_count,
                        void* stream) {
  int rc = require_comm("emk_comm_allgather2");
  if (rc) return rc;
  EMK_REQUIRE(a_send && a_recv && a_count >= 0 && b_count >= 0, EMK_E_ARG, "emk_comm_allgather2: bad arguments");
  cudaStream_t st = as_stream(stream);
  EMK_NCCL(g_api.GroupStart());
  ncclResult_t r1 = 0, r2 = 0;
  if (a_count > 0) r1 = g_api.AllGather(a_send, a_recv, (size_t)a_count, kNcclFloat32, g_comm, st);
  if (b_send && b_recv && b_count > 0) r2 = g_api.AllGather(b_send, b_recv, (size_t)b_count, kNcclFloat32, g_comm, st);
  EMK_NCCL(g_api.GroupEnd());
  EMK_NCCL(r1);
  EMK_NCCL(r2);
  return EMK_OK;
}

int emk_comm_reduce_cost_scatter(double* loss, const float* grad_full, float* grad_mine, int64_t count_per_rank, void* stream) {
  int rc = require_comm("emk_comm_reduce_cost_scatter");
  if (rc) return rc;
  EMK_REQUIRE(grad_full && grad_mine && count_per_rank >= 0, EMK_E_ARG, "emk_comm_reduce_cost_scatter: bad arguments");
  cudaStream_t st = as_stream(stream);
  EMK_NCCL(g_api.GroupStart());
  ncclResult_t r1 = 0, r2 = 0;
  if (loss) r1 = g_api.AllReduce(loss, loss, 1, kNcclFloat64, kNcclSum, g_comm, st);
  if (count_per_rank > 0) r2 = g_api.ReduceScatter(grad_full, grad_mine, (size_t)count_per_rank, kNcclFloat32, kNcclSum, g_comm, st);
  EMK_NCCL(g_api.GroupEnd());
  EMK_NCCL(r1);
  EMK_NCCL(r2);
  return EMK_OK;
}

int emk_comm_destroy(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_comm) {
    ncclComm_t c = g_comm;
    g_comm = nullptr;
    g_rank = -1;
    g_world = 0;
    EMK_NCCL(g_api.CommDestroy(c));
  }
  return EMK_OK;
}

}  // extern "C"
