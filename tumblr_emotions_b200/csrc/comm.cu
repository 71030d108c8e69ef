// Data-parallel collective of the training step behind the C ABI: one flat NCCL all-reduce (sum, fp32) over the gradient arena.
//
// Semantics follow the reference's only multi-device precedent, slim/deployment/model_deploy.py: every clone holds a full
// replica, per-clone losses are scaled by 1/num_clones (:220-223), the regularisation loss is added once (:301-302) and the
// gradients of the shared variables are SUMMED across clones (:414-444).  Here a clone is a rank (one process per GPU): the
// caller reduces the per-rank gradient arenas with ds_allreduce_sum_f32, folds 1/world into ds_adam's grad_scale and adds the
// L2 gradient once after the reduction (engine.py).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 - the copy PyTorch already mapped into the process when there is one), so
// libdeepsent.so carries no link-time dependency on it and still loads on a box without NCCL; ds_comm_* then fail with a message.
// ncclAllReduce is stream-ordered and CUDA-graph capturable: the engine captures it on a side stream inside the step graph so
// that the backward pass of the frozen layers hides it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include "common.cuh"

struct ds_comm {
  ncclComm_t comm;
  int rank, world;
};

namespace {

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

Nccl g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already in the process (PyTorch's), if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return ds::fail("NCCL not found: dlopen(libnccl.so.2) -> %s", dlerror());
#define DS_SYM(field, name)                                                              \
  do {                                                                                   \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                           \
    if (!g_nccl.field) return ds::fail("NCCL symbol %s missing: %s", name, dlerror());   \
  } while (0)
  DS_SYM(GetUniqueId, "ncclGetUniqueId");
  DS_SYM(CommInitRank, "ncclCommInitRank");
  DS_SYM(AllReduce, "ncclAllReduce");
  DS_SYM(CommDestroy, "ncclCommDestroy");
  DS_SYM(GetErrorString, "ncclGetErrorString");
  DS_SYM(GetVersion, "ncclGetVersion");
#undef DS_SYM
  g_nccl.handle = h;
  return 0;
}

#define DS_NCCL(expr)                                                                                       \
  do {                                                                                                      \
    ncclResult_t _r = (expr);                                                                               \
    if (_r != ncclSuccess) return ds::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
  } while (0)

}  // namespace

extern "C" {

int ds_comm_unique_id(uint8_t* id128) {
  DS_REQUIRE(id128 != nullptr, "id buffer is NULL");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (int r = load_nccl()) return r;
  ncclUniqueId id;
  DS_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int ds_comm_init(ds_comm** comm, int rank, int world, const uint8_t* id128) {
  DS_REQUIRE(comm != nullptr && id128 != nullptr, "NULL argument");
  DS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank must be in [0, world)");
  if (int r = load_nccl()) return r;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  DS_NCCL(g_nccl.CommInitRank(&c, world, id, rank));      // binds to the calling thread's current device (ds_init)
  *comm = new ds_comm{c, rank, world};
  return 0;
}

int ds_allreduce_sum_f32(ds_comm* comm, float* buf, int64_t n, void* stream) {
  DS_REQUIRE(comm != nullptr && comm->comm != nullptr, "communicator is not initialised");
  DS_REQUIRE(n >= 0 && (n == 0 || buf != nullptr), "bad buffer");
  if (n == 0) return 0;
  DS_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, comm->comm, ds::S(stream)));
  return 0;
}

int ds_comm_destroy(ds_comm* comm) {
  if (!comm) return 0;
  if (comm->comm && g_nccl.CommDestroy) {
    ncclResult_t r = g_nccl.CommDestroy(comm->comm);
    if (r != ncclSuccess) { delete comm; return ds::fail("ncclCommDestroy -> %s", g_nccl.GetErrorString(r)); }
  }
  delete comm;
  return 0;
}

int ds_comm_nccl_version(void) {
  if (load_nccl()) return 0;
  int v = 0;
  return g_nccl.GetVersion(&v) == ncclSuccess ? v : 0;
}

}  // extern "C"
