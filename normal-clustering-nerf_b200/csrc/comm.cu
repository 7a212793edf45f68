// Data-parallel gradient all-reduce over NVLink 5 / NVSwitch.  Replaces torch DDP's bucketed
// all-reduce (train_nerf.py:949-952): ONE ncclAllReduce(sum) on the flat fp32 gradient buffer
// [hash table | MLPs], enqueued on the caller's stream right behind the last backward kernel.
// NCCL is bound at run time (dlopen "libnccl.so.2": in a torch process that is torch's bundled
// copy, already loaded), so libncn.so itself has no link-time dependency on it.
#include "ncn_common.cuh"
#include <dlfcn.h>
#include <string.h>
#include <stdio.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclFloat32 = 7 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;
char g_err[256] = "";

bool load_api() {
  if (g_api.handle) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) { snprintf(g_err, sizeof g_err, "dlopen(libnccl.so.2) failed: %s", dlerror()); return false; }
  g_api.GetUniqueId = (decltype(g_api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_api.CommInitRank = (decltype(g_api.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_api.AllReduce = (decltype(g_api.AllReduce))dlsym(h, "ncclAllReduce");
  g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllReduce || !g_api.CommDestroy) {
    snprintf(g_err, sizeof g_err, "libnccl is missing a required symbol");
    return false;
  }
  g_api.handle = h;
  return true;
}

int nccl_fail(ncclResult_t r, const char* what) {
  snprintf(g_err, sizeof g_err, "%s: %s", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "nccl error");
  return NCN_E_NCCL;
}

}  // namespace

struct ncn_comm {
  ncclComm_t comm;
  int world_size, rank;
};

extern "C" const char* ncn_comm_last_error(void) { return g_err; }

extern "C" int ncn_comm_unique_id(void* id128_host) {
  NCN_CHECK_PTR(id128_host);
  if (!load_api()) return NCN_E_NCCL;
  ncclUniqueId id;
  ncclResult_t r = g_api.GetUniqueId(&id);
  if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(id128_host, &id, sizeof id);
  return NCN_OK;
}

extern "C" int ncn_comm_init(ncn_comm** comm, const void* id128_host, int world_size, int rank) {
  NCN_CHECK_PTR(comm); NCN_CHECK_PTR(id128_host);
  NCN_CHECK_SIZE(world_size >= 1 && rank >= 0 && rank < world_size);
  if (!load_api()) return NCN_E_NCCL;
  ncclUniqueId id;
  memcpy(&id, id128_host, sizeof id);
  ncn_comm* c = new ncn_comm;
  c->world_size = world_size; c->rank = rank; c->comm = nullptr;
  ncclResult_t r = g_api.CommInitRank(&c->comm, world_size, id, rank);
  if (r != 0) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
  *comm = c;
  return NCN_OK;
}

extern "C" int ncn_comm_allreduce_sum_f32(ncn_comm* comm, float* buf, int64_t n, ncn_stream_t stream) {
  NCN_CHECK_PTR(comm);
  NCN_CHECK_SIZE(n >= 0);
  if (n == 0 || comm->world_size == 1) return NCN_OK;
  NCN_CHECK_PTR(buf);
  ncclResult_t r = g_api.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, comm->comm, ncn::as_stream(stream));
  if (r != 0) return nccl_fail(r, "ncclAllReduce");
  return NCN_OK;
}

extern "C" int ncn_comm_destroy(ncn_comm* comm) {
  if (!comm) return NCN_OK;
  if (comm->comm && g_api.CommDestroy) g_api.CommDestroy(comm->comm);
  delete comm;
  return NCN_OK;
}
