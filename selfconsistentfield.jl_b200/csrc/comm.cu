// K6 — cross-GPU sum of the partial [J | K_a | K_b] buffer: one
// ncclAllReduce(sum, fp64) of 8 n^2 (1 + n_spin) bytes per build, on the handle's stream.
// libnccl is dlopen'ed on first use so single-GPU users need no NCCL at all (and so the
// copy torch already mapped into the process is reused instead of a second one).
#include <dlfcn.h>
#include <nccl.h>

#include "scfb_internal.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl(scfb_handle_s* h) {
  if (g_nccl.lib) return SCFB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return scfb_fail(h, SCFB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  NcclApi api;
  api.lib = lib;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  api.GetErrorString =
      reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce ||
      !api.GetErrorString)
    return scfb_fail(h, SCFB_ERR_NCCL, "libnccl lacks a required symbol");
  g_nccl = api;
  return SCFB_OK;
}

}  // namespace

int scfb_comm_allreduce_jk(scfb_handle_s* h, int n_spin) {
  const size_t count = (size_t)(1 + n_spin) * h->n * h->n;
  ncclResult_t r = g_nccl.AllReduce(h->jk, h->jk_sum, count, ncclDouble, ncclSum,
                                    static_cast<ncclComm_t>(h->nccl_comm), h->stream);
  if (r != ncclSuccess)
    return scfb_fail(h, SCFB_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
  h->launches += 1;
  return SCFB_OK;
}

int scfb_comm_destroy(scfb_handle_s* h) {
  if (h->nccl_comm && g_nccl.CommDestroy) {
    g_nccl.CommDestroy(static_cast<ncclComm_t>(h->nccl_comm));
    h->nccl_comm = nullptr;
  }
  return SCFB_OK;
}

extern "C" {

int scfb_comm_unique_id(char* id128) {
  if (!id128) return scfb_fail(nullptr, SCFB_ERR_ARG, "null id buffer");
  if (int rc = load_nccl(nullptr)) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess)
    return scfb_fail(nullptr, SCFB_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  std::memcpy(id128, &id, 128);
  return SCFB_OK;
}

int scfb_comm_init(scfb_handle_t h, const char* id128) {
  if (!h) return scfb_fail(nullptr, SCFB_ERR_ARG, "null handle");
  SCFB_REQUIRE(h, id128, "null id buffer");
  if (h->n_ranks == 1) return SCFB_OK;  // nothing to exchange
  if (h->multi || scfb_peer_ready(h))
    return scfb_fail(h, SCFB_ERR_STATE, "this handle already has a peer-memory exchange (scfb_create_multi / "
                     "scfb_peer_import); a second one cannot be added");
  if (int rc = load_nccl(h)) return rc;
  SCFB_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, h->n_ranks, id, h->rank);
  if (r != ncclSuccess)
    return scfb_fail(h, SCFB_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  h->nccl_comm = comm;
  return SCFB_OK;
}

}  // extern "C"
