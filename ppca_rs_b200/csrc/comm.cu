// comm.cu — the one collective of the path, behind the C ABI: an NCCL all-reduce of the additive statistics buffer.
//
// The reference has no distributed backend (rayon over samples on one host); the north star shards the
// samples over the GPUs of one box and sums [A | B | tdev | totals | scalars] once per iteration over NVLink.  A host
// written in any language gets that through ppca_b200_comm_init + the *_sharded entry points without linking torch.
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a process that already carries a copy (torch) the loader
// returns that copy, a bare Rust/C host gets the system library.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace ppca {

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char *names[] = {getenv("PPCA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
      err = dlerror() ? dlerror() : "dlopen failed";
    }
    if (!api.lib) return;
    auto sym = [&](const char *s) { return dlsym(api.lib, s); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) {
      err = "libnccl lacks a required symbol";
      api.lib = nullptr;
    }
  });
  if (!api.lib) PPCA_THROW(PPCA_ERR_CUDA, "NCCL is not available: %s", err.c_str());
  return api;
}

#define NCCL_CHECK(expr)                                                                                  \
  do {                                                                                                    \
    ncclResult_t r_ = (expr);                                                                             \
    if (r_ != ncclSuccess)                                                                                \
      PPCA_THROW(PPCA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)
}  // namespace

static_assert(sizeof(ncclUniqueId) == PPCA_B200_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");

void comm_unique_id(uint8_t *out) {
  ncclUniqueId id;
  NCCL_CHECK(nccl().GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
}

void *comm_create(const uint8_t *id_bytes, int rank, int world) {
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  NCCL_CHECK(nccl().CommInitRank(&comm, world, id, rank));
  return comm;
}

void comm_destroy(void *comm) {
  if (comm) nccl().CommDestroy(static_cast<ncclComm_t>(comm));
}

void comm_allreduce(void *comm, double *buf, int64_t count, int op, cudaStream_t stream) {
  if (count <= 0) return;
  NCCL_CHECK(nccl().AllReduce(buf, buf, (size_t)count, ncclDouble, op == 1 ? ncclMax : ncclSum,
                              static_cast<ncclComm_t>(comm), stream));
}

int comm_version() {
  int v = 0;
  if (nccl().GetVersion) nccl().GetVersion(&v);
  return v;
}

}  // namespace ppca
