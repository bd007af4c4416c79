// dfb200_nccl_*: thin C-ABI wrappers around NCCL for hosts that do not bring torch.distributed -- the
// gradient all-reduce of DistributedDataParallel (utils/train.py:238-243) on the flat gradient arena, on a caller-given stream.
// libnccl is resolved at run time (dlopen) so that the shared library neither links against nor pins an NCCL build: inside a
// PyTorch process the already-loaded torch-bundled libnccl.so.2 is picked up by its SONAME, a C/C++ host gets the system one.
// Only the handful of NCCL entry points below are used; their prototypes follow nccl.h (stable since NCCL 2.10, ncclAvg / ncclBfloat16).
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

struct UniqueId { char internal[128]; };                       // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed BY VALUE to ncclCommInitRank
typedef int (*GetVersionFn)(int*);
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*CommDestroyFn)(void*);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
  void* handle = nullptr;
  GetVersionFn get_version = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool ok = false;
  char why[256] = "";
};

Nccl g_nccl;
std::once_flag g_once;

void load_nccl() {
  const char* names[] = {getenv("DFB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    const char* e = dlerror();
    snprintf(g_nccl.why, sizeof(g_nccl.why), "libnccl not found (tried $DFB200_NCCL_LIB, libnccl.so.2, libnccl.so): %s", e ? e : "?");
    return;
  }
  g_nccl.get_version = reinterpret_cast<GetVersionFn>(dlsym(g_nccl.handle, "ncclGetVersion"));
  g_nccl.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(g_nccl.handle, "ncclGetUniqueId"));
  g_nccl.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(g_nccl.handle, "ncclCommInitRank"));
  g_nccl.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(g_nccl.handle, "ncclAllReduce"));
  g_nccl.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(g_nccl.handle, "ncclCommDestroy"));
  g_nccl.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(g_nccl.handle, "ncclGetErrorString"));
  if (!g_nccl.get_version || !g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy ||
      !g_nccl.error_string) {
    snprintf(g_nccl.why, sizeof(g_nccl.why), "libnccl lacks one of ncclGetVersion / GetUniqueId / CommInitRank / AllReduce / CommDestroy / GetErrorString");
    return;
  }
  g_nccl.ok = true;
}

// returns nullptr (and sets the error text) when NCCL cannot be used
const Nccl* nccl() {
  std::call_once(g_once, load_nccl);
  if (!g_nccl.ok) {
    dfb_set_error("nccl: %s", g_nccl.why);
    return nullptr;
  }
  return &g_nccl;
}

int check(const Nccl* n, int rc, const char* what) {
  if (rc == 0) return DFB_OK;                                      // ncclSuccess
  dfb_set_error("%s: %s (ncclResult %d)", what, n->error_string(rc), rc);
  return DFB_ERR_CUDA;
}

}  // namespace

extern "C" int dfb200_nccl_version(int* version) {
  const Nccl* n = nccl();
  if (!n) return DFB_ERR_UNSUPPORTED;
  if (!version) { dfb_set_error("nccl_version: null pointer"); return DFB_ERR_ARG; }
  return check(n, n->get_version(version), "ncclGetVersion");
}

extern "C" int dfb200_nccl_unique_id(void* id128) {
  const Nccl* n = nccl();
  if (!n) return DFB_ERR_UNSUPPORTED;
  if (!id128) { dfb_set_error("nccl_unique_id: null pointer"); return DFB_ERR_ARG; }
  UniqueId id;
  int rc = check(n, n->get_unique_id(&id), "ncclGetUniqueId");
  if (rc == DFB_OK) memcpy(id128, id.internal, sizeof(id.internal));
  return rc;
}

extern "C" int dfb200_nccl_comm_init(const void* id128, int nranks, int rank, void** comm) {
  const Nccl* n = nccl();
  if (!n) return DFB_ERR_UNSUPPORTED;
  if (!id128 || !comm || nranks < 1 || rank < 0 || rank >= nranks) { dfb_set_error("nccl_comm_init: bad arguments (nranks %d, rank %d)", nranks, rank); return DFB_ERR_ARG; }
  UniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  *comm = nullptr;
  return check(n, n->comm_init_rank(comm, nranks, id, rank), "ncclCommInitRank");
}

extern "C" int dfb200_nccl_all_reduce(void* comm, void* buf, long count, int dtype, int average, void* stream) {
  const Nccl* n = nccl();
  if (!n) return DFB_ERR_UNSUPPORTED;
  if (!comm || (!buf && count > 0) || count < 0) { dfb_set_error("nccl_all_reduce: bad arguments"); return DFB_ERR_ARG; }
  if (count == 0) return DFB_OK;
  int nd;                                                        // ncclDataType_t
  if (dtype == DFB200_F32) nd = 7;                               // ncclFloat32
  else if (dtype == DFB200_BF16) nd = 9;                         // ncclBfloat16
  else if (dtype == 2) nd = 8;                                   // ncclFloat64 (the SyncBatchNorm statistics)
  else { dfb_set_error("nccl_all_reduce: dtype %d (0 = float32, 1 = bfloat16, 2 = float64)", dtype); return DFB_ERR_ARG; }
  const int op = average ? 4 : 0;                                // ncclAvg : ncclSum
  return check(n, n->all_reduce(buf, buf, static_cast<size_t>(count), nd, op, comm, reinterpret_cast<cudaStream_t>(stream)), "ncclAllReduce");
}

extern "C" int dfb200_nccl_comm_destroy(void* comm) {
  const Nccl* n = nccl();
  if (!n) return DFB_ERR_UNSUPPORTED;
  if (!comm) return DFB_OK;
  return check(n, n->comm_destroy(comm), "ncclCommDestroy");
}
