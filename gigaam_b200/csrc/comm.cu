// The one exchange step of the path (north_star: "a single NCCL all-gather of transcripts over NVLink only when
// the batch is split"): every rank holds the hypotheses of its shard in ONE packed device buffer
//     [ ids  B_local x W | frames  B_local x W | counts  B_local ]   (int32)
// and gam_gather_hyps all-gathers that buffer to all ranks with a single ncclAllGather on the caller's stream -- it is
// stream-ordered behind the greedy kernels and can be captured into the same CUDA graph.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has -- torch ships one -- else the system
// one): the library stays loadable and every single-GPU entry point usable on hosts without NCCL; only the gam_comm_*
// calls fail there, loudly.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

#include "../../include/gigaam_b200.h"
#include "comm.h"

namespace gam {
namespace {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  bool ok = false;
  char why[256] = {0};
};

NcclApi& api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch (or anyone) already mapped
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    snprintf(a.why, sizeof(a.why), "libnccl.so.2 not loadable: %s", dlerror());
    return a;
  }
#define GAM_SYM(name)                                                             \
  a.name = reinterpret_cast<decltype(a.name)>(dlsym(lib, "nccl" #name));            \
  if (!a.name) { snprintf(a.why, sizeof(a.why), "nccl" #name " not found"); return a; }
  GAM_SYM(GetUniqueId) GAM_SYM(CommInitRank) GAM_SYM(AllGather) GAM_SYM(CommDestroy) GAM_SYM(GetErrorString) GAM_SYM(GetVersion)
#undef GAM_SYM
  a.ok = true;
  return a;
}

}  // namespace

const char* comm_unavailable_reason() { return api().ok ? nullptr : api().why; }

int comm_unique_id(unsigned char* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in every NCCL 2.x");
  if (!api().ok) return -1;
  ncclUniqueId id;
  if (api().GetUniqueId(&id) != ncclSuccess) return -2;
  memcpy(out128, &id, 128);
  return 0;
}

int comm_init(void** comm, const unsigned char* id128, int rank, int nranks, const char** err) {
  if (!api().ok) { *err = api().why; return -1; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  const ncclResult_t r = api().CommInitRank(&c, nranks, id, rank);
  if (r != ncclSuccess) { *err = api().GetErrorString(r); return -2; }
  *comm = c;
  return 0;
}

int comm_all_gather_i32(void* comm, const int* send, int* recv, long long count, cudaStream_t s, const char** err) {
  const ncclResult_t r = api().AllGather(send, recv, static_cast<size_t>(count), ncclInt32, static_cast<ncclComm_t>(comm), s);
  if (r != ncclSuccess) { *err = api().GetErrorString(r); return -2; }
  return 0;
}

void comm_destroy(void* comm) {
  if (comm && api().ok) api().CommDestroy(static_cast<ncclComm_t>(comm));
}

int comm_nccl_version() {
  int v = 0;
  if (api().ok) api().GetVersion(&v);
  return v;
}

}  // namespace gam
