// Data-parallel gradient exchange for head training: one ncclAllReduce(sum, fp32) over
// the flat [grads..., sum_sq, sum_abs, count] buffer per step, over NVLink/NVSwitch.
// NCCL is resolved at run time (dlsym on the already-loaded library, e.g. the one bundled with
// torch, else dlopen libnccl.so.2) so libhpose.so has no link-time dependency on it.
#include <dlfcn.h>

#include "common.cuh"

typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_get_uid)(nccl_uid_t*);
typedef int (*fn_init_rank)(void**, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

static struct {
  bool tried = false;
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

static int load_nccl() {
  if (g_nccl.tried) return g_nccl.allreduce ? HP_OK : HP_ERR_NCCL;
  g_nccl.tried = true;
  void* handles[3] = {RTLD_DEFAULT, nullptr, nullptr};
  handles[1] = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  handles[2] = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  for (int i = 0; i < 3; ++i) {
    void* hd = handles[i];
    if (i > 0 && !hd) continue;
    void* s = dlsym(hd, "ncclAllReduce");
    if (!s) continue;
    g_nccl.lib = hd;
    g_nccl.allreduce = (fn_allreduce)s;
    g_nccl.get_uid = (fn_get_uid)dlsym(hd, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(hd, "ncclCommInitRank");
    g_nccl.destroy = (fn_destroy)dlsym(hd, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(hd, "ncclGetErrorString");
    break;
  }
  if (!g_nccl.allreduce || !g_nccl.get_uid || !g_nccl.init_rank) {
    hp_set_error("NCCL not found (neither already loaded nor libnccl.so.2 on the loader path)");
    g_nccl.allreduce = nullptr;
    return HP_ERR_NCCL;
  }
  return HP_OK;
}

#define HP_NCCL(call)                                                                                  \
  do {                                                                                                 \
    int r_ = (call);                                                                                   \
    if (r_ != 0) {                                                                                     \
      hp_set_error("%s -> NCCL error %d (%s)", #call, r_, g_nccl.errstr ? g_nccl.errstr(r_) : "?");    \
      return HP_ERR_NCCL;                                                                              \
    }                                                                                                  \
  } while (0)

extern "C" int hp_comm_unique_id(void* id128_host) {
  HP_REQUIRE(id128_host, HP_ERR_INVALID, "hp_comm_unique_id: null pointer");
  HP_TRY(load_nccl());
  nccl_uid_t id;
  HP_NCCL(g_nccl.get_uid(&id));
  memcpy(id128_host, &id, sizeof(id));
  return HP_OK;
}

extern "C" int hp_comm_init(hp_handle h, const void* id_host, int rank, int nranks) {
  HP_REQUIRE(h && id_host && nranks >= 1 && rank >= 0 && rank < nranks, HP_ERR_INVALID, "hp_comm_init: bad arguments");
  HP_TRY(load_nccl());
  HP_CUDA(cudaSetDevice(h->device));
  nccl_uid_t id;
  memcpy(&id, id_host, sizeof(id));
  void* comm = nullptr;
  HP_NCCL(g_nccl.init_rank(&comm, nranks, id, rank));
  h->comm.comm = comm;
  h->comm.rank = rank;
  h->comm.nranks = nranks;
  return HP_OK;
}

extern "C" int hp_comm_destroy(hp_handle h) {
  HP_REQUIRE(h, HP_ERR_INVALID, "hp_comm_destroy: null handle");
  if (h->comm.comm && g_nccl.destroy) g_nccl.destroy(h->comm.comm);
  h->comm.comm = nullptr;
  h->comm.nranks = 1;
  return HP_OK;
}

int hp_comm_allreduce_sum(hp_ctx* h, float* buf, size_t n, cudaStream_t st) {
  HP_REQUIRE(h->comm.comm && g_nccl.allreduce, HP_ERR_STATE, "allreduce without hp_comm_init");
  // ncclFloat32 = 7, ncclSum = 0
  HP_NCCL(g_nccl.allreduce(buf, buf, n, 7, 0, h->comm.comm, st));
  return HP_OK;
}
