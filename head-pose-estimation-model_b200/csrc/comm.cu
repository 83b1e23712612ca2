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

extern "C" int hp_p2p_close(hp_handle h);

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
  if (h->comm.inbox) hp_p2p_close(h);
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

// ============================================================================ peer-memory exchange for the fused all-reduce
// One ncclAllReduce of a 0.2k-52k float gradient buffer is pure latency (measured inside the captured step: 12 us at 2 ranks,
// 35 us at 8).  The fused kernel of heads.cu (p2p_allreduce_optimizer_kernel) instead PUSHES every rank's gradient slice into
// every peer's inbox with plain NVLink stores, raises a per-slice flag, waits for the world's flags of its own slice, sums the
// copies in rank order (every rank gets bit-identical sums) and applies the optimizer -- one kernel, no second launch.
// Inbox layout (floats): [2 parities][world sources][cap + HP_P2P_SLICES flag words].
size_t hp_p2p_slot_floats(int cap) { return (size_t)cap + HP_P2P_SLICES; }

extern "C" int hp_p2p_alloc(hp_handle h, int cap_floats, int nranks, void* ipc_handle64_host) {
  HP_REQUIRE(h && ipc_handle64_host && cap_floats > 0 && nranks >= 2 && nranks <= HP_P2P_MAX_RANKS, HP_ERR_INVALID,
             "hp_p2p_alloc: bad arguments (2..%d ranks)", HP_P2P_MAX_RANKS);
  HP_CUDA(cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (h->comm.inbox) cudaFree(h->comm.inbox);
  h->comm.inbox = nullptr;
  h->comm.p2p_ready = false;
  const size_t bytes = 2 * (size_t)nranks * hp_p2p_slot_floats(cap_floats) * sizeof(float);
  HP_CUDA(cudaMalloc(&h->comm.inbox, bytes));
  HP_CUDA(cudaMemset(h->comm.inbox, 0, bytes));          // flags start at step 0: the first step waits for 1
  cudaIpcMemHandle_t hd;
  HP_CUDA(cudaIpcGetMemHandle(&hd, h->comm.inbox));
  memcpy(ipc_handle64_host, &hd, sizeof(hd));
  h->comm.p2p_cap = cap_floats;
  return HP_OK;
}

extern "C" int hp_p2p_open(hp_handle h, const void* all_handles_host, int rank, int nranks) {
  HP_REQUIRE(h && all_handles_host && h->comm.inbox && nranks >= 2 && nranks <= HP_P2P_MAX_RANKS && rank >= 0 && rank < nranks,
             HP_ERR_INVALID, "hp_p2p_open: bad arguments or hp_p2p_alloc not called");
  HP_CUDA(cudaSetDevice(h->device));
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      h->comm.peers[r] = h->comm.inbox;
      continue;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)all_handles_host + 64 * (size_t)r, sizeof(hd));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    HP_REQUIRE(e == cudaSuccess, HP_ERR_CUDA, "hp_p2p_open: cannot map the inbox of rank %d (%s): no peer access between the GPUs?", r,
               cudaGetErrorString(e));
    h->comm.peers[r] = ptr;
  }
  if (!h->comm.peers_dev) HP_CUDA(cudaMalloc((void**)&h->comm.peers_dev, HP_P2P_MAX_RANKS * sizeof(void*)));
  if (!h->comm.p2p_seq) HP_CUDA(cudaMalloc((void**)&h->comm.p2p_seq, sizeof(unsigned int)));
  HP_CUDA(cudaMemset(h->comm.p2p_seq, 0, sizeof(unsigned int)));
  HP_CUDA(cudaMemcpy(h->comm.peers_dev, h->comm.peers, HP_P2P_MAX_RANKS * sizeof(void*), cudaMemcpyHostToDevice));
  h->comm.rank = rank;
  h->comm.nranks = nranks;
  h->comm.p2p_ready = true;
  return HP_OK;
}

extern "C" int hp_p2p_close(hp_handle h) {
  HP_REQUIRE(h, HP_ERR_INVALID, "hp_p2p_close: null handle");
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < HP_P2P_MAX_RANKS; ++r) {
    if (h->comm.peers[r] && h->comm.peers[r] != h->comm.inbox) cudaIpcCloseMemHandle(h->comm.peers[r]);
    h->comm.peers[r] = nullptr;
  }
  if (h->comm.peers_dev) cudaFree(h->comm.peers_dev);
  h->comm.peers_dev = nullptr;
  if (h->comm.p2p_seq) cudaFree(h->comm.p2p_seq);
  h->comm.p2p_seq = nullptr;
  if (h->comm.inbox) cudaFree(h->comm.inbox);
  h->comm.inbox = nullptr;
  h->comm.p2p_ready = false;
  return HP_OK;
}
