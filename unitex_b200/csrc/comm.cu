// The multi-GPU plane of the library.
//  * utx_allgather_tiles: the batch-sharded path's one exchange -- every rank contributes its finished uint8 view tile and
//    receives all of them (north_star: "a single NCCL all-gather of decoded tiles before UV projection").  Thin on purpose: one
//    ncclAllGather over NVLink / NVSwitch of a few MB per asset batch, latency-bound, nothing to fuse it with (the VAE decode
//    before it and the bake after it are per-asset local).
//  * utx_comm_alltoall / comm_allgather: the collectives of the engine's sequence-parallel mode (flux_engine.cu).
//  * utx_peer_* + peer_barrier: CUDA-IPC peer regions and a flag barrier over them for the FUSED form of that mode, where the
//    QKV GEMM's and the attention kernel's epilogues store straight into the peers' buffers and no collective is left around the
//    attention.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy torch already mapped when the caller is a torch process, the
// system one otherwise), so the library links and loads on a box without NCCL and only these entry points fail there.
// Bootstrap follows NCCL's own model: one rank makes a 128-byte unique id, the host side shares it out of band
// (torch.distributed broadcast, MPI, a file), every rank joins with (id, nranks, rank).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../include/unitex_b200.h"
#include "common.h"
#include "ptx.cuh"

using namespace utx;

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0 };   // ncclChar / ncclInt8 == 0 (nccl.h: ncclDataType_t)

struct Nccl {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_LAZY | RTLD_GLOBAL);
    if (!h) return;
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    n.Send = reinterpret_cast<decltype(n.Send)>(dlsym(h, "ncclSend"));
    n.Recv = reinterpret_cast<decltype(n.Recv)>(dlsym(h, "ncclRecv"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(dlsym(h, "ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.GetErrorString && n.Send && n.Recv && n.GroupStart &&
           n.GroupEnd;
  });
  return n;
}

#define UTX_NCCL(expr)                                                                                        \
  do {                                                                                                        \
    ncclResult_t _r = (expr);                                                                                 \
    if (_r != 0) {                                                                                            \
      set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": NCCL: " + nccl().GetErrorString(_r)); \
      return 3;                                                                                               \
    }                                                                                                         \
  } while (0)

}  // namespace

struct utx_comm {
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = 0;
};

namespace {
// Cross-GPU barrier over peer memory (NVLink P2P): thread p of one CTA publishes this rank's epoch in peer p's flag array
// (after a system-scope fence, so every store this GPU issued before -- the previous kernels' writes into the peers' buffers
// included: they are complete at the kernel boundary -- is visible first) and then waits until peer p has published the same
// epoch here.  Bounded by the watchdog: a rank that never arrives becomes a trapped launch error on the others, not a hung box.
__global__ void __launch_bounds__(32) peer_barrier_kernel(unsigned* const* __restrict__ peer_flags, unsigned* local_flags, int rank,
                                                          int nranks, unsigned epoch) {
  const int p = threadIdx.x;
  if (p >= nranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[p] + rank), "r"(epoch) : "memory");
  const uint64_t t0 = globaltimer_ns();
  unsigned v = 0, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local_flags + p) : "memory");
    if (static_cast<int>(v - epoch) >= 0) break;
    if ((++spins & 0xffu) == 0 && globaltimer_ns() - t0 > UTX_WATCHDOG_NS) {
      printf("utx: peer barrier watchdog: rank %d waits for rank %d, epoch %u (has %u)\n", rank, p, epoch, v);
      __trap();
    }
  }
}
}  // namespace

namespace utx {
// flags_dev: device array [nranks] of pointers to every rank's flag array (this process's mappings); local = flags_dev[rank]
int peer_barrier(unsigned* const* flags_dev, unsigned* local_flags, int rank, int nranks, unsigned epoch, cudaStream_t stream) {
  UTX_CHECK(nranks >= 1 && nranks <= 32, "peer_barrier: 1..32 ranks");
  peer_barrier_kernel<<<1, 32, 0, stream>>>(flags_dev, local_flags, rank, nranks, epoch);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// used by the FLUX engine's sequence-parallel mode (flux_engine.cu)
int comm_nranks(const utx_comm* c) { return c ? c->nranks : 1; }
int comm_rank(const utx_comm* c) { return c ? c->rank : 0; }
// send[p * bytes .. ] goes to rank p, recv[p * bytes ..] comes from rank p (grouped ncclSend / ncclRecv: an all-to-all)
int comm_alltoall(utx_comm* c, const void* send, void* recv, size_t bytes_per_peer, cudaStream_t stream) {
  UTX_CHECK(c && c->comm, "comm_alltoall: communicator not initialised");
  if (bytes_per_peer == 0) return 0;
  UTX_NCCL(nccl().GroupStart());
  for (int p = 0; p < c->nranks; ++p) {
    UTX_NCCL(nccl().Send(static_cast<const char*>(send) + static_cast<size_t>(p) * bytes_per_peer, bytes_per_peer, ncclInt8, p, c->comm, stream));
    UTX_NCCL(nccl().Recv(static_cast<char*>(recv) + static_cast<size_t>(p) * bytes_per_peer, bytes_per_peer, ncclInt8, p, c->comm, stream));
  }
  UTX_NCCL(nccl().GroupEnd());
  return 0;
}
int comm_allgather(utx_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream) {
  UTX_CHECK(c && c->comm, "comm_allgather: communicator not initialised");
  if (bytes_per_rank == 0) return 0;
  UTX_NCCL(nccl().AllGather(send, recv, bytes_per_rank, ncclInt8, c->comm, stream));
  return 0;
}
}  // namespace utx

extern "C" {

int utx_comm_unique_id(void* id128) {
  UTX_CHECK(id128, "utx_comm_unique_id: null pointer");
  UTX_CHECK(nccl().ok, "utx_comm: libnccl.so.2 not found");
  ncclUniqueId id;
  UTX_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

int utx_comm_init(utx_comm** out, const void* id128, int nranks, int rank) {
  UTX_CHECK(out && id128 && nranks > 0 && rank >= 0 && rank < nranks, "utx_comm_init: bad argument");
  UTX_CHECK(nccl().ok, "utx_comm: libnccl.so.2 not found");
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  utx_comm* c = new utx_comm();
  c->nranks = nranks;
  c->rank = rank;
  ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);   // uses the CURRENT device, like every other utx_ call
  if (r != 0) {
    set_error(std::string("utx_comm_init: NCCL: ") + nccl().GetErrorString(r));
    delete c;
    return 3;
  }
  *out = c;
  return 0;
}

void utx_comm_destroy(utx_comm* c) {
  if (!c) return;
  if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
  delete c;
}

/* ---- peer memory: a device buffer every rank of the box can address (CUDA IPC over NVLink P2P) ---- */
int utx_peer_alloc(void** ptr, size_t bytes) {
  UTX_CHECK(ptr && bytes > 0, "utx_peer_alloc: bad argument");
  UTX_CUDA(cudaMalloc(ptr, bytes));
  UTX_CUDA(cudaMemset(*ptr, 0, bytes));
  return 0;
}
void utx_peer_free(void* ptr) {
  if (ptr) cudaFree(ptr);
}
int utx_peer_export(void* ptr, void* handle64) {
  UTX_CHECK(ptr && handle64, "utx_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hnd;
  UTX_CUDA(cudaIpcGetMemHandle(&hnd, ptr));
  std::memcpy(handle64, &hnd, 64);
  return 0;
}
int utx_peer_import(const void* handle64, void** ptr) {
  UTX_CHECK(ptr && handle64, "utx_peer_import: null pointer");
  cudaIpcMemHandle_t hnd;
  std::memcpy(&hnd, handle64, 64);
  UTX_CUDA(cudaIpcOpenMemHandle(ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
void utx_peer_close(void* ptr) {
  if (ptr) cudaIpcCloseMemHandle(ptr);
}

int utx_comm_alltoall(utx_comm* c, const void* send, void* recv, size_t bytes_per_peer, void* stream) {
  UTX_CHECK(send && recv, "utx_comm_alltoall: null pointer");
  return comm_alltoall(c, send, recv, bytes_per_peer, static_cast<cudaStream_t>(stream));
}

int utx_allgather_tiles(utx_comm* c, const void* tile, void* out, size_t bytes_per_rank, void* stream) {
  UTX_CHECK(c && c->comm, "utx_allgather_tiles: communicator not initialised");
  UTX_CHECK(tile && out, "utx_allgather_tiles: null pointer");
  if (bytes_per_rank == 0) return 0;
  UTX_NCCL(nccl().AllGather(tile, out, bytes_per_rank, ncclInt8, c->comm, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
