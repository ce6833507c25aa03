// Peer-memory plumbing for the sharded stage-2 optimiser: mapping another rank's allocation into this
// process (CUDA IPC over NVLink / NVSwitch) and a stream-ordered cross-rank barrier over peer-mapped flags.
// The data path itself lives in postopt.cu: its gather / level-0 kernels load UVT rows from, and reduce gradients into,
// whichever shard owns the row.  Replaces the autograd index_select / index_add_ + a dense gradient all-reduce
// (reference generate.py:496-517 run under data parallelism).
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"
#include "tclight.h"

namespace tcl {

static std::mutex g_ipc_mu;
static std::map<std::string, void*> g_ipc_open;
__device__ unsigned long long g_barrier_timeouts = 0ull;

struct PeerFlags { int* f[TCL_MAX_RANKS]; };

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// thread t < world: signal peer t, then wait for peer t's signal.  Everything this GPU wrote before the kernel (stream
// order) is performed system-wide before the release; the acquire orders the kernels that follow after the peers' writes.
__global__ void peer_barrier_kernel(PeerFlags pf, int world, int rank, int epoch) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  asm volatile("red.release.sys.global.add.s32 [%0], 1;" ::"l"(pf.f[t] + rank) : "memory");
  const int* mine = pf.f[rank] + t;
  const unsigned long long t0 = global_ns();
  int seen;
  for (;;) {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    if (seen >= epoch) break;
    if (global_ns() - t0 > 2000000000ull) { atomicAdd(&g_barrier_timeouts, 1ull); break; }
  }
  __threadfence_system();
}

}  // namespace tcl

using namespace tcl;

// A device allocation of its own (cudaMalloc, outside any caching allocator, zero-filled) plus the IPC handle of its base.
extern "C" int tcl_peer_alloc(size_t bytes, void** ptr, void* handle_out) {
  TCL_CHECK_ARG(bytes > 0 && ptr && handle_out, "tcl_peer_alloc: args");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (p) cudaFree(p);
    set_last_error("tcl_peer_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return TCL_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  *ptr = p;
  return TCL_OK;
}

extern "C" int tcl_peer_free(void* ptr) {
  if (!ptr) return TCL_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); set_last_error("tcl_peer_free: %s", cudaGetErrorString(e)); return TCL_ERR_CUDA; }
  return TCL_OK;
}

extern "C" int tcl_ipc_open(const void* handle, void** base) {
  TCL_CHECK_ARG(handle && base, "tcl_ipc_open: null argument");
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  const std::string key(reinterpret_cast<const char*>(handle), sizeof(cudaIpcMemHandle_t));
  auto it = g_ipc_open.find(key);
  if (it != g_ipc_open.end()) { *base = it->second; return TCL_OK; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("tcl_ipc_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    return TCL_ERR_CUDA;
  }
  g_ipc_open[key] = p;
  *base = p;
  return TCL_OK;
}

extern "C" int tcl_ipc_close(void* base) {
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  for (auto it = g_ipc_open.begin(); it != g_ipc_open.end(); ++it) {
    if (it->second != base) continue;
    cudaError_t e = cudaIpcCloseMemHandle(base);
    g_ipc_open.erase(it);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error("tcl_ipc_close: %s", cudaGetErrorString(e)); return TCL_ERR_CUDA; }
    return TCL_OK;
  }
  set_last_error("tcl_ipc_close: %p is not an open mapping", base);
  return TCL_ERR_ARG;
}

extern "C" int tcl_ipc_close_all(void) {
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  int rc = TCL_OK;
  for (auto& kv : g_ipc_open) {
    cudaError_t e = cudaIpcCloseMemHandle(kv.second);
    if (e != cudaSuccess) { cudaGetLastError(); set_last_error("tcl_ipc_close_all: %s", cudaGetErrorString(e)); rc = TCL_ERR_CUDA; }
  }
  g_ipc_open.clear();
  return rc;
}

extern "C" int tcl_peer_barrier(int32_t* const* flags, int world, int rank, int epoch, cudaStream_t stream) {
  TCL_CHECK_ARG(flags && world >= 1 && world <= TCL_MAX_RANKS && rank >= 0 && rank < world && epoch >= 1, "tcl_peer_barrier: args");
  PeerFlags pf;
  memset(&pf, 0, sizeof(pf));
  for (int r = 0; r < world; ++r) {
    TCL_CHECK_ARG(flags[r] != nullptr, "tcl_peer_barrier: null flag pointer (rank %d)", r);
    pf.f[r] = flags[r];
  }
  peer_barrier_kernel<<<1, 32, 0, stream>>>(pf, world, rank, epoch);
  TCL_CHECK_LAUNCH("tcl_peer_barrier");
  return TCL_OK;
}

extern "C" long long tcl_peer_barrier_timeouts(void) {
  unsigned long long v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_barrier_timeouts, sizeof(v));
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return (long long)v;
}
