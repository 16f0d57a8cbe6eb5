// NVLink topology + peer-memory runtime: the in-process replacement for the reference's
// WAN plumbing (DHT / STUN / WebSocket transport, bee2bee/p2p_runtime.py:343-431,
// bee2bee/dht.py, bee2bee/nat.py).  A "peer" is a GPU; a connection is a mapped pointer.
//
//  * symmetric buffers: cudaMalloc'd staging/flag memory whose CUDA IPC handles are
//    exchanged once at start-up (through torch.distributed) and opened by neighbours,
//  * same-process multi-GPU: cudaDeviceEnablePeerAccess,
//  * bulk transfers (weights / KV migration): cudaMemcpyPeerAsync on the copy engines,
//  * pinned + device-mapped host rings for the token stream back to the asyncio side.
#include "peer.h"

#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

namespace b2b {

static std::mutex g_mu;
static std::unordered_map<void*, size_t> g_allocs;     // local cudaMalloc'd buffers
static std::unordered_map<void*, int> g_opened;        // IPC-opened peer buffers

int peer_alloc(size_t bytes, void** out) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  std::lock_guard<std::mutex> g(g_mu);
  g_allocs[p] = bytes;
  *out = p;
  return 0;
}

int peer_free(void* p) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_allocs.find(p);
  if (it == g_allocs.end()) return -1;
  g_allocs.erase(it);
  return static_cast<int>(cudaFree(p));
}

int ipc_export(void* p, char* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &h, 64);
  return 0;
}

int ipc_import(const char* handle64, void** out) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return static_cast<int>(e);
  std::lock_guard<std::mutex> g(g_mu);
  g_opened[p] = 1;
  *out = p;
  return 0;
}

int ipc_close(void* p) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_opened.find(p);
  if (it == g_opened.end()) return -1;
  g_opened.erase(it);
  return static_cast<int>(cudaIpcCloseMemHandle(p));
}

int enable_peer_access(int dev, int peer) {
  int prev = 0;
  cudaGetDevice(&prev);
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, dev, peer);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (!can) return -1;
  cudaSetDevice(dev);
  e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  cudaSetDevice(prev);
  return static_cast<int>(e);
}

int can_access_peer(int dev, int peer) {
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, dev, peer) != cudaSuccess) { cudaGetLastError(); return 0; }
  return can;
}

int memcpy_peer_async(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, cudaStream_t s) {
  return static_cast<int>(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, s));
}

int host_ring_alloc(size_t bytes, void** host_ptr, void** dev_ptr) {
  void* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) return static_cast<int>(e);
  std::memset(h, 0, bytes);
  void* d = nullptr;
  e = cudaHostGetDevicePointer(&d, h, 0);
  if (e != cudaSuccess) return static_cast<int>(e);
  *host_ptr = h;
  *dev_ptr = d;
  return 0;
}

int host_ring_free(void* host_ptr) { return static_cast<int>(cudaFreeHost(host_ptr)); }

}  // namespace b2b
