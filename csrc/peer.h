#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace b2b {
int peer_alloc(size_t bytes, void** out);
int peer_free(void* p);
int ipc_export(void* p, char* handle64);
int ipc_import(const char* handle64, void** out);
int ipc_close(void* p);
int enable_peer_access(int dev, int peer);
int can_access_peer(int dev, int peer);
int memcpy_peer_async(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, cudaStream_t s);
int host_ring_alloc(size_t bytes, void** host_ptr, void** dev_ptr);
int host_ring_free(void* host_ptr);
}  // namespace b2b
