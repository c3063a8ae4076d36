// Runtime half of the CPU emulation (see cuda_emu.h): block scheduler and the slice of the
// CUDA runtime API the library calls.  Device memory is host memory.
#include "cuda_emu.h"

#include <cstdio>

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
dim3 blockDim, gridDim;
std::mutex g_atomic_mutex;

namespace emu {
BlockCtx* g_block = nullptr;
thread_local int t_lane = 0, t_warp = 0;

void run(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nt = (int)(block.x * block.y * block.z);
  BlockCtx ctx;
  ctx.warps.reset(new WarpCtx[(nt + 31) / 32]);
  ctx.smem.resize(smem + 64);
  g_block = &ctx;
  blockDim = block;
  gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        ctx.block_bar.reset(nt);
        std::vector<std::thread> threads;
        threads.reserve(nt);
        for (int t = 0; t < nt; ++t)
          threads.emplace_back([&, t] {
            threadIdx = uint3{(unsigned)t, 0, 0};
            blockIdx = uint3{bx, by, bz};
            t_lane = t & 31;
            t_warp = t >> 5;
            body();
            ctx.block_bar.drop();
          });
        for (auto& th : threads) th.join();
      }
  g_block = nullptr;
}
}  // namespace emu

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 8); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 8); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  switch (a) {
    case cudaDevAttrMultiProcessorCount: *v = 4; break;        // few SMs: small grids in the emulation
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 232448; break;
    case cudaDevAttrMaxSharedMemoryPerMultiprocessor: *v = 233472; break;
    default: *v = 0;
  }
  return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaLaunchCooperativeKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) {
  return cudaErrorNotSupported;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
cudaError_t cudaDeviceCanAccessPeer(int* c, int, int) { *c = 0; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaErrorNotSupported; }
}
