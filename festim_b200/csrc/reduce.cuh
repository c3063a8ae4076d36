// Warp-shuffle + ticket-counter grid reductions shared by the vector, Krylov
// and derived-quantity kernels.
#pragma once
#include "fb_internal.cuh"

#define FB_RED_SLOTS 72   // max simultaneous reductions (GMRES restart + 2)
#define FB_TPB 256

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction of NV values per thread.  Returns true in every
// thread of the LAST block to finish; there tot[] (thread 0) holds the totals.
template <int NV>
__device__ bool grid_reduce(double (&val)[NV], double* __restrict__ partials,
                            unsigned* __restrict__ ticket, double (&tot)[NV]) {
  __shared__ double sm[NV][FB_TPB / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(val[k]);
    if (lane == 0) sm[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int w = 0; w < nw; ++w) s += sm[k][w];
      partials[(size_t)k * gridDim.x + blockIdx.x] = s;
    }
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      s += partials[(size_t)k * gridDim.x + b];
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) sm[k][wid] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0.0;
      for (int w = 0; w < nw; ++w) r += sm[k][w];
      tot[k] = r;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
  return true;
}

