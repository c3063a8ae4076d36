// Micro-benchmark: FP64 FMA issue rate per SM (and FP32 for comparison), B200.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void k_fma(T* out, int iters, T a, T b) {
  T x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
void run(const char* name, int blocks, int threads) {
  T* out;
  cudaMalloc(&out, sizeof(T) * blocks * threads);
  const int iters = 4096;
  constexpr int ILP = 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fma<T, ILP><<<blocks, threads>>>(out, iters, (T)1.000001, (T)1e-9);
  cudaEventRecord(e0);
  k_fma<T, ILP><<<blocks, threads>>>(out, iters, (T)1.000001, (T)1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * threads * iters * ILP;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s blocks=%d threads=%d: %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM (at %d MHz nominal)\n", name, blocks,
         threads, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / (148.0 * clk * 1e3) * (148.0 / (blocks < 148 ? blocks : 148)),
         clk / 1000);
  cudaFree(out);
}

int main() {
  run<double>("fp64", 148 * 2, 1024);
  run<double>("fp64", 1, 1024);
  run<double>("fp64", 1, 128);
  run<float>("fp32", 148 * 2, 1024);
  run<float>("fp32", 1, 1024);
  return 0;
}
