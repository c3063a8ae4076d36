// Host emulation of the CUDA execution model, for running festim_b200's kernel sources on a
// CPU in the test-suite (tests/emu/build_emu.py compiles csrc/*.cu with g++ against this
// header; kernel launches are rewritten to EMU_KERNEL(...)).  Test infrastructure only: the
// product library is the nvcc build and never includes this file.
//
// One OS thread per CUDA thread of a block, blocks run one after another.  __syncthreads,
// __syncwarp and the warp collectives are real barriers between those threads, so the
// warp-synchronous logic of the kernels (shared-memory staging, ballots, bit masks, gathers)
// executes as written.  Not emulated: cooperative grid launches, peer memory, CUDA graphs.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern dim3 blockDim, gridDim;

namespace emu {

// barrier whose participant count can shrink (threads that returned from the kernel)
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int expected = 0, waiting = 0;
  unsigned long long generation = 0;
  void reset(int n) { expected = n; waiting = 0; }
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned long long gen = generation;
    if (++waiting >= expected) { waiting = 0; ++generation; cv.notify_all(); return; }
    cv.wait(lk, [&] { return gen != generation; });
  }
  void drop() {   // a participant left for good
    std::unique_lock<std::mutex> lk(m);
    --expected;
    if (expected > 0 && waiting >= expected) { waiting = 0; ++generation; cv.notify_all(); }
  }
};

struct WarpCtx {
  Barrier bar[32];            // keyed by the lowest lane of the participating mask
  unsigned long long slot[32];
};

struct BlockCtx {
  Barrier block_bar;
  std::unique_ptr<WarpCtx[]> warps;
  std::vector<unsigned char> smem;
};

extern BlockCtx* g_block;
extern thread_local int t_lane, t_warp;

inline int popc(unsigned m) { return __builtin_popcount(m); }

// rendezvous of the lanes named by mask; every participant deposits v and reads all
template <typename F>
inline void warp_exchange(unsigned mask, unsigned long long v, F&& reader) {
  WarpCtx& w = g_block->warps[t_warp];
  const int key = __builtin_ctz(mask);
  Barrier& b = w.bar[key];
  {
    std::unique_lock<std::mutex> lk(b.m);
    if (b.expected != popc(mask) && b.waiting == 0) b.expected = popc(mask);
  }
  w.slot[t_lane] = v;
  b.arrive_and_wait();
  reader(w.slot);
  b.arrive_and_wait();
}

void run(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

template <typename K>
struct Launch {
  K k;
  dim3 g, b;
  size_t smem;
  template <typename... A>
  void operator()(A... args) {
    run(g, b, smem, [&] { k(args...); });
  }
};
template <typename K>
Launch<K> make_launch(K k, dim3 g, dim3 b, size_t smem = 0, cudaStream_t = nullptr) {
  return Launch<K>{k, g, b, smem};
}

}  // namespace emu

#define EMU_KERNEL(k, ...) emu::make_launch(k, __VA_ARGS__)
#define __launch_bounds__(...)
#define __shared__ static

inline void* emu_dynamic_smem() { return emu::g_block->smem.data(); }

// ---- synchronisation and warp collectives -----------------------------------------------
inline void __syncthreads() { emu::g_block->block_bar.arrive_and_wait(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  emu::warp_exchange(mask, 0, [](unsigned long long*) {});
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned r = 0;
  emu::warp_exchange(mask, pred ? 1 : 0, [&](unsigned long long* s) {
    for (int l = 0; l < 32; ++l)
      if ((mask >> l) & 1u) r |= (unsigned)(s[l] & 1ull) << l;
  });
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  unsigned r = 0;
  emu::warp_exchange(mask, v, [&](unsigned long long* s) {
    for (int l = 0; l < 32; ++l)
      if ((mask >> l) & 1u) r |= (unsigned)s[l];
  });
  return r;
}
inline int __reduce_max_sync(unsigned mask, int v) {
  int r = INT32_MIN;
  emu::warp_exchange(mask, (unsigned long long)(long long)v, [&](unsigned long long* s) {
    for (int l = 0; l < 32; ++l)
      if ((mask >> l) & 1u) r = std::max(r, (int)(long long)s[l]);
  });
  return r;
}
inline double __shfl_xor_sync(unsigned mask, double v, int lanemask) {
  unsigned long long bits;
  memcpy(&bits, &v, 8);
  double out = v;
  emu::warp_exchange(mask, bits, [&](unsigned long long* s) {
    const int src = emu::t_lane ^ lanemask;
    if ((mask >> src) & 1u) memcpy(&out, &s[src], 8);
  });
  return out;
}
inline double __shfl_sync(unsigned mask, double v, int src) {
  unsigned long long bits;
  memcpy(&bits, &v, 8);
  double out = v;
  emu::warp_exchange(mask, bits, [&](unsigned long long* s) { memcpy(&out, &s[src & 31], 8); });
  return out;
}
inline int __shfl_sync(unsigned mask, int v, int src) {
  int out = v;
  emu::warp_exchange(mask, (unsigned long long)(unsigned)v, [&](unsigned long long* s) { out = (int)s[src & 31]; });
  return out;
}
inline unsigned __shfl_sync(unsigned mask, unsigned v, int src) { return (unsigned)__shfl_sync(mask, (int)v, src); }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }

// ---- intrinsics ---------------------------------------------------------------------------
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline long long clock64() { return 0; }
using std::isinf;
using std::max;
using std::min;

extern std::mutex g_atomic_mutex;
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double* p, double v) {
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  const double old = *p;
  *p = old + v;
  return old;
}
