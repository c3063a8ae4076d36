// Cross-rank primitives over NVLink peer memory (one process per GPU, exchange areas mapped
// with CUDA IPC by fb_peer_alloc / fb_peer_open): self-validating LL words, the cross-rank
// completion of a grid reduction, and the owner -> ghost halo exchange of a node-major vector.
// They let the whole Newton-Krylov solve of a partitioned mesh run from the library's own stream
// with no NCCL call and no host round trip inside it: what the reference gets from PETSc's
// VecScatter / MPI_Allreduce under mpirun.
#pragma once
#include "fb_device.cuh"

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_add_release_u32(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// Self-validating 8-byte words {tag : 32 | half a double : 32} (the protocol NCCL calls LL):
// the writer needs no fence or flag, the reader spins until both halves carry the tag it
// expects.  Used for data another block (or, over NVLink, another GPU) consumes right after
// it is produced: it replaces a barrier by point-to-point waiting.
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned tag, double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long t = (unsigned long long)tag << 32;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(t | (bits & 0xffffffffull)) : "memory");
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p + 1), "l"(t | (bits >> 32)) : "memory");
}
__device__ __forceinline__ unsigned long long ll_word(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spins until both halves carry `tag`; false on timeout / abort (bar[64])
__device__ __forceinline__ bool ll_load(const unsigned long long* p, unsigned tag, unsigned* bar, double& out) {
  unsigned spins = 0;
  while (true) {
    const unsigned long long lo = ll_word(p), hi = ll_word(p + 1);
    if ((unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag) {
      out = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
      return true;
    }
    if ((++spins & 0x3ffu) == 0u) {
      __nanosleep(64);
      if (spins > (1u << 23) || ld_acquire_u32(bar + 64)) { atomicExch(bar + 64, 1u); out = 0.0; return false; }
    }
  }
}


// exchange area of a rank (u64 words): [0, 128) reduction slots of the persistent CG,
// [FB_XG, FB_XG + 2*8*FB_XG_K*2) reduction slots of the Krylov kernels
// ([parity][rank of origin][value][2 halves]), [FB_XH, ...) halo words, two per ghost dof and parity
#define FB_XG 128
#define FB_XG_K 40
#define FB_XH 4096

struct XRank {
  int n, rank;                   // n <= 1: single rank, nothing crosses NVLink
  unsigned seq;                  // tag of this reduction / exchange (same on every rank, never repeats)
  unsigned long long* base[8];   // exchange area of every rank (own one included)
  unsigned* bar;                 // bar[64]: abort flag (a peer never answered)
};

// Completes a grid reduction across ranks.  Call from every thread of the block that finished
// the local reduction (grid_reduce returned true); tot[] is valid in thread 0 on entry and holds
// the sums over all ranks, added in rank order (bit-identical on every rank), on return.
template <int NV>
__device__ void xrank_sum(const XRank& X, double (&tot)[NV]) {
  if (X.n <= 1) return;
  __shared__ double xr_s[NV];
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) xr_s[k] = tot[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    const int k = threadIdx.x;
    const unsigned par = X.seq & 1u;
    const double mine = xr_s[k];
    for (int q = 0; q < X.n; ++q)
      if (q != X.rank) ll_store(X.base[q] + FB_XG + ((size_t)(par * 8 + X.rank) * FB_XG_K + k) * 2, X.seq, mine);
    double s = 0.0;
    for (int q = 0; q < X.n; ++q) {
      double v = mine;
      if (q != X.rank) ll_load(X.base[X.rank] + FB_XG + ((size_t)(par * 8 + q) * FB_XG_K + k) * 2, X.seq, X.bar, v);
      s += v;
    }
    xr_s[k] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = xr_s[k];
}

// owner -> ghost: every owned vertex another rank holds as a ghost is written into that rank's
// halo words (tagged), then the own ghost entries are collected from the own halo words
__global__ void k_halo_push(XRank X, int ns, int n_send, const int* __restrict__ send_v, const int* __restrict__ send_rank,
                            const int* __restrict__ send_dst, long long halo_words, const double* __restrict__ x) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)n_send * ns) return;
  const int q = (int)(t / ns), s = (int)(t - (long long)q * ns);
  // [ghost dof][parity][2 halves]: the position does not depend on the receiver's ghost count
  unsigned long long* dst = X.base[send_rank[q]] + FB_XH + (((size_t)send_dst[q] * ns + s) * 2 + (X.seq & 1u)) * 2;
  ll_store(dst, X.seq, x[(long long)send_v[q] * ns + s]);
}
// both directions in one launch: the first n_send * ns threads push, every thread below
// n_ghost_dofs pulls (a pull only waits for remote pushes, never for the local ones)
__global__ void k_halo_exchange(XRank X, int ns, int n_send, const int* __restrict__ send_v,
                                const int* __restrict__ send_rank, const int* __restrict__ send_dst,
                                long long n_owned_dofs, long long n_ghost_dofs, double* __restrict__ x) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < (long long)n_send * ns) {
    const int q = (int)(t / ns), s = (int)(t - (long long)q * ns);
    unsigned long long* dst = X.base[send_rank[q]] + FB_XH + (((size_t)send_dst[q] * ns + s) * 2 + (X.seq & 1u)) * 2;
    ll_store(dst, X.seq, x[(long long)send_v[q] * ns + s]);
  }
  if (t < n_ghost_dofs) {
    double v = 0.0;
    ll_load(X.base[X.rank] + FB_XH + ((size_t)t * 2 + (X.seq & 1u)) * 2, X.seq, X.bar, v);
    x[n_owned_dofs + t] = v;
  }
}
__global__ void k_halo_pull(XRank X, long long n_owned_dofs, long long n_ghost_dofs, long long halo_words,
                            double* __restrict__ x) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_ghost_dofs) return;
  double v = 0.0;
  ll_load(X.base[X.rank] + FB_XH + ((size_t)t * 2 + (X.seq & 1u)) * 2, X.seq, X.bar, v);
  x[n_owned_dofs + t] = v;
}
