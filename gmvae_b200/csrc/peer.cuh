// peer.cuh -- the data-parallel exchange step as our own kernels over NVLink peer memory, fused with the optimiser:
// a two-shot all-reduce(sum) of the flat gradient buffer whose last phase IS the Adam kernel, in place of ncclAllReduce.
// On when the host attaches the ranks' symmetric regions (gmvae_peer_export / gmvae_peer_attach; GMVAE_DP_PEER=0 keeps NCCL).
//
// Why: at cfg4 the exchange is 8.4 MB.  The bandwidth term over NVSwitch is ~10 us per phase, but an in-stream ncclAllReduce
// costs 53 us (2 GPUs) to 100 us (8 GPUs) per 0.39 ms step, all exposed: the last weight gradient of the backward pass is the
// first thing Adam needs.  Here the gradient buffer of every rank lives in a symmetric region mapped by all ranks (cudaIpc):
//   exchange_kernel  rank j publishes "my gradients are final" (epoch flag), waits for every rank's flag, PULLS shard j of every
//                    rank's gradients over NVLink, adds them in rank order (deterministic: every replica gets the same bits) and
//                    PUSHES the sum into every rank's `red` buffer; its last block publishes "shard j has landed".
//   adam_kernel      (kernels.cuh) waits for the `world` landed-flags, reads the reduced gradients from its own `red` buffer and
//                    clears its own gradient buffer -- safe: a shard owner publishes only after it has pulled from everybody.
// One extra launch per step instead of a collective.  Flags are 64-bit epochs in the destination's memory, written with
// st.release.sys after a system fence and polled with ld.acquire.sys by one thread per block; peer memory is read with
// ld.global.cg.  The epoch lives in device memory, so a captured CUDA graph replays.  No buffer needs double buffering: a rank
// publishes epoch e+1 only after its own Adam of epoch e (stream order), and an owner pushes epoch e+1 sums into my `red` only
// after it has seen MY flag of epoch e+1.
#pragma once
#include "common.cuh"

namespace gmvae {
namespace peer {

constexpr int MAX_WORLD = 16;
constexpr int THREADS = 256;

// One symmetric region per rank (same layout everywhere), exported through cudaIpc.
struct Layout {
  int world;
  int64_t n4;        // float4 elements of the gradient buffer (parameters + loss accumulators)
  int64_t cap4;      // float4 elements per shard (shard j = [j cap4, min(n4, (j+1) cap4)))
  size_t grad_off, red_off, flags_off, local_off, bytes;
};
__host__ __device__ inline Layout make_layout(int world, int64_t n_floats) {
  Layout L;
  L.world = world;
  L.n4 = n_floats / 4;
  L.cap4 = (L.n4 + world - 1) / world;
  const size_t buf_bytes = ((size_t)L.n4 * 16 + 255) / 256 * 256;
  L.grad_off = 0;                                                  // [n4] float4: THIS rank's gradient buffer (what the backward pass accumulates into)
  L.red_off = buf_bytes;                                           // [n4] float4: the reduced gradients (shard j written by rank j)
  L.flags_off = 2 * buf_bytes;                                     // [2][MAX_WORLD] u64 epochs: row 0 "rank r's gradients are final", row 1 "shard r has landed"
  L.local_off = L.flags_off + 2 * MAX_WORLD * 8;                   // Local (never written by peers)
  L.bytes = L.local_off + 256;
  return L;
}
struct Local { unsigned long long epoch; unsigned int done[3]; unsigned int pad; };
struct Peers {           // mapped base pointers of every rank's region, own rank included
  float4* grad[MAX_WORLD]; float4* red[MAX_WORLD]; unsigned long long* flags[MAX_WORLD];
};

// ---- element arithmetic (host + device: the host build is what the CPU tests run) -------------------------------
__host__ __device__ inline int64_t shard_begin4(const Layout& L, int j) { return (int64_t)j * L.cap4 < L.n4 ? (int64_t)j * L.cap4 : L.n4; }
__host__ __device__ inline int64_t shard_len4(const Layout& L, int j) {
  const int64_t rest = L.n4 - (int64_t)j * L.cap4;
  return rest < 0 ? 0 : (rest < L.cap4 ? rest : L.cap4);
}
// memory another GPU wrote / owns: read through L2 (peer addresses bypass it), never from a stale L1 line
struct LoadPeer {
  __host__ __device__ float4 operator()(const float4* p) const {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
  }
};
// sum over the ranks, in rank order, of element i4 of the gradient buffers
template <class Load>
__host__ __device__ inline float4 reduce_ranks(const Layout& L, const Peers& P, int64_t i4, Load ld) {
  float4 s = ld(P.grad[0] + i4);
  for (int r = 1; r < L.world; ++r) {
    const float4 t = ld(P.grad[r] + i4);
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  return s;
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// One thread per block polls.  Bounded by `timeout_cycles` (host: GMVAE_PEER_TIMEOUT_S, default 120 s -- a rank that writes a
// checkpoint or a summary between two steps keeps the others waiting here, exactly as it would inside a collective): a lost
// peer ends as a trapped launch instead of a hung GPU.
__device__ __forceinline__ void wait_epochs(const unsigned long long* flags, int world, unsigned long long epoch, long long timeout_cycles) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < world; ++r)
      while (ld_acquire_sys(flags + r) < epoch) {
        __nanosleep(64);
        if (clock64() - t0 > timeout_cycles) __trap();
      }
  }
  __syncthreads();
}
// every thread has fenced its peer stores; the last block to arrive publishes
__device__ __forceinline__ bool last_block(unsigned int* done) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    last = atomicAdd(done, 1u) == gridDim.x - 1;
    if (last) { *done = 0; __threadfence_system(); }
  }
  __syncthreads();
  return last;
}

// volatile: the loads of a pass stay in program order ahead of the additions (all of them in flight together)
__device__ __forceinline__ float4 ld_peer_v4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// W = world size known at compile time (2, 4, 8: every rank's loads of a pass are in flight together -- a peer load is a ~2 us
// NVLink round trip, and a loop over the ranks would pay it `world` times) or 0 (any world size: one rank at a time).
template <int W>
__global__ void __launch_bounds__(THREADS) exchange_kernel(Layout L, int rank, Peers P, Local* loc, long long timeout_cycles) {
  const unsigned long long epoch = loc->epoch + 1;
  // (a) this kernel is stream-ordered after the backward pass: my gradients are final and visible
  if (blockIdx.x == 0 && threadIdx.x < L.world) { __threadfence_system(); st_release_sys(P.flags[threadIdx.x] + rank, epoch); }
  // (b) every rank's gradients are final
  wait_epochs(P.flags[rank], L.world, epoch, timeout_cycles);
  // (c) my shard: pull, add in rank order (the same bits on every replica), push to everybody
  const int64_t b4 = shard_begin4(L, rank), len4 = shard_len4(L, rank);
  if constexpr (W > 0) {
    // thread (e, r) of a block loads element e of rank r's gradients -- one NVLink round trip for all ranks -- and, after the
    // exchange through shared memory, adds the W values of its element in rank order and stores the sum to rank r's `red`
    constexpr int EPB = THREADS / W;                           // elements (float4) per block and sub-pass
    constexpr int U = W <= 2 ? 4 : 2;                          // sub-passes whose loads are in flight together
    __shared__ float4 sh[U][THREADS];
    const int e = threadIdx.x / W, r = threadIdx.x % W;
    for (int64_t i0 = (int64_t)blockIdx.x * EPB * U; i0 < len4; i0 += (int64_t)gridDim.x * EPB * U) {
      float4 t[U];
#pragma unroll
      for (int u = 0; u < U; ++u) t[u] = ld_peer_v4(P.grad[r] + b4 + min(i0 + u * EPB + e, len4 - 1));   // (clamped: the tail re-reads the last element)
#pragma unroll
      for (int u = 0; u < U; ++u) sh[u][e * W + r] = t[u];
      __syncthreads();
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * EPB + e;
        if (i < len4) {
          float4 s = sh[u][e * W];
#pragma unroll
          for (int q = 1; q < W; ++q) { const float4 v = sh[u][e * W + q]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
          P.red[r][b4 + i] = s;
        }
      }
      __syncthreads();
    }
  } else {
    constexpr int U = 4;                                       // float4 per thread, rank and pass
    for (int64_t i0 = (int64_t)blockIdx.x * THREADS * U + threadIdx.x; i0 < len4; i0 += (int64_t)gridDim.x * THREADS * U) {
      float4 s[U];
#pragma unroll
      for (int u = 0; u < U; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < L.world; ++r) {
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i0 + u * THREADS < len4) t[u] = __ldcg(P.grad[r] + b4 + i0 + u * THREADS);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i0 + u * THREADS < len4) {
            if (r == 0) s[u] = t[u];
            else { s[u].x += t[u].x; s[u].y += t[u].y; s[u].z += t[u].z; s[u].w += t[u].w; }
          }
      }
      for (int p = 0; p < L.world; ++p)
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i0 + u * THREADS < len4) P.red[p][b4 + i0 + u * THREADS] = s[u];
    }
  }
  // (d) shard `rank` has landed everywhere
  if (last_block(&loc->done[0]) && threadIdx.x < L.world) st_release_sys(P.flags[threadIdx.x] + MAX_WORLD + rank, epoch);
}
// float4 elements a block covers per pass
constexpr int exchange_block_items(int world) {
  return world == 2 ? (THREADS / 2) * 4 : (world == 4 || world == 8 || world == 16) ? (THREADS / world) * 2 : THREADS * 4;
}

// Stand-alone form of the last phase (gmvae_allreduce_grads without the fused Adam): reduced gradients -> my gradient buffer.
__global__ void __launch_bounds__(THREADS) gather_kernel(Layout L, int rank, Peers P, Local* loc, long long timeout_cycles) {
  const unsigned long long epoch = loc->epoch + 1;
  wait_epochs(P.flags[rank] + MAX_WORLD, L.world, epoch, timeout_cycles);
  const float4* red = P.red[rank]; float4* g = P.grad[rank];
  for (int64_t i4 = (int64_t)blockIdx.x * THREADS + threadIdx.x; i4 < L.n4; i4 += (int64_t)gridDim.x * THREADS) g[i4] = __ldcg(red + i4);
  if (last_block(&loc->done[1]) && threadIdx.x == 0) loc->epoch = epoch;
}
#endif  // __CUDACC__

}  // namespace peer
}  // namespace gmvae
