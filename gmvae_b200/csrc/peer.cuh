// peer.cuh -- the data-parallel exchange step as our own kernels over NVLink peer memory:
// a two-shot all-reduce(sum) of the flat gradient buffer (reduce-scatter by push, all-gather by push),
// in place of ncclAllReduce.  EXPERIMENTAL, OFF BY DEFAULT (GMVAE_DP_PEER=1 in the Python host): written
// at the end of round 1 after the GPU budget was spent -- the index arithmetic is checked on the host
// (tests/native/host_peer.cu), the kernels have not run on hardware yet.
//
// Why: at cfg4 the exchange is 8.4 MB.  The bandwidth term over NVSwitch is ~10 us per phase, but the
// NCCL call costs 70-100 us per step (0.397 -> 0.467 ms at N = 2, 0.497 ms at N = 8), all exposed because
// the last weight gradient of the backward pass is the first thing Adam needs.  Here every rank owns one
// shard of the buffer:
//   A  push    : rank r stores shard j of its gradients into rank j's receive slot r          (NVLink stores)
//   B  reduce  : rank j waits for its `world` slots, adds them in rank order (deterministic, so every
//                replica sees the same bits) and stores the sum into every rank's `red` buffer (NVLink stores)
//   C  gather  : every rank waits for the `world` reduced shards and copies them over its gradient buffer
// Flags are 64-bit epochs in the destination's memory, written with st.release.sys after a system fence by
// the last block of the writing kernel and polled with ld.acquire.sys by one thread per block; data written
// by a peer is read with ld.global.cg (L2 is the point of coherence for peer writes).  The epoch lives in
// device memory so a captured CUDA graph replays correctly.  Buffers need no double buffering: a rank can
// start phase A of step s+1 only after its phase C of step s, which needed every rank's phase B of step s.
#pragma once
#include "common.cuh"

namespace gmvae {
namespace peer {

constexpr int MAX_WORLD = 16;
constexpr int THREADS = 256;

// One symmetric region per rank (same layout everywhere), exported through cudaIpc.
struct Layout {
  int world;
  int64_t n4;        // float4 elements to reduce
  int64_t cap4;      // float4 elements per shard (shard j = [j cap4, min(n4, (j+1) cap4)))
  size_t recv_off, red_off, flags_off, local_off, bytes;
};
__host__ __device__ inline Layout make_layout(int world, int64_t n_floats) {
  Layout L;
  L.world = world;
  L.n4 = n_floats / 4;
  L.cap4 = (L.n4 + world - 1) / world;
  const size_t shard_bytes = (size_t)L.cap4 * 16;
  L.recv_off = 0;                                                  // [world][cap4] float4: slot r = rank r's contribution to MY shard
  L.red_off = ((size_t)world * shard_bytes + 255) / 256 * 256;     // [world][cap4] float4: the reduced buffer (shard j from rank j)
  L.flags_off = L.red_off + ((size_t)world * shard_bytes + 255) / 256 * 256;   // [2][MAX_WORLD] u64 epochs
  L.local_off = L.flags_off + 2 * MAX_WORLD * 8;                   // Local (never written by peers)
  L.bytes = L.local_off + 256;
  return L;
}
struct Local { unsigned long long epoch; unsigned int done[3]; unsigned int pad; };
struct Peers {           // mapped base pointers of every rank's region, own rank included
  float4* recv[MAX_WORLD]; float4* red[MAX_WORLD]; unsigned long long* flags[MAX_WORLD];
};

// ---- element arithmetic (host + device: the host build is what the CPU tests run) -------------------------------
__host__ __device__ inline int64_t shard_len4(const Layout& L, int j) {
  const int64_t rest = L.n4 - (int64_t)j * L.cap4;
  return rest < 0 ? 0 : (rest < L.cap4 ? rest : L.cap4);
}
// phase A: gradient element i4 of rank `rank` goes to rank `owner`, slot `rank`, position i4 - owner cap4
__host__ __device__ inline void push_target(const Layout& L, int rank, int64_t i4, int& owner, int64_t& dst4) {
  owner = (int)(i4 / L.cap4);
  dst4 = (int64_t)rank * L.cap4 + (i4 - (int64_t)owner * L.cap4);
}
// data written by a peer: read through L2 (the point of coherence for peer writes), never from a stale L1 line
struct LoadPeerWritten {
  __host__ __device__ float4 operator()(const float4* p) const {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
  }
};
// phase B: sum of the `world` slots of element i of my shard, in rank order
template <class Load>
__host__ __device__ inline float4 reduce_slots(const Layout& L, const float4* recv, int64_t i, Load ld) {
  float4 s = ld(recv + i);
  for (int r = 1; r < L.world; ++r) {
    const float4 t = ld(recv + (int64_t)r * L.cap4 + i);
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  return s;
}
// where rank j's reduced element i lands in every rank's `red` buffer == its index in the flat gradient
__host__ __device__ inline int64_t red_index(const Layout& L, int j, int64_t i) { return (int64_t)j * L.cap4 + i; }

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one thread per block polls; bounded, so a lost peer ends as a trapped launch instead of a hung GPU
__device__ __forceinline__ void wait_epochs(const unsigned long long* flags, int world, unsigned long long epoch) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < world; ++r)
      while (ld_acquire_sys(flags + r) < epoch) {
        __nanosleep(100);
        if (clock64() - t0 > 8000000000LL) __trap();
      }
  }
  __syncthreads();
}
// every thread has fenced its peer stores; the last block to arrive publishes `epoch` in every rank's flag row
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

__global__ void __launch_bounds__(THREADS) push_kernel(const float4* __restrict__ grads, Layout L, int rank, Peers P, Local* loc) {
  const unsigned long long epoch = loc->epoch + 1;
  for (int64_t i4 = (int64_t)blockIdx.x * THREADS + threadIdx.x; i4 < L.n4; i4 += (int64_t)gridDim.x * THREADS) {
    int owner; int64_t dst4;
    push_target(L, rank, i4, owner, dst4);
    P.recv[owner][dst4] = grads[i4];
  }
  if (last_block(&loc->done[0]) && threadIdx.x < L.world) st_release_sys(P.flags[threadIdx.x] + rank, epoch);
}

__global__ void __launch_bounds__(THREADS) reduce_kernel(Layout L, int rank, Peers P, Local* loc) {
  const unsigned long long epoch = loc->epoch + 1;
  wait_epochs(P.flags[rank], L.world, epoch);                      // every rank's slot of MY shard has arrived
  const int64_t len4 = shard_len4(L, rank);
  const float4* recv = P.recv[rank];
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < len4; i += (int64_t)gridDim.x * THREADS) {
    const float4 s = reduce_slots(L, recv, i, LoadPeerWritten());
    const int64_t d = red_index(L, rank, i);
    for (int p = 0; p < L.world; ++p) P.red[p][d] = s;
  }
  if (last_block(&loc->done[1]) && threadIdx.x < L.world) st_release_sys(P.flags[threadIdx.x] + MAX_WORLD + rank, epoch);
}

__global__ void __launch_bounds__(THREADS) gather_kernel(float4* __restrict__ grads, Layout L, int rank, Peers P, Local* loc) {
  const unsigned long long epoch = loc->epoch + 1;
  wait_epochs(P.flags[rank] + MAX_WORLD, L.world, epoch);          // every owner's reduced shard has arrived
  const float4* red = P.red[rank];
  for (int64_t i4 = (int64_t)blockIdx.x * THREADS + threadIdx.x; i4 < L.n4; i4 += (int64_t)gridDim.x * THREADS) grads[i4] = __ldcg(red + i4);
  if (last_block(&loc->done[2]) && threadIdx.x == 0) loc->epoch = epoch;
}
#endif  // __CUDACC__

}  // namespace peer
}  // namespace gmvae
