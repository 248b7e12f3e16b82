// gemm_tc.cuh -- bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (sm_100a), hand-written:
// TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (one issuing
// thread, accumulator in TMEM) -> tcgen05.ld -> fused epilogue functor (epilogue.cuh).
//
//   C[M,N] (+)= A1[M,K1] * B1[K1,N] + A2[M,K2] * B2[K2,N]        (second segment optional)
//
// The second K segment accumulates into the same TMEM tile; it removes the reference's
// tf.concat([x, y]) (base.py:66): [x,y] @ W = x @ W[:D] + y @ W[D:].
// Operand "majorness" is a template parameter: K-major operands come from row-major
// [rows, K] tensors (activations in forward/dgrad, transposed weight copies); MN-major
// operands come from row-major [K, rows] tensors and are what the weight-gradient
// contraction dW = X^T dY needs (both X and dY are stored [batch, features], the reduction
// runs over the batch), so no transposed activation copies are ever written.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#pragma once
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "epilogue.cuh"

namespace gmvae {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must end as a trapped launch (an error the host sees), never as
// a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a converged warp (the lowest active one; the same lane on every call of a fully active warp).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint64_t adesc, uint64_t bdesc, uint32_t tmem_d, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// tcgen05.ld is asynchronous: the destination registers are valid only after tcgen05.wait::ld.
// The wait takes the registers as in/out operands so the compiler cannot move a use above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster (the two SMs of a TPC) run ONE 256-row MMA -------------------
// Each CTA holds its 128 rows of A, HALF of the B tile's columns and its 128 rows of the accumulator (own TMEM); the leader
// (cluster rank 0) issues the instruction for both.  What it buys here: a 256 x 256 x 64 step needs 32 KB of operands per SM
// instead of 48 KB -- the main loop of the chained kernel is bound by the ~72 GB/s each SM gets from L2, not by the tensor pipe.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's even (leader) CTA
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's copy of `bar` (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
// TMA load into THIS CTA's shared memory whose transaction bytes are counted on the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// the same, multicast: the box also lands at the same offset in every CTA of `mask` (cluster ranks), each destination's bytes
// counted on the barrier of ITS pair's leader (quad mode of the chained kernel: the A rows two pairs share)
__device__ __forceinline__ void tma_load_2d_pair_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint64_t adesc, uint64_t bdesc, uint32_t tmem_d, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` (same offset) in BOTH CTAs of the pair when every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------
// Shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"; same bit layout as
// cute::UMMA::SmemDescriptor): [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48)
// version=1, [61,64) layout (2 = 128-byte swizzle).
//   K-major, 128B swizzle: rows are 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major, 128B swizzle: slabs of 64 MN-elements x BLOCK_K k-rows; k-rows 128 B apart,
//   8-k-row groups 1024 B apart (SBO), slabs BLOCK_K*128 B apart (LBO).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint64_t lbo = MN_MAJOR ? (uint64_t)(BLOCK_K * 128) : 0;
  const uint64_t sbo = 1024;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// bytes to add to the start address per UMMA_K (=16 elements of K)
template <bool MN_MAJOR>
__device__ __forceinline__ constexpr uint32_t desc_k_step() { return MN_MAJOR ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4; }

// Instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6), a=bf16 [7,10), b=bf16
// [10,13), a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29).
template <int N, bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__host__ __device__ constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }
template <int BLOCK_N> __host__ __device__ constexpr int stage_bytes() { return A_STAGE_BYTES + BLOCK_N * BLOCK_K * 2; }
// as many ring stages as fit in ~196 KB (one persistent CTA per SM), at most 8
template <int BLOCK_N> __host__ __device__ constexpr int stages_for() {
  return (196 * 1024) / stage_bytes<BLOCK_N>() > 8 ? 8 : (196 * 1024) / stage_bytes<BLOCK_N>();
}
constexpr int EPI_WARPS = 16;                      // four warps per TMEM lane quadrant (the epilogue is latency-bound: TLP)
constexpr int PATCH_BYTES = 1024;                  // per-epilogue-warp transposition scratch: 32 rows x 16 bf16 (epilogue.cuh)
template <int BLOCK_N> __host__ __device__ constexpr int smem_bytes() {
  return stages_for<BLOCK_N>() * stage_bytes<BLOCK_N>() + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * PATCH_BYTES +
         2 * 256 * 4 /*double-buffered bias*/ + 256 * 4 /*column-sum accumulator*/;
}

constexpr int NUM_THREADS2 = 64 + EPI_WARPS * 32;  // + TMA producer warp + MMA warp

struct GemmMaps {
  CUtensorMap a1, b1, a2, b2;
};

// Persistent, warp-specialised: each CTA walks tiles t = blockIdx.x, +gridDim.x, ... of the
// (n fastest, then m, then k-split) tile space.  The accumulator is double-buffered in TMEM so
// the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
template <int BLOCK_N, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(NUM_THREADS2, 1)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, int M, int N, int kb1, int kb2, int kb_per_split, int num_splits, long long* trace, Epi epi_in) {
  constexpr int STAGES = stages_for<BLOCK_N>();
  constexpr int STAGE_BYTES = stage_bytes<BLOCK_N>();
  constexpr int ACC_COLS = tmem_cols_for(BLOCK_N);         // column stride between the two accumulators
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  constexpr int CW = 16;                               // columns per epilogue chunk
  constexpr int NCHUNK = BLOCK_N / CW;
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N constraint for M=128");
  static_assert(!B_MN || BLOCK_N % 64 == 0, "MN-major B is loaded in 64-column slabs");
  static_assert((BLOCK_N * BLOCK_K * 2) % 1024 == 0, "stage buffers must stay 1024-byte aligned");
  static_assert(TMEM_COLS <= 512, "two accumulators must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;        // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* patches = smem + STAGES * STAGE_BYTES + 256;
  float* sbias_all = reinterpret_cast<float*>(patches + EPI_WARPS * PATCH_BYTES);   // [2][256]
  float* scs_all = sbias_all + 2 * 256;                                              // [256] column sums of the current n-tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N, tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_mn = tiles_m * tiles_n;
  const int total_tiles = tiles_mn * num_splits;
  const int kb_total = kb1 + kb2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.b1);
    if (kb2 > 0) { tma_prefetch_desc(&maps.a2); tma_prefetch_desc(&maps.b2); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], EPI_WARPS); }
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on global memory written by it is read.
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the tiles (uniform control flow: loop state and addresses stay in uniform registers),
    // one elected lane issues.  A single thread inside `if (lane == 0)` compiled to ~130 dependent instructions per k-block (vote
    // loop around every UTMALDG) and paced the main loop (profiles/r2_ablation_loads.md).
    const bool tr = trace && blockIdx.x == 0 && lane == 0;
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int z = t / tiles_mn, mn = t - z * tiles_mn;
      const int m0 = (mn / tiles_n) * BLOCK_M, n0 = (mn % tiles_n) * BLOCK_N;
      const int kb_begin = z * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
      if (tr) trace[16 * it + 0] = clock64();
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const bool seg2 = kb >= kb1;
        const CUtensorMap* ta = seg2 ? &maps.a2 : &maps.a1;
        const CUtensorMap* tb = seg2 ? &maps.b2 : &maps.b1;
        const int k_elem = (seg2 ? kb - kb1 : kb) * BLOCK_K;
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        uint64_t* const fb = &full_bar[stage];
        if (elect_one()) {
          mbar_expect_tx(fb, STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d(ta, fb, sa + i * (BLOCK_K * 128), m0 + i * 64, k_elem);
          } else {
            tma_load_2d(ta, fb, sa, k_elem, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i) tma_load_2d(tb, fb, sb + i * (BLOCK_K * 128), n0 + i * 64, k_elem);
          } else {
            tma_load_2d(tb, fb, sb, k_elem, n0);
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (tr) trace[16 * it + 1] = clock64();
    }
  } else if (warp == 1) {
    // ===== MMA issuer (the warp walks the tiles uniformly; one elected lane -- the same on every call -- issues and commits) =====
    constexpr uint32_t idesc = make_idesc<BLOCK_N, A_MN, B_MN>();
    const bool tr = trace && blockIdx.x == 0 && lane == 0;
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int z = t / tiles_mn;
      const int kb_begin = z * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
      const int as = it & 1; const uint32_t ap = (it >> 1) & 1;
      if (tr) trace[16 * it + 2] = clock64();
      mbar_wait(&tmem_empty_bar[as], ap ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      if (tr) trace[16 * it + 3] = clock64();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * ACC_COLS);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (tr && kb == kb_begin) trace[16 * it + 4] = clock64();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc<A_MN>(sa);
        const uint64_t bdesc = make_smem_desc<B_MN>(sa + A_STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            umma_bf16(adesc + (uint64_t)(k * desc_k_step<A_MN>()), bdesc + (uint64_t)(k * desc_k_step<B_MN>()), tmem_d, idesc,
                      (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[as]);   // accumulator complete
      if (tr) trace[16 * it + 5] = clock64();
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> fused epilogue -> global =====
    Epi epi = epi_in;
    const int e = warp - 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = e >> 2;             // the EPI_WARPS/4 warps of a quadrant interleave column chunks
    // Fused column sums (bias gradients): accumulated per CTA in shared memory and flushed with one
    // global atomic per column when the CTA moves to another n-tile and at the end.
    float* const cs_dst = epi.colsum_dst();
    const int et = (int)threadIdx.x - 64;               // index within the epilogue threads
    int cs_n0 = -1;
    if (cs_dst) { for (int i = et; i < 256; i += EPI_WARPS * 32) scs_all[i] = 0.f; }
    auto cs_flush = [&]() {                               // caller guarantees all adds are done (named barrier)
      for (int i = et; i < BLOCK_N; i += EPI_WARPS * 32) {
        const float v = scs_all[i];
        if (v != 0.f && cs_n0 + i < N) atomicAdd(cs_dst + cs_n0 + i, v);
        scs_all[i] = 0.f;
      }
    };
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int z = t / tiles_mn, mn = t - z * tiles_mn;
      const int m0 = (mn / tiles_n) * BLOCK_M, n0 = (mn % tiles_n) * BLOCK_N;
      const int as = it & 1; const uint32_t ap = (it >> 1) & 1;
      const int m = m0 + quad * 32 + lane;
      // stage this tile's bias while the MMA is still running; double-buffered, one named barrier per tile
      float* sbias = sbias_all + (it & 1) * 256;
      epi.template tile_begin<BLOCK_N>(n0, N, sbias, (int)threadIdx.x - 64, EPI_WARPS * 32);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (cs_dst && cs_n0 != n0) {                        // uniform over the epilogue warps
        if (cs_n0 >= 0) {
          cs_flush();
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        }
        cs_n0 = n0;
      }
      EpiCtx ctx{patches + e * PATCH_BYTES, max(0, min(32, M - (m0 + quad * 32))), sbias, scs_all};
      const bool tr = trace && blockIdx.x == 0 && e == 0 && lane == 0;
      if (tr) trace[16 * it + 6] = clock64();
      const bool mvalid = m < M;
      constexpr int CSTEP = EPI_WARPS / 4;
      auto chunk_ok = [&](int ci) { return ci < NCHUNK && n0 + ci * CW < N; };   // warp-uniform
      // The epilogue's own global loads (ReLU mask, image bytes) run one chunk ahead: the first
      // chunk's are issued before the accumulator is even complete, the next chunk's while the
      // current one is processed.
      typename Epi::template Pre<CW> pre_cur, pre_next;
      if (chunk_ok(half)) pre_cur = epi.template prefetch<CW>(m, n0 + half * CW, min(CW, N - (n0 + half * CW)), mvalid, ctx);
      mbar_wait(&tmem_full_bar[as], ap);
      tc_fence_after();
      if (tr) trace[16 * it + 7] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * ACC_COLS);
#pragma unroll 1
      for (int ci = half; chunk_ok(ci); ci += CSTEP) {
        const int n = n0 + ci * CW;
        const int nv = min(CW, N - n);
        uint32_t r[CW];
        ctx.sbias = sbias + ci * CW;
        ctx.scs = scs_all + ci * CW;
        if constexpr (CW == 32) tmem_ld32_issue(taddr + ci * CW, r); else tmem_ld16_issue(taddr + ci * CW, r);
        const int cn = ci + CSTEP;
        if (chunk_ok(cn)) pre_next = epi.template prefetch<CW>(m, n0 + cn * CW, min(CW, N - (n0 + cn * CW)), mvalid, ctx);
        if constexpr (CW == 32) tmem_ld32_wait(r); else tmem_ld16_wait(r);
        if (tr && ci / CSTEP < 3) trace[16 * it + 13 + ci / CSTEP] = clock64();
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
        epi.template row<CW>(m, n, v, nv, mvalid, pre_cur, ctx);
        if (tr && ci / CSTEP < 4) trace[16 * it + 8 + ci / CSTEP] = clock64();
        pre_cur = pre_next;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
      if (tr) trace[16 * it + 12] = clock64();
    }
    // once per warp and kernel (not per tile): the scalar loss accumulator sees 148 x 16 atomics
    // instead of one per warp and tile, which keeps its fp32 rounding error at the 1e-6 level
    epi.finish_warp();
    if (cs_dst && cs_n0 >= 0) {
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      cs_flush();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

// 2-D bf16 tensor, row-major [outer, inner] with `outer_stride` elements between rows;
// box = {box_inner (<= 64), box_outer}, 128-byte swizzle, out-of-bounds elements read as zero.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t outer_stride,
                   uint32_t box_inner, uint32_t box_outer);
// general form: element size 1/2/4 bytes, swizzle span 128/64/0 bytes
int make_tmap(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t outer_stride,
              uint32_t box_inner, uint32_t box_outer, int swizzle);

// Operand description given to launch_gemm_tc: a row-major bf16 matrix.
//   K-major operand  : stored [rows(M or N), K],  ld = row stride
//   MN-major operand : stored [K, rows(M or N)],  ld = row stride
struct Operand {
  const bf16* ptr; int64_t ld; int rows; int k;
};

int num_sms();
int current_device();
extern int g_reserved_sms;
extern long long* g_trace;   // device buffer of clock64 stamps of CTA 0 (test hook), normally null

template <int BLOCK_N, bool A_MN, bool B_MN, class Epi>
int launch_gemm_tc(const Operand& A1, const Operand& B1, const Operand* A2, const Operand* B2, int M, int N, int split_k,
                   const Epi& epi, cudaStream_t st, bool persistent = true) {
  static bool attr_set_by_dev[64] = {};              // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_by_dev[current_device()];
  auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, Epi>;
  if (!attr_set) {
    GM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BLOCK_N>()));
    attr_set = true;
  }
  GemmMaps maps;
  auto mk = [&](CUtensorMap* m, const Operand& o, bool mn, int box_rows) -> int {
    if (mn) return make_tmap_bf16(m, o.ptr, (uint64_t)o.rows, (uint64_t)o.k, (uint64_t)o.ld, 64, BLOCK_K);
    return make_tmap_bf16(m, o.ptr, (uint64_t)o.k, (uint64_t)o.rows, (uint64_t)o.ld, BLOCK_K, (uint32_t)box_rows);
  };
  GM_TRY(mk(&maps.a1, A1, A_MN, BLOCK_M));
  GM_TRY(mk(&maps.b1, B1, B_MN, BLOCK_N));
  int kb1 = (A1.k + BLOCK_K - 1) / BLOCK_K, kb2 = 0;
  if (A2 && B2) {
    GM_TRY(mk(&maps.a2, *A2, A_MN, BLOCK_M));
    GM_TRY(mk(&maps.b2, *B2, B_MN, BLOCK_N));
    kb2 = (A2->k + BLOCK_K - 1) / BLOCK_K;
  } else {
    maps.a2 = maps.a1; maps.b2 = maps.b1;
  }
  int kb_total = kb1 + kb2;
  if (split_k < 1) split_k = 1;
  int per = (kb_total + split_k - 1) / split_k;
  split_k = (kb_total + per - 1) / per;
  const int total = ((N + BLOCK_N - 1) / BLOCK_N) * ((M + BLOCK_M - 1) / BLOCK_M) * split_k;
  const int grid = persistent ? std::min(total, num_sms()) : total;
  GM_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(NUM_THREADS2), (size_t)smem_bytes<BLOCK_N>(), st, true, maps, M, N, kb1, kb2, per, split_k,
                         g_trace, epi));
  return 0;
}

}  // namespace tc
}  // namespace gmvae
