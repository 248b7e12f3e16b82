// gemm_chain.cuh -- a CHAIN of dependent GEMMs in ONE persistent tcgen05 kernel.
//
// At the reference's layer sizes (784-512-512, batch 16 384) one layer is 256 output tiles on 148
// SMs: 1.73 waves.  Launched one kernel per GEMM (gemm_tc.cuh), every layer pays the pipeline fill,
// the wave quantisation and the exposed epilogue of its last tile.  Here the GEMMs of a whole
// section of the step (e.g. decoder forward + Bernoulli likelihood + decoder backward: 9 GEMMs) are
// jobs of one launch.  The tiles of all jobs form one sequence, dealt round-robin to the persistent
// CTAs, so the tail of job j overlaps the head of job j+1 and the TMA->MMA->epilogue pipeline never
// drains in between.
//
// Data dependencies are per 128-row block: a job that consumes the rows another job of the chain
// produces waits (TMA producer thread, ld.acquire.gpu spin) on a counter the producing tiles'
// epilogue warps increment (st -> fence -> red.add) -- forward/dgrad jobs wait for ONE row block,
// weight-gradient jobs for the row blocks their batch slice covers.  Every tile depends only on
// tiles with a smaller index in the sequence and each CTA walks its tiles in increasing order with
// all CTAs co-resident (grid <= number of SMs, 1 CTA/SM), so the waits cannot deadlock.
//
// Everything the one-GEMM kernel fixes at compile time (tile width, operand majorness, epilogue
// functor) is a run-time field of the job here; the mechanics (TMA ring, single-thread MMA issue,
// double-buffered TMEM accumulator, 16 epilogue warps with smem-staged coalesced I/O) are the same.
#pragma once
#include "gemm_tc.cuh"

namespace gmvae {
namespace tc {

enum EpiKind : int { EK_NONE = -1, EK_STORE_BF16 = 0, EK_STORE_F32 = 1, EK_BCE = 2, EK_RELUMASK = 3, EK_ATOMIC = 4 };
template <class Epi> struct epi_kind { static constexpr int value = EK_NONE; };
template <> struct epi_kind<EpiStore<bf16, EPI_PLAIN>> { static constexpr int value = EK_STORE_BF16; };
template <> struct epi_kind<EpiStore<float, EPI_PLAIN>> { static constexpr int value = EK_STORE_F32; };
template <> struct epi_kind<EpiBCE<bf16>> { static constexpr int value = EK_BCE; };
template <> struct epi_kind<EpiReluMask<bf16, bf16>> { static constexpr int value = EK_RELUMASK; };
template <> struct epi_kind<EpiAtomicAdd> { static constexpr int value = EK_ATOMIC; };

constexpr int CHAIN_MAX_JOBS = 10;
constexpr int CHAIN_MAX_DEPS = 3;
constexpr int CHAIN_STAGES = 4;
constexpr int CHAIN_STAGE_BYTES = A_STAGE_BYTES + 256 * BLOCK_K * 2;   // room for the widest tile (48 KB)
constexpr int CHAIN_EPI_BYTES = 128;
constexpr int CHAIN_SMEM_BYTES = CHAIN_STAGES * CHAIN_STAGE_BYTES + 1024 + 256 + EPI_WARPS * PATCH_BYTES + 2 * 256 * 4 + 256 * 4;

// counters[base + row_block] >= target.  by_k = 0: the row block of the consumer's own tile;
// by_k = 1 (weight gradients: the contraction runs over the batch): every row block its k-range covers.
struct ChainDep { int base, target, by_k, nblocks; };

struct alignas(64) ChainJob {
  CUtensorMap a1, b1, a2, b2;
  int M, N, kb1, kb2, kb_per_split, num_splits;
  int block_n, a_mn, b_mn, kind;
  int tiles_n, tiles_mn, total_tiles, tile_base;
  int sig_base;                      // counters[sig_base + m_block] += 1 per epilogue warp per finished tile; -1: nobody waits
  int ndeps;
  int epi_dep;                       // index into deps of the job that wrote the epilogue's own operand (ReLU mask source), or -1
  ChainDep deps[CHAIN_MAX_DEPS];
  alignas(16) unsigned char epi[CHAIN_EPI_BYTES];
};
struct ChainParams {
  int njobs;
  int* counters;
  long long* trace;    // test hook: clock64 stamps of CTA `trace_cta`, 16 per processed tile (null in production)
  int trace_cta;
  ChainJob jobs[CHAIN_MAX_JOBS];
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// bounded like mbar_wait: a scheduling bug ends as a trapped launch, not as a hung GPU
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  if (ld_acquire_gpu(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(40);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ uint64_t make_smem_desc_rt(uint32_t smem_addr, int mn_major) {
  const uint64_t lbo = mn_major ? (uint64_t)(BLOCK_K * 128) : 0;
  const uint64_t sbo = 1024;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct ChainShared {
  uint8_t* smem; uint64_t* full_bar; uint64_t* empty_bar; uint64_t* tmem_full_bar; uint64_t* tmem_empty_bar;
  uint8_t* patches; float* sbias_all; float* scs_all; uint32_t tmem_base;
};

// The epilogue warps' share of one job.  `it` counts the tiles this CTA has processed since the
// start of the kernel (accumulator buffer = it & 1).
template <class Epi>
__device__ __forceinline__ void chain_epilogue_job(const ChainJob& J, int* counters, const ChainShared& S, int& it, int warp, int lane,
                                                   long long* trace, int jidx) {
  constexpr int CW = 16;
  Epi epi = *reinterpret_cast<const Epi*>(J.epi);
  const int G = gridDim.x;
  const int first = (((int)blockIdx.x - J.tile_base) % G + G) % G;
  if (first >= J.total_tiles) return;
  const int M = J.M, N = J.N, BN = J.block_n, tiles_n = J.tiles_n, tiles_mn = J.tiles_mn;
  const int nchunk = BN / CW;
  const int e = warp - 2, quad = warp & 3, half = e >> 2;
  const int et = (int)threadIdx.x - 64;
  float* const cs_dst = epi.colsum_dst();
  const float* const bias = epi.bias_ptr();
  float* const scs_all = S.scs_all;
  int cs_n0 = -1;
  if (cs_dst) { for (int i = et; i < 256; i += EPI_WARPS * 32) scs_all[i] = 0.f; }
  auto cs_flush = [&]() {
    for (int i = et; i < BN; i += EPI_WARPS * 32) {
      const float v = scs_all[i];
      if (v != 0.f && cs_n0 + i < N) atomicAdd(cs_dst + cs_n0 + i, v);
      scs_all[i] = 0.f;
    }
  };
  for (int l = first; l < J.total_tiles; l += G, ++it) {
    const int z = l / tiles_mn, mn = l - z * tiles_mn;
    const int mb = mn / tiles_n;
    const int m0 = mb * BLOCK_M, n0 = (mn - mb * tiles_n) * BN;
    const int as = it & 1; const uint32_t ap = (it >> 1) & 1;
    const int m = m0 + quad * 32 + lane;
    float* sbias = S.sbias_all + (it & 1) * 256;
    {
      const uint32_t sb = smem_addr(sbias);
      for (int i = et; i < BN; i += EPI_WARPS * 32) sts32f(sb + 4 * i, (bias && n0 + i < N) ? __ldg(bias + n0 + i) : 0.f);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (cs_dst && cs_n0 != n0) {
      if (cs_n0 >= 0) {
        cs_flush();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
      cs_n0 = n0;
    }
    EpiCtx ctx{S.patches + e * PATCH_BYTES, max(0, min(32, M - (m0 + quad * 32))), sbias, scs_all};
    const bool mvalid = m < M;
    constexpr int CSTEP = EPI_WARPS / 4;
    auto chunk_ok = [&](int ci) { return ci < nchunk && n0 + ci * CW < N; };
    const bool tr = trace && e == 0 && lane == 0 && it < 64;
    if (tr) { trace[16 * it + 6] = clock64(); trace[16 * it + 15] = jidx; trace[16 * it + 14] = l; }
    typename Epi::template Pre<CW> pre_cur, pre_next;
    // The epilogue's own operand (ReLU mask source) is fetched ahead of the accumulator, i.e. possibly
    // before the TMA producer has seen this tile's dependencies: the warp checks the operand's producer itself.
    if (J.epi_dep >= 0) {
      if (lane == 0) wait_counter(counters + J.deps[J.epi_dep].base + mb, J.deps[J.epi_dep].target);
      __syncwarp();
    }
    if (chunk_ok(half)) pre_cur = epi.template prefetch<CW>(m, n0 + half * CW, min(CW, N - (n0 + half * CW)), mvalid, ctx);
    if (tr) trace[16 * it + 7] = clock64();
    mbar_wait(&S.tmem_full_bar[as], ap);
    tc_fence_after();
    if (tr) trace[16 * it + 8] = clock64();
    const uint32_t taddr = S.tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256);
#pragma unroll 1
    for (int ci = half; chunk_ok(ci); ci += CSTEP) {
      const int n = n0 + ci * CW;
      const int nv = min(CW, N - n);
      uint32_t r[CW];
      ctx.sbias = sbias + ci * CW;
      ctx.scs = scs_all + ci * CW;
      tmem_ld16_issue(taddr + ci * CW, r);
      const int cn = ci + CSTEP;
      if (chunk_ok(cn)) pre_next = epi.template prefetch<CW>(m, n0 + cn * CW, min(CW, N - (n0 + cn * CW)), mvalid, ctx);
      tmem_ld16_wait(r);
      float v[CW];
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
      epi.template row<CW>(m, n, v, nv, mvalid, pre_cur, ctx);
      pre_cur = pre_next;
    }
    tc_fence_before();
    if (tr) trace[16 * it + 9] = clock64();
    if (J.sig_base >= 0) fence_proxy_async_global();   // these rows are read back through TMA (async proxy) by later jobs
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&S.tmem_empty_bar[as]);
      if (J.sig_base >= 0) {
        __threadfence();
        atomicAdd(counters + J.sig_base + mb, 1);
      }
    }
    if (tr) trace[16 * it + 10] = clock64();
  }
  epi.finish_warp();
  if (cs_dst && cs_n0 >= 0) {
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    cs_flush();
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
  }
}

__global__ void __launch_bounds__(NUM_THREADS2, 1) gemm_chain_kernel(const __grid_constant__ ChainParams p) {
  constexpr int STAGES = CHAIN_STAGES, STAGE_BYTES = CHAIN_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* patches = smem + STAGES * STAGE_BYTES + 256;
  float* sbias_all = reinterpret_cast<float*>(patches + EPI_WARPS * PATCH_BYTES);
  float* scs_all = sbias_all + 2 * 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, c = blockIdx.x;
  long long* const trace = (p.trace && c == p.trace_cta) ? p.trace : nullptr;

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < p.njobs; ++j) {
      tma_prefetch_desc(&p.jobs[j].a1);
      tma_prefetch_desc(&p.jobs[j].b1);
      if (p.jobs[j].kb2 > 0) { tma_prefetch_desc(&p.jobs[j].a2); tma_prefetch_desc(&p.jobs[j].b2); }
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], EPI_WARPS); }
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      int pit = 0;
      for (int j = 0; j < p.njobs; ++j) {
        const ChainJob& J = p.jobs[j];
        const int tiles_n = J.tiles_n, tiles_mn = J.tiles_mn, BN = J.block_n, kb1 = J.kb1, kb_total = J.kb1 + J.kb2;
        const uint32_t tx_bytes = (uint32_t)(A_STAGE_BYTES + BN * BLOCK_K * 2);
        const int first = ((c - J.tile_base) % G + G) % G;
        for (int l = first; l < J.total_tiles; l += G, ++pit) {
          const int z = l / tiles_mn, mn = l - z * tiles_mn;
          const int mb = mn / tiles_n;
          const int m0 = mb * BLOCK_M, n0 = (mn - mb * tiles_n) * BN;
          const int kb_begin = z * J.kb_per_split, kb_end = min(kb_total, kb_begin + J.kb_per_split);
          const bool tr = trace && pit < 64;
          if (tr) trace[16 * pit + 0] = clock64();
          if (J.ndeps > 0) {
            for (int d = 0; d < J.ndeps; ++d) {
              const ChainDep& D = J.deps[d];
              if (!D.by_k) {
                wait_counter(p.counters + D.base + mb, D.target);
              } else {
                const int lo = (kb_begin * BLOCK_K) / BLOCK_M, hi = min(D.nblocks - 1, (kb_end * BLOCK_K - 1) / BLOCK_M);
                for (int b = lo; b <= hi; ++b) wait_counter(p.counters + D.base + b, D.target);
              }
            }
            fence_proxy_async_global();
          }
          if (tr) trace[16 * pit + 1] = clock64();
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const bool seg2 = kb >= kb1;
            const CUtensorMap* ta = seg2 ? &J.a2 : &J.a1;
            const CUtensorMap* tb = seg2 ? &J.b2 : &J.b1;
            const int k_elem = (seg2 ? kb - kb1 : kb) * BLOCK_K;
            uint8_t* sa = smem + stage * STAGE_BYTES;
            uint8_t* sb = sa + A_STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            if (J.a_mn) {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d(ta, &full_bar[stage], sa + i * (BLOCK_K * 128), m0 + i * 64, k_elem);
            } else {
              tma_load_2d(ta, &full_bar[stage], sa, k_elem, m0);
            }
            if (J.b_mn) {
              for (int i = 0; i < BN / 64; ++i) tma_load_2d(tb, &full_bar[stage], sb + i * (BLOCK_K * 128), n0 + i * 64, k_elem);
            } else {
              tma_load_2d(tb, &full_bar[stage], sb, k_elem, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (tr) trace[16 * pit + 2] = clock64();
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (single thread) =====
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int j = 0; j < p.njobs; ++j) {
        const ChainJob& J = p.jobs[j];
        const int tiles_mn = J.tiles_mn, kb_total = J.kb1 + J.kb2;
        const int a_mn = J.a_mn, b_mn = J.b_mn;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                               ((uint32_t)(J.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
        const uint64_t a_step = a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        const uint64_t b_step = b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        const int first = ((c - J.tile_base) % G + G) % G;
        for (int l = first; l < J.total_tiles; l += G, ++it) {
          const int z = l / tiles_mn;
          const int kb_begin = z * J.kb_per_split, kb_end = min(kb_total, kb_begin + J.kb_per_split);
          const int as = it & 1; const uint32_t ap = (it >> 1) & 1;
          const bool tr = trace && it < 64;
          if (tr) trace[16 * it + 3] = clock64();
          mbar_wait(&tmem_empty_bar[as], ap ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (tr && kb == kb_begin) trace[16 * it + 4] = clock64();
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t adesc = make_smem_desc_rt(sa, a_mn);
            const uint64_t bdesc = make_smem_desc_rt(sa + A_STAGE_BYTES, b_mn);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_bf16(adesc + (uint64_t)k * a_step, bdesc + (uint64_t)k * b_step, tmem_d, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tmem_full_bar[as]);
          if (tr) trace[16 * it + 5] = clock64();
        }
      }
    }
  } else {
    // ===== epilogue warps =====
    ChainShared S{smem, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, patches, sbias_all, scs_all, tmem_base};
    int it = 0;
    for (int j = 0; j < p.njobs; ++j) {
      const ChainJob& J = p.jobs[j];
      switch (J.kind) {
        case EK_STORE_BF16: chain_epilogue_job<EpiStore<bf16, EPI_PLAIN>>(J, p.counters, S, it, warp, lane, trace, j); break;
        case EK_STORE_F32: chain_epilogue_job<EpiStore<float, EPI_PLAIN>>(J, p.counters, S, it, warp, lane, trace, j); break;
        case EK_BCE: chain_epilogue_job<EpiBCE<bf16>>(J, p.counters, S, it, warp, lane, trace, j); break;
        case EK_RELUMASK: chain_epilogue_job<EpiReluMask<bf16, bf16>>(J, p.counters, S, it, warp, lane, trace, j); break;
        case EK_ATOMIC: chain_epilogue_job<EpiAtomicAdd>(J, p.counters, S, it, warp, lane, trace, j); break;
        default: break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace gmvae
