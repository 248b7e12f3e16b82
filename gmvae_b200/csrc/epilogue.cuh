// epilogue.cuh -- GEMM epilogue functors shared by the tcgen05 (bf16) and SIMT (fp32) kernels.
//
// A GEMM kernel hands an epilogue one row fragment at a time:
//     epi.row<NV>(m, n0, acc, nvalid)   acc[0..NV) = C[m, n0..n0+NV), the first `nvalid` are in range
// and calls epi.finish_warp() once per warp (all 32 lanes converged) when the tile is done.
// The ELBO body lives here: bias+ReLU (base.py MLP layers), the Bernoulli log-likelihood and its
// gradient (gmvae.py:254 / vae.py:177), ReLU masks of the backward pass, atomic accumulation of
// split-K weight gradients.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace gmvae {

// ---- row-fragment stores -------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void store_frag(float* dst, const float* v, int nvalid) {
  if (nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}
template <int NV>
__device__ __forceinline__ void store_frag(bf16* dst, const float* v, int nvalid) {
  if (NV % 8 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      uint4 p;
      p.x = pack_bf16x2(v[i], v[i + 1]); p.y = pack_bf16x2(v[i + 2], v[i + 3]);
      p.z = pack_bf16x2(v[i + 4], v[i + 5]); p.w = pack_bf16x2(v[i + 6], v[i + 7]);
      *reinterpret_cast<uint4*>(dst + i) = p;
    }
  } else if (nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 2) *reinterpret_cast<uint32_t*>(dst + i) = pack_bf16x2(v[i], v[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}

// ---- row-fragment loads (converted to fp32) ------------------------------------------------
template <int NV>
__device__ __forceinline__ void load_frag(const float* src, float* v, int nvalid) {
  if (nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(src + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < nvalid ? src[i] : 0.f;
  }
}
template <int NV>
__device__ __forceinline__ void load_frag(const bf16* src, float* v, int nvalid) {
  if (NV % 8 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      uint4 t = *reinterpret_cast<const uint4*>(src + i);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(p[j]);
        v[i + 2 * j] = f.x; v[i + 2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < nvalid ? __bfloat162float(src[i]) : 0.f;
  }
}
template <int NV>
__device__ __forceinline__ void load_frag(const uint8_t* src, float* v, int nvalid) {
  if (NV % 16 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 16) {
      uint4 t = *reinterpret_cast<const uint4*>(src + i);
      const uint8_t* p = reinterpret_cast<const uint8_t*>(&t);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[i + j] = (float)p[j];
    }
  } else if (NV % 4 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      uint32_t t = *reinterpret_cast<const uint32_t*>(src + i);
      v[i] = (float)(t & 0xff); v[i + 1] = (float)((t >> 8) & 0xff);
      v[i + 2] = (float)((t >> 16) & 0xff); v[i + 3] = (float)(t >> 24);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < nvalid ? (float)src[i] : 0.f;
  }
}

// ---- per-column sum over the 32 rows a warp holds (lane = row): recursive halving, 31 shuffles.
// On return lane l holds the sum of column (l % NV) ... for NV == 32 exactly column l; for
// NV == 16 lanes l and l+16 both hold column l % 16 partial sums of their half-warps.
template <int NV>
__device__ __forceinline__ float warp_colsum(const float* v, int lane) {
  float a[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = v[i];
  // step s halves the number of live columns per lane: keep the half selected by lane bit
  if constexpr (NV >= 32) {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float send = up ? a[i] : a[i + 16];
      float recv = __shfl_xor_sync(0xffffffffu, send, 16);
      a[i] = (up ? a[i + 16] : a[i]) + recv;
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = up ? a[i] : a[i + 8];
      float recv = __shfl_xor_sync(0xffffffffu, send, 8);
      a[i] = (up ? a[i + 8] : a[i]) + recv;
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = up ? a[i] : a[i + 4];
      float recv = __shfl_xor_sync(0xffffffffu, send, 4);
      a[i] = (up ? a[i + 4] : a[i]) + recv;
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = up ? a[i] : a[i + 2];
      float recv = __shfl_xor_sync(0xffffffffu, send, 2);
      a[i] = (up ? a[i + 2] : a[i]) + recv;
    }
  }
  {
    const bool up = lane & 1;
    float send = up ? a[0] : a[1];
    float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    a[0] = (up ? a[1] : a[0]) + recv;
  }
  if (NV < 32) a[0] += __shfl_xor_sync(0xffffffffu, a[0], 16);   // both half-warps hold 16 columns
  return a[0];
}
// column index (within the NV-wide chunk) whose sum lane `lane` holds after warp_colsum
template <int NV>
__device__ __forceinline__ int warp_colsum_index(int lane) {
  // lane bit 16 chose columns [16,32), bit 8 the upper 8 of those, ... : the index equals the lane
  // bits themselves for NV == 32; for NV == 16 the low four bits.
  return NV >= 32 ? (lane & 31) : (lane & 15);
}
template <int NV>
__device__ __forceinline__ void warp_colsum_atomic(float* dst, int n0, int nvalid, const float* v) {
  if constexpr (NV == 16 || NV == 32) {   // only the tensor-core epilogue (one row per lane) fuses column sums
    const int lane = threadIdx.x & 31;
    float s = warp_colsum<NV>(v, lane);
    int c = warp_colsum_index<NV>(lane);
    if (c < nvalid && (NV >= 32 || lane < 16) && s != 0.f) atomicAdd(dst + n0 + c, s);
  }
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- coalesced epilogue I/O for the tensor-core kernel ---------------------------------------
// In the tcgen05 epilogue lane l of a warp owns row (m_base + l) of a 32-row x NV-column chunk.
// Storing that row by row makes every store instruction touch 32 different cache lines with a
// 16-byte piece each (measured: ~2300 cycles per chunk).  Instead the warp transposes the chunk
// through a private 2 KB shared-memory patch (XOR-swizzled, conflict-free both ways) so that one
// global instruction covers 8 rows x 64 contiguous bytes (NV = 32) or 16 rows x 32 bytes (NV = 16).
// Explicit shared-space accesses: through a generic pointer the compiler emits LD.E/ST.E (generic
// loads tracked on the long scoreboard) instead of LDS/STS.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}

struct EpiCtx {
  uint8_t* patch;   // per-warp scratch (>= 2 KB) or null (CUDA-core GEMM: rows are stored directly)
  int rows_valid;   // valid rows of the warp's 32-row group (lane l <-> row m_base + l)
  const float* sbias;  // tensor-core kernel: this chunk's bias values staged in shared memory by tile_begin()
  float* scs;          // tensor-core kernel: this chunk's slice of the CTA-wide column-sum accumulator (shared memory)
};
// Chunks are moved in 16-byte pieces; a ragged last chunk (N not a multiple of NV) rounds its
// width up to 8 columns, which stays inside the zero padding of the row (strides are multiples of 8).
__device__ __forceinline__ int cols8(int nvalid) { return (nvalid + 7) & ~7; }
template <int NV> __device__ __forceinline__ int swz(int r) { return NV == 32 ? ((r >> 1) & 3) : 0; }

// all 32 lanes call; lane l passes its row's NV values.  rows_valid = number of valid rows of the chunk.
// v[i] must be 0 for i >= nvalid (the rounded-up columns land in the row padding, which stays zero).
template <int NV>
__device__ __forceinline__ void store_chunk_bf16(bf16* base /*row m_base, col n0*/, int64_t ld, const float* v, int nvalid,
                                                 const EpiCtx& ctx, bool keep_patch = false) {
  constexpr int U = NV / 8;                         // 16-byte units per row
  const int lane = threadIdx.x & 31;
  const int rows_valid = ctx.rows_valid, ncols = cols8(nvalid);
  const bool fast = (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
  const uint32_t patch = smem_addr(ctx.patch);
  if (!fast) {                                      // unaligned view: plain per-row store (patch still filled if kept)
    if (lane < rows_valid) store_frag<NV>(base + (int64_t)lane * ld, v, nvalid);
    if (!keep_patch) return;
  }
#pragma unroll
  for (int c = 0; c < U; ++c) {
    uint4 p;
    p.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]); p.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
    p.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]); p.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    sts128(patch + 16 * (lane * U + (c ^ swz<NV>(lane))), p);
  }
  __syncwarp();
  constexpr int ROWS_PER_IT = 32 / U;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int r = ROWS_PER_IT * j + lane / U, c = lane % U;
    const uint4 val = lds128(patch + 16 * (r * U + (c ^ swz<NV>(r))));
    if (fast && r < rows_valid && 8 * c < ncols) *reinterpret_cast<uint4*>(base + (int64_t)r * ld + 8 * c) = val;
  }
  if (!keep_patch) __syncwarp();
}

// Column sums of the chunk just staged in the patch by store_chunk_bf16 (call with keep_patch = true
// there, then this, which ends with the releasing __syncwarp): lanes 0..15 / 16..31 each walk 16 rows
// of a 32-bit column pair, one shuffle merges the halves, NV/2 lanes issue two atomics each.
// The sums go into the CTA-wide shared accumulator ctx.scs (flushed to global memory by the kernel
// once per CTA and n-tile): per-warp global atomics would put M/32 same-address atomics on every
// column, which L2 serialises.
template <int NV>
__device__ __forceinline__ void colsum_from_patch(int nvalid, const EpiCtx& ctx) {
  constexpr int U = NV / 8, WORDS = NV / 2;              // 32-bit words (column pairs) per row
  const int lane = threadIdx.x & 31;
  const uint32_t patch = smem_addr(ctx.patch);
  constexpr int LANES_PER_HALF = WORDS;                   // 16 (NV=32) or 8 (NV=16)
  constexpr int GROUPS = 32 / LANES_PER_HALF;              // 2 or 4 row groups
  constexpr int ROWS_PER_GROUP = 32 / GROUPS;
  const int w = lane % LANES_PER_HALF, grp = lane / LANES_PER_HALF;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < ROWS_PER_GROUP; ++i) {
    const int r = grp * ROWS_PER_GROUP + i;
    const int c = w / 4;                                   // 16-byte unit holding this word
    const uint32_t t = lds32(patch + 4 * ((r * U + (c ^ swz<NV>(r))) * 4 + (w & 3)));
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t));
    s0 += f.x; s1 += f.y;
  }
#pragma unroll
  for (int o = LANES_PER_HALF; o < 32; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if (grp == 0) {
    const uint32_t a = smem_addr(ctx.scs) + 8 * w;
    if (2 * w < nvalid && s0 != 0.f) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(s0) : "memory");
    if (2 * w + 1 < nvalid && s1 != 0.f) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 4), "f"(s1) : "memory");
  }
  __syncwarp();
}

// Coalesced load of a 32-row x NV-column bf16 chunk: phase 1 (issue) returns the raw 16-byte
// pieces this lane fetched; phase 2 exchanges them through the patch and returns the lane's own row.
template <int NV> struct ChunkRegs { uint4 q[NV / 8]; };
template <int NV>
__device__ __forceinline__ ChunkRegs<NV> load_chunk_bf16_issue(const bf16* base, int64_t ld, int nvalid, const EpiCtx& ctx) {
  constexpr int U = NV / 8, ROWS_PER_IT = 32 / U;
  const int lane = threadIdx.x & 31;
  const int ncols = cols8(nvalid);
  ChunkRegs<NV> g;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int r = ROWS_PER_IT * j + lane / U, c = lane % U;
    g.q[j] = (r < ctx.rows_valid && 8 * c < ncols) ? *reinterpret_cast<const uint4*>(base + (int64_t)r * ld + 8 * c)
                                                   : make_uint4(0, 0, 0, 0);
  }
  return g;
}
template <int NV>
__device__ __forceinline__ void load_chunk_bf16_finish(const ChunkRegs<NV>& g, float* v, const EpiCtx& ctx) {
  constexpr int U = NV / 8, ROWS_PER_IT = 32 / U;
  const int lane = threadIdx.x & 31;
  const uint32_t patch = smem_addr(ctx.patch);
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int r = ROWS_PER_IT * j + lane / U, c = lane % U;
    sts128(patch + 16 * (r * U + (c ^ swz<NV>(r))), g.q[j]);
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < U; ++c) {
    const uint4 t = lds128(patch + 16 * (lane * U + (c ^ swz<NV>(lane))));
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(p[k]);
      v[8 * c + 2 * k] = f.x; v[8 * c + 2 * k + 1] = f.y;
    }
  }
  __syncwarp();
}

// Every functor has a two-phase interface so that a kernel can issue the epilogue's own global
// loads (bias, ReLU mask, image bytes) BEFORE it waits for the accumulator:
//     auto pre = epi.template prefetch<NV>(m, n0, nvalid, valid, ctx);
//     ... wait for acc ...
//     epi.template row<NV>(m, n0, acc, nvalid, valid, pre, ctx);
// `valid` is false for rows beyond M.  With ctx.patch != null (tensor-core kernel) both calls are
// warp-collective: all 32 lanes call them with consecutive rows m = m_base + lane.
// tile_begin(n0, N, sbias, t, nt): called by the nt epilogue threads (t = 0..nt-1) of the tensor-core
// kernel before they wait for a tile's accumulator; stages the tile's bias values in shared memory.
template <int BN>
__device__ __forceinline__ void stage_bias(const float* bias, int n0, int N, float* sbias, int t, int nt) {
  const uint32_t sb = smem_addr(sbias);
  for (int i = t; i < BN; i += nt) sts32f(sb + 4 * i, (bias && n0 + i < N) ? __ldg(bias + n0 + i) : 0.f);
}

template <typename T, int NV> struct staged_io { static constexpr bool value = std::is_same<T, bf16>::value && (NV == 16 || NV == 32); };

// ---- out = act(acc + bias[n] (+ addend[m,n]))  (out += ... when ACCUM) -------------------------
// Forward MLP layers (base.py:46-60: relu(h W + b), last layer linear) and plain stores.
enum { EPI_PLAIN = 0, EPI_ADDEND = 1, EPI_ACCUM = 2 };
template <typename OutT, int MODE = EPI_PLAIN>
struct EpiStore {
  OutT* out; int64_t ld;
  const float* bias;        // [N] or null
  const float* addend;      // EPI_ADDEND: [M, ld_add] fp32 (the y-part of encoder_gmm layer 0 in fp32 mode)
  int64_t ld_add;
  int relu;                 // 0 none, 1 ReLU, 2 sigmoid (Bernoulli mean of the decoder, inference only)
  float scale;
  float shift;              // constant added with the bias (gen_bias_init); 0 when omitted from the initialiser
  uint32_t* relu_bits;      // chained kernel only: [N/32, ld_bits >= M] words, bit (31 - j) of word [w][m] = (out[m, 32 w + j] > 0); or null
  int ld_bits;

  template <int NV> struct Pre { float b[NV >= 16 ? 1 : NV]; float a[(MODE != EPI_PLAIN) ? NV : 1]; };

  template <int BN>
  __device__ __forceinline__ void tile_begin(int n0, int N, float* sbias, int t, int nt) const { stage_bias<BN>(bias, n0, N, sbias, t, nt); }
  __device__ __forceinline__ float* colsum_dst() const { return nullptr; }
  __host__ __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  const void* out_ptr() const { return out; }      // host: dependency inference of GEMM chains (gemm_chain.cuh)
  const void* read_ptr() const { return MODE == EPI_ADDEND ? (const void*)addend : nullptr; }

  template <int NV>
  __device__ __forceinline__ Pre<NV> prefetch(int m, int n0, int nvalid, bool valid, const EpiCtx&) const {
    Pre<NV> p = {};
    if constexpr (NV < 16) {
      if (bias) {
        load_frag<NV>(bias + n0, p.b, nvalid);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) p.b[i] = 0.f;
      }
    }
    if constexpr (MODE == EPI_ADDEND) {
      if (valid) load_frag<NV>(addend + (int64_t)m * ld_add + n0, p.a, nvalid);
    } else if constexpr (MODE == EPI_ACCUM) {
      if (valid) load_frag<NV>(out + (int64_t)m * ld + n0, p.a, nvalid);
    }
    return p;
  }
  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid, bool valid, const Pre<NV>& p, const EpiCtx& ctx) {
    constexpr bool STAGED = staged_io<OutT, NV>::value && MODE == EPI_PLAIN;
    if (!STAGED && !valid) return;
    float v[NV];
    float bb[NV >= 16 ? NV : 1];
    if constexpr (NV >= 16) {
      const uint32_t sb = smem_addr(ctx.sbias);
#pragma unroll
      for (int i = 0; i < NV; i += 4) { float4 t = lds128f(sb + 4 * i); bb[i] = t.x; bb[i + 1] = t.y; bb[i + 2] = t.z; bb[i + 3] = t.w; }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if constexpr (NV >= 16) v[i] = fmaf(acc[i], scale, bb[i]); else v[i] = fmaf(acc[i], scale, p.b[i]);
      if constexpr (MODE == EPI_ADDEND) v[i] += p.a[i];
      if (relu == 1) v[i] = fmaxf(v[i], 0.f);
      else if (relu == 2) v[i] = sigmoid_f(v[i] + shift);
      if constexpr (MODE == EPI_ACCUM) v[i] += p.a[i];
    }
    // columns >= nvalid need no masking on the staged path: their accumulators are exactly 0 (the B
    // operand rows beyond N are zero-filled by TMA) and tile_begin staged a 0 bias for them.
    if constexpr (STAGED) {
      if (ctx.patch) {
        store_chunk_bf16<NV>(reinterpret_cast<bf16*>(out) + (int64_t)(m - (int)(threadIdx.x & 31)) * ld + n0, ld, v, nvalid, ctx);
        return;
      }
      if (!valid) return;
    }
    store_frag<NV>(out + (int64_t)m * ld + n0, v, nvalid);
  }
  __device__ __forceinline__ void finish_warp() {}
};

// ---- decoder output layer fused with the Bernoulli log-likelihood ---------------------------
// logits = acc + b + gen_bias_init (base.py:135); log p(x|z) = sum_d [x l - softplus(l)]
// (gmvae.py:254, vae.py:177; TFP Bernoulli = -sigmoid_cross_entropy); d nll / d logits =
// (sigmoid(l) - x) / B.  Logits never reach HBM; only the gradient does.
template <typename OutT>
struct EpiBCE {
  OutT* dlogits; int64_t ld;
  const float* bias; float gen_bias;
  const uint8_t* x; int64_t ldx;
  int x_row_div;            // x row = m / x_row_div (objective M: K consecutive rows share one image)
  const float* row_weight;  // [M] q(y=k|x) weights (objective M) or null -> 1
  double* row_sum;          // [M] per-row log-likelihood (objective M; double: |rec| ~ 550, its differences ~ 1) or null
  float* nll_acc;           // base of the loss accumulators (acc_add): nll += -inv_bg * sum(w * loglik)
  float inv_bg;             // 1 / global batch
  float partial;
  float* colsum;            // fused bias gradient db[n] += sum_m dlogits[m,n] (tensor-core path only) or null
  int fast;                 // 1: fast intrinsics (bf16 mode), 0: accurate libm (fp32 validation mode)

  template <int NV> struct Pre { float b[NV >= 16 ? 1 : NV]; float x[NV]; float w; };

  template <int BN>
  __device__ __forceinline__ void tile_begin(int n0, int N, float* sbias, int t, int nt) const { stage_bias<BN>(bias, n0, N, sbias, t, nt); }
  __device__ __forceinline__ float* colsum_dst() const { return colsum; }
  __host__ __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  const void* out_ptr() const { return dlogits; }
  const void* read_ptr() const { return nullptr; }

  template <int NV>
  __device__ __forceinline__ Pre<NV> prefetch(int m, int n0, int nvalid, bool valid, const EpiCtx&) const {
    Pre<NV> p;
    if constexpr (NV < 16) {
      if (bias) {
        load_frag<NV>(bias + n0, p.b, nvalid);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) p.b[i] = 0.f;
      }
    }
    if (valid) {
      load_frag<NV>(x + (int64_t)(m / x_row_div) * ldx + n0, p.x, nvalid);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) p.x[i] = 0.f;
    }
    p.w = (valid && row_weight) ? __ldg(row_weight + m) : 1.f;
    return p;
  }
  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid, bool valid, const Pre<NV>& p, const EpiCtx& ctx) {
    float d[NV];
    float ll = 0.f;
    const float wscale = p.w * inv_bg;
    float bb[NV >= 16 ? NV : 1];
    if constexpr (NV >= 16) {
      const uint32_t sb = smem_addr(ctx.sbias);
#pragma unroll
      for (int i = 0; i < NV; i += 4) { float4 t = lds128f(sb + 4 * i); bb[i] = t.x; bb[i + 1] = t.y; bb[i + 2] = t.z; bb[i + 3] = t.w; }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float l;
      if constexpr (NV >= 16) l = acc[i] + gen_bias + bb[i]; else l = acc[i] + gen_bias + p.b[i];
      float e, sp, inv1pe;
      if (fast) {
        // exp(-|l|) in (0,1]; log(1+e) with 1+e in (1,2]: absolute error of the intrinsics ~1e-7
        e = __expf(-fabsf(l));
        sp = fmaxf(l, 0.f) + __logf(1.f + e);
        inv1pe = __fdividef(1.f, 1.f + e);
      } else {
        e = expf(-fabsf(l));
        sp = fmaxf(l, 0.f) + log1pf(e);
        inv1pe = 1.f / (1.f + e);
      }
      const float sg = l >= 0.f ? inv1pe : e * inv1pe;
      const bool ok = valid && i < nvalid;
      ll += ok ? fmaf(p.x[i], l, -sp) : 0.f;
      d[i] = ok ? (sg - p.x[i]) * wscale : 0.f;
    }
    if (valid) {
      if (row_sum) atomicAdd(row_sum + m, (double)ll);
      partial += p.w * ll;
    }
    if constexpr (staged_io<OutT, NV>::value) {
      // tensor-core kernel: coalesced store through the warp's patch; the bias gradient (column sums
      // of the values as stored, bf16-rounded) is read back from the same patch
      store_chunk_bf16<NV>(reinterpret_cast<bf16*>(dlogits) + (int64_t)(m - (int)(threadIdx.x & 31)) * ld + n0, ld, d, nvalid, ctx,
                           colsum != nullptr);
      if (colsum) colsum_from_patch<NV>(nvalid, ctx);
    } else {
      if (valid) store_frag<NV>(dlogits + (int64_t)m * ld + n0, d, nvalid);
    }
  }
  __device__ __forceinline__ void finish_warp() {
    float s = warp_sum(partial);
    if ((threadIdx.x & 31) == 0 && s != 0.f) acc_add(nll_acc, ACC_NLL, -inv_bg * s);
    partial = 0.f;
  }
};

// ---- backward through a ReLU layer: out = acc * [h > 0] -------------------------------------
template <typename OutT, typename HT>
struct EpiReluMask {
  OutT* out; int64_t ld;
  const HT* h; int64_t ldh;
  float* colsum;            // fused bias gradient of the layer below (tensor-core path only) or null
  const uint32_t* relu_bits;  // chained kernel only: the 1-bit form of [h > 0] written by the forward job (EpiStore::relu_bits), or null
  int ld_bits;

  template <int NV> struct Pre {
    float h[staged_io<HT, NV>::value ? 1 : NV];
    ChunkRegs<staged_io<HT, NV>::value ? NV : 8> g;
  };

  template <int BN>
  __device__ __forceinline__ void tile_begin(int, int, float*, int, int) const {}
  __device__ __forceinline__ float* colsum_dst() const { return colsum; }
  __host__ __device__ __forceinline__ const float* bias_ptr() const { return nullptr; }
  const void* out_ptr() const { return out; }
  const void* read_ptr() const { return h; }

  template <int NV>
  __device__ __forceinline__ Pre<NV> prefetch(int m, int n0, int nvalid, bool valid, const EpiCtx& ctx) const {
    Pre<NV> p;
    if constexpr (staged_io<HT, NV>::value) {
      // coalesced: 16-byte pieces of the 32 x NV mask chunk, exchanged through the patch in row()
      p.g = load_chunk_bf16_issue<NV>(reinterpret_cast<const bf16*>(h) + (int64_t)(m - (int)(threadIdx.x & 31)) * ldh + n0, ldh, nvalid, ctx);
    } else {
      if (valid) {
        load_frag<NV>(h + (int64_t)m * ldh + n0, p.h, nvalid);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) p.h[i] = 0.f;
      }
    }
    return p;
  }
  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid, bool valid, const Pre<NV>& p, const EpiCtx& ctx) {
    float v[NV];
    if constexpr (staged_io<HT, NV>::value) {
      float hv[NV];
      load_chunk_bf16_finish<NV>(p.g, hv, ctx);
#pragma unroll
      // rows >= M and columns >= N have zero accumulators (TMA zero fill), so only the mask is applied
      for (int i = 0; i < NV; ++i) v[i] = hv[i] > 0.f ? acc[i] : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = (p.h[i] > 0.f && i < nvalid) ? acc[i] : 0.f;
    }
    if constexpr (staged_io<OutT, NV>::value) {
      store_chunk_bf16<NV>(reinterpret_cast<bf16*>(out) + (int64_t)(m - (int)(threadIdx.x & 31)) * ld + n0, ld, v, nvalid, ctx,
                           colsum != nullptr);
      if (colsum) colsum_from_patch<NV>(nvalid, ctx);   // bias gradient of the layer below, as stored
    } else {
      if (valid) store_frag<NV>(out + (int64_t)m * ld + n0, v, nvalid);
    }
  }
  __device__ __forceinline__ void finish_warp() {}
};

// ---- split-K weight gradient: out[m,n] += acc (fp32 reductions into the flat gradient buffer) ----
struct EpiAtomicAdd {
  float* out; int64_t ld;
  template <int NV> struct Pre {};
  template <int BN>
  __device__ __forceinline__ void tile_begin(int, int, float*, int, int) const {}
  __device__ __forceinline__ float* colsum_dst() const { return nullptr; }
  __host__ __device__ __forceinline__ const float* bias_ptr() const { return nullptr; }
  const void* out_ptr() const { return nullptr; }   // accumulated into the gradient buffer: consumed after the step's last GEMM only
  const void* read_ptr() const { return nullptr; }
  template <int NV>
  __device__ __forceinline__ Pre<NV> prefetch(int, int, int, bool, const EpiCtx&) const { return Pre<NV>(); }
  template <int NV>
  __device__ __forceinline__ void row(int m, int n0, const float* acc, int nvalid, bool valid, const Pre<NV>&, const EpiCtx&) {
    if (!valid) return;
    float* dst = out + (int64_t)m * ld + n0;
    if (NV % 4 == 0 && nvalid == NV && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) red_add_v4(dst + i, acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i < nvalid) atomicAdd(dst + i, acc[i]);
    }
  }
  __device__ __forceinline__ void finish_warp() {}
};

}  // namespace gmvae
