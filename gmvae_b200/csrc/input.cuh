// input.cuh -- the input pipeline's per-sample arithmetic on the device: dynamic binarisation of
// image intensities, runners.create_dataset._preprocess (/root/reference/scripts/runners.py:44-47):
//     image = tf.cast(sample['image'], tf.float32) / 255.
//     image = image < tf.random.uniform(tf.shape(image))          (inverted: pixel is 1 with prob 1 - intensity)
// HBM-bound byte work (D bytes in, D bytes out per sample); the uniforms come from the step's own
// Philox generator so no random tensor ever touches memory.  At the step's throughput a host
// pipeline cannot feed one GPU (4e7 samples/s x 784 B = 32 GB/s of freshly binarised bytes per GPU),
// so the intensities stay resident in HBM (MNIST train: 47 MB) and every batch is binarised in place.
//
// What the instruction count is spent on decides the speed here (a Philox call per 4 bytes), so:
//   * the ten round keys depend only on (seed, draw): the host expands them once and passes them as
//     kernel parameters -- they are constant-bank operands of the XORs, no instructions;
//   * the comparison `fl(v / 255) < u` with u = ((r >> 9) + 0.5) * 2^-23 is `(r >> 9) >= T[v]` for a
//     256-entry threshold table (built per block in shared memory with the exact IEEE division), which
//     replaces two int->float conversions, a division and a float compare per byte;
//   * 16 bytes per thread and iteration (one 128-bit load, four independent Philox calls, one 128-bit
//     store) when rows are multiples of 16 bytes; no 64-bit division on the contiguous path.
#pragma once
#include "common.cuh"

namespace gmvae {

// Philox stream of the binarisation draws: bit 63 set, so it can never collide with the step's
// noise streams (2*rank, 2*rank+1).
constexpr uint64_t BINARIZE_STREAM = 0x8000000000000000ull;
constexpr uint64_t DRAW_MIX = 0x9E3779B97F4A7C15ull;

// Philox key / stream of one binarisation draw (shared by the kernel launch and by the host build the tests run)
__host__ __device__ __forceinline__ uint64_t binarize_key(uint64_t seed, uint64_t draw) { return seed ^ (draw * DRAW_MIX); }
__host__ __device__ __forceinline__ uint64_t binarize_stream(uint64_t rank) { return BINARIZE_STREAM + rank; }

// The key schedule of Philox4x32-10 (key += (0x9E3779B9, 0xBB67AE85) per round), expanded once.
struct PhiloxKeys { uint32_t k[20]; };
__host__ __device__ __forceinline__ void philox_schedule(uint64_t key, PhiloxKeys& rk) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  for (int i = 0; i < 10; ++i) {
    rk.k[2 * i] = k0; rk.k[2 * i + 1] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// Same function as Philox::gen(key, stream, ctr) of common.cuh with the schedule precomputed; each
// round is two 32x32->64 multiplies and two three-input XORs.
__host__ __device__ __forceinline__ void philox_gen_scheduled(const PhiloxKeys& rk, uint64_t stream, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk.k[2 * i], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk.k[2 * i + 1];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// intensity byte -> the reference's fp32 intensity (IEEE division, round to nearest: same bits on host and device)
__host__ __device__ __forceinline__ float unit_intensity(uint8_t v) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn((float)v, 255.0f);
#else
  return (float)v / 255.0f;
#endif
}
// The reference's comparison, literally (runners.py:45-46), for one intensity byte and one random word.
__host__ __device__ __forceinline__ uint32_t binarize_direct(uint8_t v, uint32_t r) { return unit_intensity(v) < u01(r) ? 1u : 0u; }
// Smallest m = r >> 9 for which binarize_direct is 1 (2^23 = never, v = 255): with a = fl(v/255),
// a < (m + 0.5) 2^-23  <=>  m > a 2^23 - 0.5  <=>  m >= floor(a 2^23 + 0.5); a 2^23 + 0.5 is exact in fp64.
__host__ __device__ __forceinline__ uint32_t binarize_threshold(uint32_t v) {
  return (uint32_t)((double)unit_intensity((uint8_t)v) * 8388608.0 + 0.5);
}
// four intensity bytes of a 32-bit word against four random words
__host__ __device__ __forceinline__ uint32_t binarize_word(uint32_t w, const uint32_t (&r)[4], const uint32_t* __restrict__ T) {
  uint32_t o = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) o |= ((r[j] >> 9) >= T[(w >> (8 * j)) & 0xFFu] ? 1u : 0u) << (8 * j);
  return o;
}
// output row of flat element e (n_out < 2^32 takes the 32-bit division)
__host__ __device__ __forceinline__ int64_t row_of(int64_t e, int D, bool small) {
  return small ? (int64_t)((uint32_t)e / (uint32_t)D) : e / D;
}

// Output bytes [4q, 4q+4) of the flat [batch, D] result: one Philox call (counter q), four comparisons.
// Source row of output row r: row_index[r] if given, else r.  `vec` = D % 4 == 0 and both bases
// 4-byte aligned: the quad lies inside one row and moves as one 32-bit word.
__host__ __device__ __forceinline__ void binarize_quad(const uint8_t* __restrict__ src, const int64_t* __restrict__ row_index, int D,
                                                       int64_t n_out, const PhiloxKeys& rk, uint64_t stream, int64_t q, bool vec,
                                                       const uint32_t* __restrict__ T, uint8_t* __restrict__ out) {
  uint32_t r[4];
  philox_gen_scheduled(rk, stream, (uint64_t)q, r);
  const int64_t e0 = q * 4;
  const bool small = n_out <= 0xFFFFFFFFll;
  if (vec) {
    int64_t s = e0;
    if (row_index) {
      const int64_t row = row_of(e0, D, small);
      s = row_index[row] * D + (e0 - row * D);
    }
    *reinterpret_cast<uint32_t*>(out + e0) = binarize_word(*reinterpret_cast<const uint32_t*>(src + s), r, T);
  } else {
    for (int j = 0; j < 4; ++j) {
      const int64_t e = e0 + j;
      if (e >= n_out) break;
      const int64_t row = row_of(e, D, small);
      const int64_t srow = row_index ? row_index[row] : row;
      out[e] = (uint8_t)((r[j] >> 9) >= T[src[srow * D + (e - row * D)]] ? 1 : 0);
    }
  }
}

// Output bytes [16g, 16g+16): quads 4g .. 4g+3 (same counters, same bytes as four binarize_quad calls).
// Requires D % 16 == 0 and 16-byte aligned bases, so the group lies inside one row.
struct alignas(16) Bytes16 { uint32_t w[4]; };
__host__ __device__ __forceinline__ void binarize_group16(const uint8_t* __restrict__ src, const int64_t* __restrict__ row_index, int D,
                                                          int64_t n_out, const PhiloxKeys& rk, uint64_t stream, int64_t g,
                                                          const uint32_t* __restrict__ T, uint8_t* __restrict__ out) {
  const int64_t e0 = g * 16;
  int64_t s = e0;
  if (row_index) {
    const int64_t row = row_of(e0, D, n_out <= 0xFFFFFFFFll);
    s = row_index[row] * D + (e0 - row * D);
  }
  const Bytes16 in = *reinterpret_cast<const Bytes16*>(src + s);
  Bytes16 o;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t r[4];
    philox_gen_scheduled(rk, stream, (uint64_t)(g * 4 + i), r);
    o.w[i] = binarize_word(in.w[i], r, T);
  }
  *reinterpret_cast<Bytes16*>(out + e0) = o;
}

enum { BINARIZE_BYTES = 0, BINARIZE_VEC4 = 1, BINARIZE_VEC16 = 2 };
// widest path the shapes and alignments allow
inline int binarize_mode(const void* src, const void* out, int D) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out);
  if (D % 16 == 0 && (a & 15) == 0) return BINARIZE_VEC16;
  if (D % 4 == 0 && (a & 3) == 0) return BINARIZE_VEC4;
  return BINARIZE_BYTES;
}

#ifdef __CUDACC__
constexpr int BINARIZE_THREADS = 256;
__global__ void __launch_bounds__(BINARIZE_THREADS)
binarize_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ row_index, int D, int64_t n_out,
                const __grid_constant__ PhiloxKeys rk, uint64_t rank, int mode, uint8_t* __restrict__ out) {
  __shared__ uint32_t T[256];
  griddep_wait();
  griddep_launch();
  T[threadIdx.x] = binarize_threshold(threadIdx.x);          // blockDim.x == 256
  __syncthreads();
  const uint64_t stream = binarize_stream(rank);
  const int64_t first = (int64_t)blockIdx.x * BINARIZE_THREADS + threadIdx.x, step = (int64_t)gridDim.x * BINARIZE_THREADS;
  if (mode == BINARIZE_VEC16) {
    const int64_t n_groups = n_out / 16;
    for (int64_t g = first; g < n_groups; g += step) binarize_group16(src, row_index, D, n_out, rk, stream, g, T, out);
  } else {
    const int64_t n_quads = (n_out + 3) / 4;
    for (int64_t q = first; q < n_quads; q += step) binarize_quad(src, row_index, D, n_out, rk, stream, q, mode == BINARIZE_VEC4, T, out);
  }
}
#endif

// ---- bit-packed binary images ----------------------------------------------------------------------------------
// The step's input is binary ({0,1} per pixel, runners.py:44-47), so the host side of a batch is 1 bit per pixel:
// packed[b, j] holds pixels 8j .. 8j+7 of row b, most significant bit first (numpy.packbits order); D = 784 -> 98 bytes per
// image instead of 784 over PCIe.  One thread per packed byte: one 8-byte store.
__global__ void unpack_bits_kernel(const uint8_t* __restrict__ packed, int64_t n_bytes, int row_bytes, int D, uint8_t* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bytes) return;
  const int64_t b = i / row_bytes; const int j = (int)(i - b * row_bytes);
  const uint32_t v = packed[i];
  uint8_t* dst = x + b * D + 8 * j;
  if (8 * j + 8 <= D && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    // byte k of the output (little endian) = bit (7 - k) of v
    const uint32_t lo = ((v >> 7) & 1u) | (((v >> 6) & 1u) << 8) | (((v >> 5) & 1u) << 16) | (((v >> 4) & 1u) << 24);
    const uint32_t hi = ((v >> 3) & 1u) | (((v >> 2) & 1u) << 8) | (((v >> 1) & 1u) << 16) | ((v & 1u) << 24);
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  } else {
    for (int k = 0; k < 8 && 8 * j + k < D; ++k) dst[k] = (uint8_t)((v >> (7 - k)) & 1u);
  }
}

}  // namespace gmvae
