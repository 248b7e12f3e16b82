// input.cuh -- the input pipeline's per-sample arithmetic on the device: dynamic binarisation of
// image intensities, runners.create_dataset._preprocess (/root/reference/scripts/runners.py:44-47):
//     image = tf.cast(sample['image'], tf.float32) / 255.
//     image = image < tf.random.uniform(tf.shape(image))          (inverted: pixel is 1 with prob 1 - intensity)
// HBM-bound byte work (D bytes in, D bytes out per sample); the uniforms come from the step's own
// Philox generator so no random tensor ever touches memory.  At the step's throughput a host
// pipeline cannot feed one GPU (4e7 samples/s x 784 B = 32 GB/s of freshly binarised bytes per GPU),
// so the intensities stay resident in HBM (MNIST train: 47 MB) and every batch is binarised in place.
#pragma once
#include "common.cuh"

namespace gmvae {

// Philox stream of the binarisation draws: bit 63 set, so it can never collide with the step's
// noise streams (2*rank, 2*rank+1).
constexpr uint64_t BINARIZE_STREAM = 0x8000000000000000ull;
constexpr uint64_t DRAW_MIX = 0x9E3779B97F4A7C15ull;

// Philox key / stream of one binarisation draw (shared by the kernel and by the host build the tests run)
__host__ __device__ __forceinline__ uint64_t binarize_key(uint64_t seed, uint64_t draw) { return seed ^ (draw * DRAW_MIX); }
__host__ __device__ __forceinline__ uint64_t binarize_stream(uint64_t rank) { return BINARIZE_STREAM + rank; }

// intensity byte -> the reference's fp32 intensity (IEEE division, round to nearest: same bits on host and device)
__host__ __device__ __forceinline__ float unit_intensity(uint8_t v) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn((float)v, 255.0f);
#else
  return (float)v / 255.0f;
#endif
}

// Output bytes [4q, 4q+4) of the flat [batch, D] result: one Philox call, four comparisons.
// Source row of output row r: row_index[r] if given, else r.  `vec` = D % 4 == 0 and both bases
// 4-byte aligned: the quad lies inside one row and moves as one 32-bit word.
__host__ __device__ __forceinline__ void binarize_quad(const uint8_t* __restrict__ src, const int64_t* __restrict__ row_index, int D,
                                                       int64_t n_out, uint64_t key, uint64_t stream, int64_t q, bool vec,
                                                       uint8_t* __restrict__ out) {
  uint32_t r[4];
  Philox::gen(key, stream, (uint64_t)q, r);
  const int64_t e0 = q * 4;
  if (vec) {
    const int64_t row = e0 / D, col = e0 - row * D;
    const int64_t srow = row_index ? row_index[row] : row;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(src + srow * D + col);
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint8_t v = (uint8_t)(w >> (8 * j));
      o |= (unit_intensity(v) < u01(r[j]) ? 1u : 0u) << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(out + e0) = o;
  } else {
    for (int j = 0; j < 4; ++j) {
      const int64_t e = e0 + j;
      if (e >= n_out) break;
      const int64_t row = e / D, col = e - row * D;
      const int64_t srow = row_index ? row_index[row] : row;
      out[e] = unit_intensity(src[srow * D + col]) < u01(r[j]) ? 1 : 0;
    }
  }
}

#ifdef __CUDACC__
struct DeviceState;
// grid-stride over quads; `seed` is read from the handle's device state like the step's noise
__global__ void binarize_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ row_index, int D, int64_t n_out,
                                const unsigned long long* __restrict__ seed, uint64_t draw, uint64_t rank, int vec,
                                uint8_t* __restrict__ out) {
  griddep_wait();
  griddep_launch();
  const uint64_t key = binarize_key((uint64_t)(*seed), draw), stream = binarize_stream(rank);
  const int64_t n_quads = (n_out + 3) / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x)
    binarize_quad(src, row_index, D, n_out, key, stream, q, vec != 0, out);
}
#endif

}  // namespace gmvae
