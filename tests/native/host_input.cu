// Host build of the input pipeline's __host__ __device__ arithmetic (gmvae_b200/csrc/input.cuh, the Philox generator of
// common.cuh): the SAME source the CUDA kernel runs, compiled for the CPU, so that tests without a GPU can compare its
// indexing and arithmetic with the numpy oracle bit for bit.  Test infrastructure; not part of libgmvae_b200.so.
#include "../../gmvae_b200/csrc/input.cuh"

extern "C" {

__attribute__((visibility("default"))) void host_philox(uint64_t seed, uint64_t stream, uint64_t ctr, uint32_t* out4) {
  uint32_t r[4];
  gmvae::Philox::gen(seed, stream, ctr, r);
  for (int i = 0; i < 4; ++i) out4[i] = r[i];
}

__attribute__((visibility("default"))) float host_u01(uint32_t r) { return gmvae::u01(r); }

// what binarize_kernel's threads do, one work item after the other; mode = BINARIZE_BYTES / VEC4 / VEC16 (-1: the widest allowed)
__attribute__((visibility("default"))) int host_binarize(const uint8_t* src, const int64_t* row_index, int D, int64_t n_out, uint64_t seed,
                                                         uint64_t draw, uint64_t rank, int mode, uint8_t* out) {
  uint32_t T[256];
  for (uint32_t v = 0; v < 256; ++v) T[v] = gmvae::binarize_threshold(v);
  gmvae::PhiloxKeys rk;
  gmvae::philox_schedule(gmvae::binarize_key(seed, draw), rk);
  const uint64_t stream = gmvae::binarize_stream(rank);
  if (mode < 0) mode = gmvae::binarize_mode(src, out, D);
  if (mode == gmvae::BINARIZE_VEC16) {
    for (int64_t g = 0; g < n_out / 16; ++g) gmvae::binarize_group16(src, row_index, D, n_out, rk, stream, g, T, out);
  } else {
    for (int64_t q = 0; q < (n_out + 3) / 4; ++q)
      gmvae::binarize_quad(src, row_index, D, n_out, rk, stream, q, mode == gmvae::BINARIZE_VEC4, T, out);
  }
  return mode;
}

// the threshold table against the reference's comparison taken literally: 1 = they agree for this (v, r)
__attribute__((visibility("default"))) int host_threshold_agrees(uint32_t v, uint32_t r) {
  return ((r >> 9) >= gmvae::binarize_threshold(v) ? 1u : 0u) == gmvae::binarize_direct((uint8_t)v, r);
}
__attribute__((visibility("default"))) uint32_t host_threshold(uint32_t v) { return gmvae::binarize_threshold(v); }

// the scheduled Philox against Philox::gen of common.cuh
__attribute__((visibility("default"))) void host_philox_scheduled(uint64_t seed, uint64_t stream, uint64_t ctr, uint32_t* out4) {
  gmvae::PhiloxKeys rk;
  gmvae::philox_schedule(seed, rk);
  uint32_t r[4];
  gmvae::philox_gen_scheduled(rk, stream, ctr, r);
  for (int i = 0; i < 4; ++i) out4[i] = r[i];
}
}
