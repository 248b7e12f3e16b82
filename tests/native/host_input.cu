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

// what binarize_kernel's threads do, one quad after the other
__attribute__((visibility("default"))) void host_binarize(const uint8_t* src, const int64_t* row_index, int D, int64_t n_out, uint64_t seed,
                                                          uint64_t draw, uint64_t rank, int vec, uint8_t* out) {
  const uint64_t key = gmvae::binarize_key(seed, draw), stream = gmvae::binarize_stream(rank);
  const int64_t n_quads = (n_out + 3) / 4;
  for (int64_t q = 0; q < n_quads; ++q) gmvae::binarize_quad(src, row_index, D, n_out, key, stream, q, vec != 0, out);
}
}
