// Host simulation of the peer-memory all-reduce (gmvae_b200/csrc/peer.cuh): the same layout and element functions the
// kernels use, run for `world` ranks one after the other over host memory.  Checks the index arithmetic (shards, ragged last
// shard, buffers shorter than the world size) and the summation order; the flags / fences / NVLink part needs GPUs.
// Test infrastructure.
#include <cstring>
#include <vector>

#include "../../gmvae_b200/csrc/peer.cuh"

extern "C" __attribute__((visibility("default"))) int host_peer_allreduce(int world, long long n_floats, float* grads /* [world][n_floats], in/out */) {
  using namespace gmvae::peer;
  if (world < 1 || world > MAX_WORLD || n_floats % 4 != 0) return -1;
  const Layout L = make_layout(world, n_floats);
  if (L.red_off < L.grad_off + (size_t)L.n4 * 16 || L.flags_off < L.red_off + (size_t)L.n4 * 16 || L.local_off < L.flags_off + 2 * MAX_WORLD * 8 ||
      L.bytes < L.local_off + sizeof(Local) || L.cap4 * world < L.n4)
    return -2;                                                      // regions overlap or shards do not cover the buffer
  std::vector<std::vector<char>> region(world, std::vector<char>(L.bytes, (char)0x7F));   // poison: unwritten memory must never be read
  Peers P;
  for (int r = 0; r < world; ++r) {
    P.grad[r] = reinterpret_cast<float4*>(region[r].data() + L.grad_off);
    P.red[r] = reinterpret_cast<float4*>(region[r].data() + L.red_off);
    P.flags[r] = reinterpret_cast<unsigned long long*>(region[r].data() + L.flags_off);
    memcpy(P.grad[r], grads + (size_t)r * n_floats, (size_t)n_floats * 4);       // the backward pass accumulates into the region
  }
  int64_t covered = 0;
  for (int rank = 0; rank < world; ++rank) {                        // exchange_kernel on every rank
    const int64_t b4 = shard_begin4(L, rank), len4 = shard_len4(L, rank);
    if (b4 < 0 || b4 + len4 > L.n4) return -3;
    covered += len4;
    for (int64_t i = 0; i < len4; ++i) {
      const float4 s = reduce_ranks(L, P, b4 + i, LoadPeer());
      for (int p = 0; p < world; ++p) P.red[p][b4 + i] = s;
    }
  }
  if (covered != L.n4) return -4;                                   // the shards tile the buffer exactly
  for (int rank = 0; rank < world; ++rank)                          // what Adam reads (gather_kernel in the stand-alone form)
    memcpy(grads + (size_t)rank * n_floats, P.red[rank], (size_t)n_floats * 4);
  return 0;
}
