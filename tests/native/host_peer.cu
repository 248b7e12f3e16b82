// Host simulation of the peer-memory all-reduce (gmvae_b200/csrc/peer.cuh): the same layout and element functions the
// kernels use, run for `world` ranks one after the other over host memory.  Checks the index arithmetic (shards, slots,
// ragged last shard) and the summation order; the flags / fences / NVLink part needs GPUs.  Test infrastructure.
#include <cstring>
#include <vector>

#include "../../gmvae_b200/csrc/peer.cuh"

extern "C" __attribute__((visibility("default"))) int host_peer_allreduce(int world, long long n_floats, float* grads /* [world][n_floats], in/out */) {
  using namespace gmvae::peer;
  if (world < 1 || world > MAX_WORLD || n_floats % 4 != 0) return -1;
  const Layout L = make_layout(world, n_floats);
  if (L.red_off < (size_t)world * L.cap4 * 16 || L.flags_off < L.red_off + (size_t)world * L.cap4 * 16 || L.local_off < L.flags_off + 2 * MAX_WORLD * 8 ||
      L.bytes < L.local_off + sizeof(Local) || L.cap4 * world < L.n4)
    return -2;                                                      // regions overlap or shards do not cover the buffer
  std::vector<std::vector<char>> region(world, std::vector<char>(L.bytes, (char)0x7F));   // poison: unwritten slots must never be read
  Peers P;
  for (int r = 0; r < world; ++r) {
    P.recv[r] = reinterpret_cast<float4*>(region[r].data() + L.recv_off);
    P.red[r] = reinterpret_cast<float4*>(region[r].data() + L.red_off);
    P.flags[r] = reinterpret_cast<unsigned long long*>(region[r].data() + L.flags_off);
  }
  for (int rank = 0; rank < world; ++rank) {                        // phase A on every rank
    const float4* g = reinterpret_cast<const float4*>(grads + (size_t)rank * n_floats);
    for (int64_t i4 = 0; i4 < L.n4; ++i4) {
      int owner; int64_t dst4;
      push_target(L, rank, i4, owner, dst4);
      if (owner < 0 || owner >= world || dst4 < 0 || dst4 >= (int64_t)world * L.cap4) return -3;
      P.recv[owner][dst4] = g[i4];
    }
  }
  for (int rank = 0; rank < world; ++rank)                          // phase B
    for (int64_t i = 0; i < shard_len4(L, rank); ++i) {
      const float4 s = reduce_slots(L, P.recv[rank], i, LoadPeerWritten());
      for (int p = 0; p < world; ++p) P.red[p][red_index(L, rank, i)] = s;
    }
  for (int rank = 0; rank < world; ++rank) {                        // phase C
    float4* g = reinterpret_cast<float4*>(grads + (size_t)rank * n_floats);
    for (int64_t i4 = 0; i4 < L.n4; ++i4) g[i4] = P.red[rank][i4];
  }
  return 0;
}
