// Concurrency model of the peer-memory all-reduce protocol (gmvae_b200/csrc/peer.cuh), run on the CPU under ThreadSanitizer.
//
// peer.cuh's kernels synchronise ranks with epoch flags (st.release.sys / ld.acquire.sys), publish them from the last block of
// a grid (fence + ticket counter) and reuse the gradient / reduced buffers across steps without double buffering.  This program
// restates exactly that protocol with host atomics -- one std::thread per (rank, block); the kernels of one rank are separated by a
// per-rank barrier (stream order); NOTHING orders different ranks except the flags -- and uses peer.cuh's own layout and element
// functions for the data movement.  Plain (non-atomic) accesses to the buffers make every missing happens-before edge a TSan
// report; random delays shake the interleavings; every rank checks every step's result.  It validates the protocol's design (no
// deadlock, no race when buffers are reused by the next step), not the CUDA code itself.  Test infrastructure.
//
//   host_peer_protocol <world> <n_floats> <steps> <blocks_per_rank>      exit code 0 = all ranks saw the exact sums
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "../../gmvae_b200/csrc/peer.cuh"

using namespace gmvae::peer;

static float value_of(int rank, int step, int64_t i) { return (float)((rank + 1) * 3 + (step % 7) * 5 + (int)(i % 11)); }   // small integers: sums are exact

struct Rank {
  std::vector<char> region;
  std::unique_ptr<std::barrier<>> stream;     // kernel boundaries of this rank (stream order)
};

static unsigned long long ld_acquire(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static void st_release(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
// last_block(): every block fences and takes a ticket; the last one resets the counter and publishes
// (device: __threadfence_system() + relaxed atomicAdd on both sides; here the acq_rel read-modify-write carries the same edges,
// which is also the form ThreadSanitizer understands -- it does not model stand-alone fences)
static bool last_block_host(unsigned int* done, int nblocks) {
  const bool last = __atomic_fetch_add(done, 1u, __ATOMIC_ACQ_REL) == (unsigned)nblocks - 1;
  if (last) __atomic_store_n(done, 0u, __ATOMIC_RELAXED);
  return last;
}
static int g_mutation = 0;   // PEER_MUTATE: 1 = Adam does not wait for the landed shards, 2 = the exchange does not wait for the gradients
static void wait_epochs_host(const unsigned long long* flags, int world, unsigned long long epoch) {
  for (int r = 0; r < world; ++r)
    while (ld_acquire(flags + r) < epoch) std::this_thread::yield();
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 4;
  const long long n = argc > 2 ? atoll(argv[2]) : 4096;
  const int steps = argc > 3 ? atoi(argv[3]) : 20;
  const int blocks = argc > 4 ? atoi(argv[4]) : 3;
  if (world < 2 || world > MAX_WORLD || n % 4 != 0 || blocks < 1) return 2;
  if (const char* m = getenv("PEER_MUTATE")) g_mutation = atoi(m);
  const Layout L = make_layout(world, n);
  std::vector<Rank> ranks(world);
  Peers P;
  for (int r = 0; r < world; ++r) {
    ranks[r].region.assign(L.bytes, 0);
    ranks[r].stream = std::make_unique<std::barrier<>>(blocks);
    P.grad[r] = reinterpret_cast<float4*>(ranks[r].region.data() + L.grad_off);
    P.red[r] = reinterpret_cast<float4*>(ranks[r].region.data() + L.red_off);
    P.flags[r] = reinterpret_cast<unsigned long long*>(ranks[r].region.data() + L.flags_off);
  }
  std::atomic<int> errors{0};
  auto worker = [&](int rank, int block) {
    std::mt19937 rng(rank * 131 + block);
    auto jitter = [&]() { if (rng() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 200)); };
    Rank& me = ranks[rank];
    Local* loc = reinterpret_cast<Local*>(me.region.data() + L.local_off);
    float4* g = P.grad[rank];
    for (int step = 0; step < steps; ++step) {
      // "backward pass": this block's part of the gradients, ACCUMULATED into the (cleared) buffer
      for (int64_t i4 = block; i4 < L.n4; i4 += blocks) {
        g[i4].x += value_of(rank, step, 4 * i4); g[i4].y += value_of(rank, step, 4 * i4 + 1);
        g[i4].z += value_of(rank, step, 4 * i4 + 2); g[i4].w += value_of(rank, step, 4 * i4 + 3);
      }
      me.stream->arrive_and_wait();
      jitter();
      // ---- exchange_kernel
      unsigned long long epoch = loc->epoch + 1;      // advanced by this rank's Adam of the previous step, ordered by the stream barrier
      if (block == 0)
        for (int r = 0; r < world; ++r) st_release(P.flags[r] + rank, epoch);                    // (a) my gradients are final
      if (g_mutation != 2) wait_epochs_host(P.flags[rank], world, epoch);                        // (b)
      const int64_t b4 = shard_begin4(L, rank);
      for (int64_t i = block; i < shard_len4(L, rank); i += blocks) {                            // (c) pull, add, push
        const float4 s = reduce_ranks(L, P, b4 + i, LoadPeer());
        for (int p = 0; p < world; ++p) P.red[p][b4 + i] = s;
      }
      if (last_block_host(&loc->done[0], blocks))
        for (int r = 0; r < world; ++r) st_release(P.flags[r] + MAX_WORLD + rank, epoch);        // (d) shard `rank` has landed
      me.stream->arrive_and_wait();
      jitter();
      // ---- adam_kernel: wait for the shards, read the reduced gradients, clear my own gradient buffer
      if (g_mutation != 1) wait_epochs_host(P.flags[rank] + MAX_WORLD, world, epoch);
      for (int64_t i4 = block; i4 < L.n4; i4 += blocks) {
        const float* v = reinterpret_cast<const float*>(P.red[rank] + i4);
        for (int c = 0; c < 4; ++c) {
          float want = 0.f;
          for (int r = 0; r < world; ++r) want += value_of(r, step, 4 * i4 + c);
          if (v[c] != want) errors.fetch_add(1);
        }
        g[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (last_block_host(&loc->done[2], blocks)) loc->epoch = epoch;
      me.stream->arrive_and_wait();
    }
  };
  std::vector<std::thread> threads;
  for (int r = 0; r < world; ++r)
    for (int b = 0; b < blocks; ++b) threads.emplace_back(worker, r, b);
  for (auto& t : threads) t.join();
  printf("{\"world\": %d, \"n_floats\": %lld, \"steps\": %d, \"blocks_per_rank\": %d, \"errors\": %d}\n", world, n, steps, blocks, errors.load());
  return errors.load() == 0 ? 0 : 1;
}
