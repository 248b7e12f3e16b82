// Host build of the chained kernel's schedule arithmetic (gmvae_b200/csrc/chain_sched.cuh, the functions the device code calls):
// simulates every walker and every CTA of a walker over one job and checks that the tile space is covered exactly once.
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../../gmvae_b200/csrc/chain_sched.cuh"

using namespace gmvae::tc;

template <int CL>
static int check_gemm_job(const ChainJob& J, int G) {
  const int tiles_m = (J.M + SCHED_BLOCK_M - 1) / SCHED_BLOCK_M;
  const int real_blocks = tiles_m;                       // row blocks that exist
  std::map<std::tuple<int, int, int>, int> seen;         // (z, mb, n0) -> count
  for (int c = 0; c < G; ++c) {
    int first = -1, stride = -1;
    // every CTA of a walker must take the same steps
    if (!chain_walk<1>(J, c, G, first, stride)) continue;
    if (stride <= 0 || first < 0 || first >= stride) return 10;
    for (int l = first; l < J.walk_total; l += stride) {
      int zs[4], mbs[4], n0s[4];
      for (int r = 0; r < CL; ++r) chain_tile<CL>(J, l, r, zs[r], mbs[r], n0s[r]);
      for (int r = 0; r < CL; ++r) {
        if (zs[r] != zs[0]) return 11;                   // one k-split per step: the pairs of a cluster run the same number of k-blocks
        if (zs[r] < 0 || zs[r] >= J.num_splits) return 12;
        if (n0s[r] < 0 || n0s[r] % J.block_n != 0 || n0s[r] / J.block_n >= J.tiles_n) return 13;
      }
      if (CL >= 2) {
        for (int h = 0; h < CL / 2; ++h) {               // the two CTAs of a pair: same columns, row blocks 2 pm and 2 pm + 1
          if (n0s[2 * h] != n0s[2 * h + 1]) return 14;
          if (mbs[2 * h] % 2 != 0 || mbs[2 * h + 1] != mbs[2 * h] + 1) return 15;
        }
      }
      if (CL == 4 && J.share) {                          // shared A rows: same row blocks, different n-tiles
        if (mbs[0] != mbs[2] || mbs[1] != mbs[3]) return 16;
        if (n0s[0] == n0s[2]) return 17;
      }
      for (int r = 0; r < CL; ++r) {
        const int limit = CL >= 2 ? 2 * ((tiles_m + 1) / 2) : tiles_m;
        if (mbs[r] >= limit) {                           // phantom pair tile: only in quad mode, only beyond the job's own counters
          if (CL != 4) return 18;
          if (mbs[r] >= chain_job_counters(J.M, CL)) return 19;
          continue;
        }
        seen[std::make_tuple(zs[r], mbs[r], n0s[r])]++;
      }
    }
  }
  // every (k-split, existing-or-phantom-half row block of a pair tile, n-tile) exactly once
  const int mb_count = CL >= 2 ? 2 * ((tiles_m + 1) / 2) : tiles_m;
  (void)real_blocks;
  for (int z = 0; z < J.num_splits; ++z)
    for (int mb = 0; mb < mb_count; ++mb)
      for (int nt = 0; nt < J.tiles_n; ++nt) {
        auto it = seen.find(std::make_tuple(z, mb, nt * J.block_n));
        if (it == seen.end()) return 20;
        if (it->second != 1) return 21;
      }
  if ((int)seen.size() != J.num_splits * mb_count * J.tiles_n) return 22;
  // the k-splits cover the k-blocks
  if (J.kb_per_split * J.num_splits < 1) return 23;
  return 0;
}

extern "C" int host_sched_check_gemm(int M, int N, int block_n, int split_k, int kb_total, int a_mn, int cl, int G, int tile_base, int wfirst,
                                     int wcount) {
  ChainJob J;
  std::memset(&J, 0, sizeof(J));
  chain_job_geometry(J, M, N, block_n, split_k, kb_total, a_mn != 0, cl);
  J.tile_base = tile_base; J.wfirst = wfirst; J.wcount = wcount;
  if ((J.kb_per_split * (J.num_splits - 1) >= kb_total) || J.kb_per_split * J.num_splits < kb_total) return 30;   // no empty split, all k-blocks covered
  return cl == 1 ? check_gemm_job<1>(J, G) : cl == 2 ? check_gemm_job<2>(J, G) : check_gemm_job<4>(J, G);
}

// Row jobs are dealt to single CTAs: G walkers of `cl` CTAs each.
extern "C" int host_sched_check_rows(int total_tiles, int cl, int G, int tile_base, int wfirst, int wcount) {
  ChainJob J;
  std::memset(&J, 0, sizeof(J));
  J.total_tiles = total_tiles; J.tile_base = tile_base; J.wfirst = wfirst; J.wcount = wcount;
  std::vector<int> seen(total_tiles, 0);
  for (int cta = 0; cta < G * cl; ++cta) {
    int first, stride;
    const bool ok = cl == 1 ? chain_walk<1>(J, cta, G * cl, first, stride) : cl == 2 ? chain_walk<2>(J, cta, G * cl, first, stride)
                                                                                  : chain_walk<4>(J, cta, G * cl, first, stride);
    if (!ok) continue;
    for (int l = first; l < total_tiles; l += stride) seen[l]++;
  }
  for (int l = 0; l < total_tiles; ++l)
    if (seen[l] != 1) return 40;
  return 0;
}
