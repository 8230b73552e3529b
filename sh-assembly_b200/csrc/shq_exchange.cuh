// shq_exchange.cuh -- second level of the multi-GPU binning.
//
// With N shards a source GPU cannot bin straight into N x 4096 fine bins (the staged scatter of shq_roll.cuh keeps
// three words per bin in shared memory and writes runs of (round size / bins) words).  So the exchange is two-level:
//   1. every rank bins its k-mers over the WHOLE key space into <= 4096 COARSE bins (bins of one owner are
//      contiguous, so the all-to-all ships plain slices of the binned array, 4 bytes per k-mer);
//   2. the owner receives, per source, its coarse bins in order, and splits every coarse bin into 2^fb FINE bins
//      (these kernels) -- a 4-bit counting sort per tile, words of one fine bin written next to each other;
//   3. the fine bins go through the single-GPU aggregator / merge unchanged.
// Received word = low lbc key bits (lbc = hb - coarse_bits); fine word = low (lbc - fb) bits.
#pragma once
#include "shq_common.cuh"
#include "shq_scan.cuh"

namespace shq {

// seg_beg[s * (cps + 1) + c] = first word of coarse bin c from source s inside the receive buffer;
// counts[s * cps + c] = its length; source s starts at seg_start[s].  One thread per source.
__global__ void seg_table_kernel(const uint32_t* __restrict__ counts, const uint64_t* __restrict__ seg_start, int world, int cps,
                                 uint64_t* __restrict__ seg_beg) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= world) return;
  uint64_t o = seg_start[s];
  for (int c = 0; c < cps; c++) { seg_beg[(size_t)s * (cps + 1) + c] = o; o += counts[(size_t)s * cps + c]; }
  seg_beg[(size_t)s * (cps + 1) + cps] = o;
}

constexpr int kSplitThreads = 256;
constexpr int kSplitItems = 8;
constexpr int kSplitTile = kSplitThreads * kSplitItems;
constexpr int kSplitMaxFine = 64;           // fb <= 6

// grid = cps * parts; CTA (c, part) walks every parts-th tile of the sub-segments of coarse bin c.
// COUNT: g_fine[c << fb | sub] += occurrences.   !COUNT: g_fine holds running cursors (start = fine bin offset);
// the tile is counting-sorted by sub in shared memory and written out linearly.
template <bool COUNT>
__global__ void __launch_bounds__(kSplitThreads) resplit_kernel(const uint32_t* __restrict__ recv, const uint64_t* __restrict__ seg_beg, int world, int cps,
                                                               int parts, int lbc, int fb, uint32_t* __restrict__ g_fine, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_cnt[kSplitMaxFine], s_off[kSplitMaxFine + 1], s_delta[kSplitMaxFine];
  __shared__ uint32_t s_sorted[kSplitTile];
  __shared__ uint8_t s_sub[kSplitTile];
  const int c = blockIdx.x / parts, part = blockIdx.x % parts;
  const int nfine = 1 << fb;
  const int lbf = lbc - fb;
  const uint32_t cmask = lbc >= 32 ? 0xFFFFFFFFu : ((1u << lbc) - 1), fmask = (1u << lbf) - 1;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < nfine) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t tile_no = 0;
  for (int s = 0; s < world; s++) {
    const uint64_t beg = seg_beg[(size_t)s * (cps + 1) + c], end = seg_beg[(size_t)s * (cps + 1) + c + 1];
    for (uint64_t t0 = beg; t0 < end; t0 += kSplitTile, tile_no++) {
      if ((int)(tile_no % (uint32_t)parts) != part) continue;
      uint32_t w[kSplitItems], sr[kSplitItems];     // word, sub << 16 | rank inside (tile, sub)
#pragma unroll
      for (int i = 0; i < kSplitItems; i++) {
        uint64_t idx = t0 + (uint64_t)i * kSplitThreads + threadIdx.x;
        const bool live = idx < end;
        w[i] = live ? __ldg(&recv[idx]) : 0;
        const uint32_t sub = live ? ((w[i] & cmask) >> lbf) : 0xFFu;
        // lanes with the same sub: one shared-memory atomic per (warp, sub), ranks from the lane order
        const uint32_t peers = __match_any_sync(0xffffffffu, sub);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (live && lane == leader) base = atomicAdd(&s_cnt[sub], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        sr[i] = (sub << 16) | (base + (uint32_t)__popc(peers & ((1u << lane) - 1)));
      }
      __syncthreads();
      if (COUNT) continue;     // counts accumulate over all tiles of this CTA
      if (threadIdx.x == 0) {
        uint32_t o = 0;
        for (int j = 0; j < nfine; j++) {
          uint32_t n = s_cnt[j];
          s_off[j] = o;
          s_delta[j] = (n ? atomicAdd(&g_fine[((size_t)c << fb) | j], n) : 0) - o;
          o += n; s_cnt[j] = 0;
        }
        s_off[nfine] = o;
      }
      __syncthreads();
      const uint32_t total = s_off[nfine];
#pragma unroll
      for (int i = 0; i < kSplitItems; i++) {
        const uint32_t sub = sr[i] >> 16;
        if (sub != 0xFFu) {
          const uint32_t dst = s_off[sub] + (sr[i] & 0xFFFFu);
          s_sorted[dst] = w[i] & fmask;
          s_sub[dst] = (uint8_t)sub;
        }
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < total; i += kSplitThreads) out[s_delta[s_sub[i]] + i] = s_sorted[i];
      __syncthreads();
    }
  }
  if (COUNT) {
    if (threadIdx.x < nfine) { uint32_t v = s_cnt[threadIdx.x]; if (v) atomicAdd(&g_fine[((size_t)c << fb) | threadIdx.x], v); }
  }
}

}  // namespace shq
