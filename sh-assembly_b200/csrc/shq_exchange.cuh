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
// counts[s * cps + c] = its length; source s starts at seg_start[s].  One CTA per source (block scan).
__global__ void __launch_bounds__(kScanThreads) seg_table_kernel(const uint32_t* __restrict__ counts, const uint64_t* __restrict__ seg_start, int world,
                                                                int cps, uint64_t* __restrict__ seg_beg) {
  const int s = blockIdx.x;
  const int per = (cps + kScanThreads - 1) / kScanThreads;
  const int c0 = threadIdx.x * per;
  uint64_t sum = 0;
  for (int j = 0; j < per; j++) if (c0 + j < cps) sum += counts[(size_t)s * cps + c0 + j];
  uint64_t tot;
  uint64_t o = seg_start[s] + block_exclusive_scan<SumU64>(sum, &tot);
  for (int j = 0; j < per; j++) {
    if (c0 + j < cps) { seg_beg[(size_t)s * (cps + 1) + c0 + j] = o; o += counts[(size_t)s * cps + c0 + j]; }
  }
  if (threadIdx.x == 0) seg_beg[(size_t)s * (cps + 1) + cps] = seg_start[s] + tot;
}

// ---- pull variant: the owner reads its coarse bins straight out of every source's binned array over NVLink ----
// fine[s * nfg + f] = words source s holds in GLOBAL fine bin f (nfg = world * F fine bins over the whole key space,
// 2^fb fine bins per coarse bin).  One CTA per source s:
//   seg_beg[s * (cps + 1) + c] = offset of MY coarse bin c inside source s's binned array (prefix over all its bins)
// and, CTA-independent (grid-stride over my F fine bins): fine_tot[f] = sum over s of fine[s][rank * F + f].
__global__ void __launch_bounds__(kScanThreads) pull_table_kernel(const uint32_t* __restrict__ fine, int world, int rank, int cps, int fb, int nfg,
                                                                 uint64_t* __restrict__ seg_beg, uint32_t* __restrict__ fine_tot) {
  const int s = blockIdx.x;
  const int F = cps << fb;                                  // my fine bins
  const uint32_t* row = fine + (size_t)s * nfg;
  // prefix of everything before my first fine bin
  const int before = rank * F;
  uint64_t sum = 0;
  for (int i = threadIdx.x; i < before; i += kScanThreads) sum += row[i];
  uint64_t tot;
  block_exclusive_scan<SumU64>(sum, &tot);
  const uint64_t start = tot;
  // my coarse bins: prefix over groups of 2^fb fine bins
  const int per = (cps + kScanThreads - 1) / kScanThreads;
  const int c0 = threadIdx.x * per;
  uint64_t mine = 0;
  for (int j = 0; j < per; j++) if (c0 + j < cps) for (int k = 0; k < (1 << fb); k++) mine += row[before + ((c0 + j) << fb) + k];
  uint64_t tot2;
  uint64_t o = start + block_exclusive_scan<SumU64>(mine, &tot2);
  for (int j = 0; j < per; j++) {
    if (c0 + j < cps) {
      seg_beg[(size_t)s * (cps + 1) + c0 + j] = o;
      for (int k = 0; k < (1 << fb); k++) o += row[before + ((c0 + j) << fb) + k];
    }
  }
  if (threadIdx.x == 0) seg_beg[(size_t)s * (cps + 1) + cps] = start + tot2;
  for (int f = blockIdx.x * kScanThreads + threadIdx.x; f < F; f += gridDim.x * kScanThreads) {
    uint32_t t = 0;
    for (int q = 0; q < world; q++) t += fine[(size_t)q * nfg + before + f];
    fine_tot[f] = t;
  }
}
// ---- layout metadata: what a source publishes about its binned array, whatever way it was laid out ----------------
// meta[b] = first word of GLOBAL coarse bin b inside the source's array, meta[ncoarse + b] = one past its last word.
// Fixed-capacity bins (one hashing pass): begin = b * capbin, end = the scatter cursor.  Exact bins: begin = prefix, end =
// next prefix.  *total += words held.
__global__ void layout_meta_kernel(const uint32_t* __restrict__ bin_off, const uint32_t* __restrict__ bin_end, int ncoarse,
                                   uint32_t* __restrict__ meta, unsigned long long* __restrict__ total) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t len = 0;
  if (b < ncoarse) {
    const uint32_t beg = bin_off[b];
    uint32_t end = bin_off[b + 1];
    if (bin_end) { const uint32_t e = bin_end[b]; end = e < end ? e : end; }
    meta[b] = beg; meta[ncoarse + b] = end; len = end - beg;
  }
  for (int d = 16; d; d >>= 1) len += __shfl_down_sync(0xffffffffu, len, d);
  if ((threadIdx.x & 31) == 0 && len) atomicAdd(total, (unsigned long long)len);
}
// Owner side: seg_beg / seg_end[s * cps + c] = range of MY coarse bin c inside source s's array, from the all-gathered
// metadata (row s = meta of source s, row stride = meta_stride words); *total += words to pull.
__global__ void pull_layout_kernel(const uint32_t* __restrict__ all_meta, size_t meta_stride, int world, int rank, int cps, int ncoarse,
                                   uint64_t* __restrict__ seg_beg, uint64_t* __restrict__ seg_end, unsigned long long* __restrict__ total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t len = 0;
  if (i < world * cps) {
    const int s = i / cps, c = i % cps;
    const uint32_t* row = all_meta + (size_t)s * meta_stride;
    const uint32_t beg = row[rank * cps + c], end = row[ncoarse + rank * cps + c];
    seg_beg[i] = beg; seg_end[i] = end; len = end - beg;
  }
  for (int d = 16; d; d >>= 1) len += __shfl_down_sync(0xffffffffu, len, d);
  if ((threadIdx.x & 31) == 0 && len) atomicAdd(total, (unsigned long long)len);
}

// coarse[b] = sum of the 2^fb fine bins of global coarse bin b
__global__ void coarse_from_fine_kernel(const uint32_t* __restrict__ fine, int ncoarse, int fb, uint32_t* __restrict__ coarse) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ncoarse) return;
  uint32_t t = 0;
  for (int k = 0; k < (1 << fb); k++) t += fine[((size_t)b << fb) + k];
  coarse[b] = t;
}

constexpr int kSplitThreads = 256;
constexpr int kSplitItems = 16;
constexpr int kSplitTile = kSplitThreads * kSplitItems;
constexpr int kSplitMaxFine = 128;          // fb <= 7 on one GPU (split_bins_kernel); the exchange's resplit_kernel uses fb <= 5

// grid = cps * parts; CTA (c, part) walks every parts-th tile of the sub-segments of coarse bin c.
// COUNT: g_fine[c << fb | sub] += occurrences.   !COUNT: g_fine holds running cursors (start = fine bin offset);
// the tile is counting-sorted by sub in shared memory and written out linearly.
// seg_end == nullptr: segments are contiguous, seg_beg has cps + 1 entries per source (end = next begin); otherwise
// seg_beg / seg_end hold cps entries per source (pull_layout_kernel).  capfine > 0: the fine bins have a fixed capacity --
// a reservation that does not fit raises *ovf and lands in a scratch area at slack_base (one tile long).
template <bool COUNT>
__global__ void __launch_bounds__(kSplitThreads) resplit_kernel(const uint32_t* const* __restrict__ bases, const uint64_t* __restrict__ seg_beg, int world, int cps,
                                                               int parts, int lbc, int fb, uint32_t* __restrict__ g_fine, uint32_t* __restrict__ out,
                                                               const uint64_t* __restrict__ seg_end = nullptr, uint32_t capfine = 0,
                                                               uint32_t slack_base = 0, int* __restrict__ ovf = nullptr) {
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
    const uint64_t beg = seg_end ? seg_beg[(size_t)s * cps + c] : seg_beg[(size_t)s * (cps + 1) + c];
    const uint64_t end = seg_end ? seg_end[(size_t)s * cps + c] : seg_beg[(size_t)s * (cps + 1) + c + 1];
    const uint32_t* __restrict__ recv = bases[s];      // the receive buffer, or source s's own binned array (peer memory)
    for (uint64_t t0 = beg; t0 < end; t0 += kSplitTile, tile_no++) {
      if ((int)(tile_no % (uint32_t)parts) != part) continue;
      uint32_t w[kSplitItems], sr[kSplitItems];     // word, sub << 16 | rank inside (tile, sub)
#pragma unroll
      for (int i = 0; i < kSplitItems; i++) {        // all loads of the tile are in flight before the first is used
        uint64_t idx = t0 + (uint64_t)i * kSplitThreads + threadIdx.x;
        w[i] = idx < end ? recv[idx] : 0xFFFFFFFFu;
      }
#pragma unroll
      for (int i = 0; i < kSplitItems; i++) {
        uint64_t idx = t0 + (uint64_t)i * kSplitThreads + threadIdx.x;
        const bool live = idx < end;
        const uint32_t sub = (w[i] & cmask) >> lbf;
        // rank inside (tile, sub) straight from the shared-memory counter (a few-way conflict per warp)
        if (COUNT) { if (live) atomicAdd(&s_cnt[sub], 1u); }
        else sr[i] = live ? ((sub << 16) | atomicAdd(&s_cnt[sub], 1u)) : 0xFFFFFFFFu;
      }
      if (COUNT) continue;     // counts accumulate over all tiles of this CTA (one barrier at the end)
      __syncthreads();
      if (threadIdx.x < 32) {  // nfine <= 128: lane j owns fine bins j, j + 32, ... -- one global reservation per bin, offsets by warp scans
        uint32_t carry = 0;
        for (int q0 = 0; q0 < nfine; q0 += 32) {          // one pass for fb <= 5
          const int b = q0 + lane;
          const uint32_t n = b < nfine ? s_cnt[b] : 0;
          const uint32_t f = ((uint32_t)c << fb) | (uint32_t)b;
          uint32_t bs = n ? atomicAdd(&g_fine[f], n) : 0;      // in flight while the scan runs
          uint32_t inc = n;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
          const uint32_t o = carry + inc - n;
          if (b < nfine) {
            if (capfine && n && (uint64_t)bs + n > (uint64_t)(f + 1) * capfine) { *ovf = 1; bs = slack_base + o; }
            s_off[b] = o;
            s_delta[b] = bs - o;
            s_cnt[b] = 0;
          }
          carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_off[nfine] = carry;
      }
      __syncthreads();
      const uint32_t total = s_off[nfine];
#pragma unroll
      for (int i = 0; i < kSplitItems; i++) {
        const uint32_t sub = sr[i] >> 16;
        if (sr[i] != 0xFFFFFFFFu) {
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
    __syncthreads();
    if (threadIdx.x < nfine) { uint32_t v = s_cnt[threadIdx.x]; if (v) atomicAdd(&g_fine[((size_t)c << fb) | threadIdx.x], v); }      // nfine <= 128 < kSplitThreads
  }
}

// ---- second binning level on ONE GPU ------------------------------------------------------------------------------
// A device batch is binned by the hashing pass into <= 4096 coarse bins (roll_scatter_kernel keeps three words per
// bin in shared memory).  When a coarse bin would hold more distinct keys than the aggregator's table -- large batches
// of a large genome, where almost every k-mer of a batch is distinct -- every coarse bin is split once more into 2^fb
// fine bins: a counting sort of 4096-word tiles in shared memory, one global reservation per (tile, fine bin), the tile
// written out linearly.  Both levels use fixed-capacity bins (ntHash keys are uniform); a reservation that does not fit
// raises *ovf and lands in a scratch area, and the host falls back to the coarse bins.
// in:  coarse bin c = [bin_off[c], bin_end[c]);  word = chunk << lb | low lb key bits
// out: fine bin (c << fb | sub) = [f * capfine, fine_cursor[f]);  word = chunk << (lb - fb) | low (lb - fb) key bits
__global__ void __launch_bounds__(kSplitThreads) split_bins_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ bin_off,
                                                                   const uint32_t* __restrict__ bin_end, int parts, int lb, int fb,
                                                                   uint32_t* __restrict__ fine_cursor, uint32_t capfine, uint32_t slack_base,
                                                                   int* __restrict__ ovf, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_cnt[kSplitMaxFine], s_off[kSplitMaxFine + 1], s_delta[kSplitMaxFine];
  __shared__ uint32_t s_sorted[kSplitTile];
  __shared__ uint8_t s_sub[kSplitTile];
  const int c = blockIdx.x / parts, part = blockIdx.x % parts;
  const int nfine = 1 << fb;
  const int lbf = lb - fb;
  const uint32_t cmask = lb >= 32 ? 0xFFFFFFFFu : ((1u << lb) - 1), fmask = (1u << lbf) - 1;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < nfine) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t beg = bin_off[c];
  uint32_t end = bin_off[c + 1];
  { const uint32_t e = bin_end[c]; end = e < end ? e : end; }
  for (uint32_t t0 = beg + (uint32_t)part * kSplitTile; t0 < end; t0 += (uint32_t)parts * kSplitTile) {
    uint32_t w[kSplitItems], sr[kSplitItems];     // word, sub << 16 | rank inside (tile, sub)
#pragma unroll
    for (int i = 0; i < kSplitItems; i++) {        // all loads of the tile are in flight before the first is used
      const uint32_t idx = t0 + (uint32_t)i * kSplitThreads + threadIdx.x;
      w[i] = idx < end ? __ldg(&in[idx]) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 0; i < kSplitItems; i++) {
      const uint32_t idx = t0 + (uint32_t)i * kSplitThreads + threadIdx.x;
      const uint32_t sub = (w[i] & cmask) >> lbf;
      sr[i] = idx < end ? ((sub << 16) | atomicAdd(&s_cnt[sub], 1u)) : 0xFFFFFFFFu;
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // nfine <= 128: lane j owns fine bins j, j + 32, ... -- one global reservation per bin, offsets by warp scans
      uint32_t carry = 0;
      for (int q0 = 0; q0 < nfine; q0 += 32) {          // one pass for fb <= 5
        const int b = q0 + lane;
        const uint32_t n = b < nfine ? s_cnt[b] : 0;
        const uint32_t f = ((uint32_t)c << fb) | (uint32_t)b;
        uint32_t bs = n ? atomicAdd(&fine_cursor[f], n) : 0;      // in flight while the scan runs
        uint32_t inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        const uint32_t o = carry + inc - n;
        if (b < nfine) {
          if (n && (uint64_t)bs + n > (uint64_t)(f + 1) * capfine) { *ovf = 1; bs = slack_base + o; }
          s_off[b] = o;
          s_delta[b] = bs - o;
          s_cnt[b] = 0;
        }
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) s_off[nfine] = carry;
    }
    __syncthreads();
    const uint32_t total = s_off[nfine];
#pragma unroll
    for (int i = 0; i < kSplitItems; i++) {
      if (sr[i] != 0xFFFFFFFFu) {
        const uint32_t sub = sr[i] >> 16;
        const uint32_t dst = s_off[sub] + (sr[i] & 0xFFFFu);
        s_sorted[dst] = (w[i] & fmask) | ((lb >= 32 ? 0u : (w[i] >> lb)) << lbf);
        s_sub[dst] = (uint8_t)sub;
      }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += kSplitThreads) out[s_delta[s_sub[i]] + i] = s_sorted[i];
    __syncthreads();
  }
}
__global__ void fine_bin_offsets_kernel(uint32_t* __restrict__ bin_off, uint32_t* __restrict__ cursor, uint32_t n_bins, uint32_t capfine) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p <= n_bins) { bin_off[p] = p * capfine; if (p < n_bins) cursor[p] = p * capfine; }
}

}  // namespace shq
