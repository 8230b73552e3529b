// shq_aggregate.cuh -- K3: turning a batch of binned k-mer keys into a sorted run of distinct
// (key, count[, first chunk]) and merging that run into the resident sorted state.
//
// Replaces the per-key insert1_advance loop (reference cqf/gqf.c:1614-1915) and its locks /
// local-QF spill (cqf/CQF_mt.h:637-645): because the CQF image is a pure function of the
// (key,count) multiset, the GPU never shifts slots; it aggregates and later materialises.
#pragma once
#include "shq_common.cuh"
#include "shq_scan.cuh"

namespace shq {

// ----------------------------------------------------- per-bin aggregation ---
constexpr int kAggThreads = 512;
constexpr int kAggTable = 8192;          // open-addressing slots per CTA
constexpr int kAggMaxDistinct = 4096;    // per round (load factor <= 0.5)
constexpr uint32_t kAggEmpty = 0xFFFFFFFFu;
constexpr size_t kAggSmemBytes = (size_t)kAggTable * (4 + 4 + 2) + (size_t)kAggMaxDistinct * 4;

__device__ __forceinline__ uint32_t agg_hash(uint32_t k) { return (k * 2654435761u) >> (32 - 13); }

// One CTA per bin.  packed[bin_off[p] .. bin_off[p+1]) holds (key_low | chunk << lb).
// Occurrences whose chunk is outside [chunk_lo, chunk_hi] are ignored (deNoise boundary replay).
// Output (sorted by key) goes to out_*[bin_off[p] + i]; n_out[p] = number of distinct keys,
// occ_out[p] = number of occurrences consumed.
// If a bin holds more than kAggMaxDistinct distinct keys the CTA re-streams it in 2^k sub-ranges.
__global__ void __launch_bounds__(kAggThreads) bin_aggregate_kernel(
    const uint32_t* __restrict__ packed, const uint32_t* __restrict__ bin_off, int lb, uint64_t key_lo,
    uint32_t chunk_lo, uint32_t chunk_hi, uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts,
    uint16_t* __restrict__ out_chunk, uint32_t* __restrict__ n_out, uint32_t* __restrict__ occ_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* t_key = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* t_cnt = t_key + kAggTable;
  uint32_t* s_sort = t_cnt + kAggTable;
  uint16_t* t_chk = reinterpret_cast<uint16_t*>(s_sort + kAggMaxDistinct);
  __shared__ uint32_t s_ndist, s_nocc, s_overflow, s_nsorted;

  const uint32_t p = blockIdx.x;
  const uint32_t beg = bin_off[p], end = bin_off[p + 1];
  const uint32_t lowmask = lb >= 32 ? 0xFFFFFFFFu : ((1u << lb) - 1);
  uint32_t written = 0, occ_total = 0;
  int sub_bits = 0;

restart:
  written = 0; occ_total = 0;
  for (uint32_t r = 0; r < (1u << sub_bits); r++) {
    for (int i = threadIdx.x; i < kAggTable; i += blockDim.x) { t_key[i] = kAggEmpty; t_cnt[i] = 0; t_chk[i] = 0xFFFF; }
    if (threadIdx.x == 0) { s_ndist = 0; s_nocc = 0; s_overflow = 0; }
    __syncthreads();
    uint32_t my_occ = 0;
    for (uint32_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      uint32_t w = packed[i];
      uint32_t kl = w & lowmask, ch = lb >= 32 ? 0 : (w >> lb);
      if (ch < chunk_lo || ch > chunk_hi) continue;
      if (sub_bits && (kl >> (lb - sub_bits)) != r) continue;
      my_occ++;
      uint32_t h = agg_hash(kl);
      bool placed = false;
      for (int probe = 0; probe < kAggTable; probe++) {
        uint32_t cur = t_key[h];
        if (cur == kl) { placed = true; break; }
        if (cur == kAggEmpty) {
          if (*reinterpret_cast<volatile uint32_t*>(&s_overflow)) break;
          uint32_t old = atomicCAS(&t_key[h], kAggEmpty, kl);
          if (old == kAggEmpty) {
            if (atomicAdd(&s_ndist, 1u) >= (uint32_t)kAggMaxDistinct) s_overflow = 1;
            placed = true; break;
          }
          if (old == kl) { placed = true; break; }
        }
        h = (h + 1) & (kAggTable - 1);
      }
      if (placed) {
        atomicAdd(&t_cnt[h], 1u);
        if (out_chunk) {  // 16-bit atomic min of the first chunk through the containing 32-bit word
          unsigned int* wp = reinterpret_cast<unsigned int*>(t_chk) + (h >> 1);
          unsigned int sh = (h & 1) * 16;
          unsigned int old = *reinterpret_cast<volatile unsigned int*>(wp);
          while (((old >> sh) & 0xFFFFu) > ch) {
            unsigned int nw = (old & ~(0xFFFFu << sh)) | (ch << sh);
            unsigned int prev = atomicCAS(wp, old, nw);
            if (prev == old) break;
            old = prev;
          }
        }
      }
    }
    __syncthreads();
    if (*reinterpret_cast<volatile uint32_t*>(&s_overflow)) { sub_bits++; __syncthreads(); goto restart; }
    // gather distinct keys, sort them
    if (threadIdx.x == 0) s_nsorted = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < kAggTable; i += blockDim.x) {
      uint32_t k = t_key[i];
      if (k != kAggEmpty) s_sort[atomicAdd(&s_nsorted, 1u)] = k;
    }
    __syncthreads();
    const uint32_t nd = s_nsorted;
    uint32_t m = 1; while (m < nd) m <<= 1;
    for (uint32_t i = nd + threadIdx.x; i < m; i += blockDim.x) s_sort[i] = kAggEmpty;
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
          uint32_t lo = (t / stride) * (stride << 1) + (t % stride);
          uint32_t hi = lo + stride;
          bool up = ((lo & size) == 0);
          uint32_t x = s_sort[lo], y = s_sort[hi];
          if ((x > y) == up) { s_sort[lo] = y; s_sort[hi] = x; }
        }
        __syncthreads();
      }
    }
    for (uint32_t i = threadIdx.x; i < nd; i += blockDim.x) {
      uint32_t kl = s_sort[i];
      uint32_t h = agg_hash(kl);
      while (t_key[h] != kl) h = (h + 1) & (kAggTable - 1);
      size_t o = (size_t)beg + written + i;
      out_keys[o] = key_lo + (((uint64_t)p << lb) | kl);
      out_counts[o] = t_cnt[h];
      if (out_chunk) out_chunk[o] = t_chk[h];
    }
    atomicAdd(&s_nocc, my_occ);
    __syncthreads();
    written += nd;
    occ_total += s_nocc;
    __syncthreads();
  }
  if (threadIdx.x == 0) { n_out[p] = written; occ_out[p] = occ_total; }
}

// Dense gather: bin p's n_out[p] entries move from [bin_off[p], ...) to [dense_off[p], ...).
__global__ void __launch_bounds__(256) compact_bins_kernel(const uint32_t* __restrict__ bin_off, const uint32_t* __restrict__ dense_off,
                                                          const uint32_t* __restrict__ n_out, const uint64_t* __restrict__ in_keys,
                                                          const uint32_t* __restrict__ in_counts, const uint16_t* __restrict__ in_chunk,
                                                          uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts,
                                                          uint16_t* __restrict__ out_chunk) {
  uint32_t p = blockIdx.x;
  uint32_t n = n_out[p];
  size_t src = bin_off[p], dst = dense_off[p];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    out_keys[dst + i] = in_keys[src + i];
    out_counts[dst + i] = in_counts[src + i];
    if (in_chunk) out_chunk[dst + i] = in_chunk[src + i];
  }
}

// ----------------------------------------------------------------- merge ----
// State A (na sorted distinct keys) and batch run B (nb sorted distinct keys) -> merged state.
// Merge-path tiles of kMergeTile merged elements; ties order A before B.
constexpr int kMergeThreads = 256;
constexpr int kMergeItems = 8;
constexpr int kMergeTile = kMergeThreads * kMergeItems;
constexpr int kMergeHistMax = 1024;

__device__ __forceinline__ size_t merge_path_search(const uint64_t* A, size_t na, const uint64_t* B, size_t nb, size_t diag) {
  size_t lo = diag > nb ? diag - nb : 0, hi = diag < na ? diag : na;
  while (lo < hi) {
    size_t mid = lo + ((hi - lo) >> 1);
    if (A[mid] <= B[diag - mid - 1]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void merge_partition_kernel(const uint64_t* __restrict__ A, size_t na, const uint64_t* __restrict__ B, size_t nb,
                                       uint64_t* __restrict__ a_split, size_t ntiles) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  size_t diag = t * kMergeTile;
  if (diag > na + nb) diag = na + nb;
  a_split[t] = merge_path_search(A, na, B, nb, diag);
}

// WRITE=false: heads[t] = number of output entries of tile t, and (if chunk_hist) the histogram of
//              first-chunk ids over batch keys that are NOT in the state ("new" keys).
// WRITE=true : emits merged (key,count) at out_off[t] + rank.  count overflow sets *err.
template <bool WRITE>
__global__ void __launch_bounds__(kMergeThreads) merge_tiles_kernel(
    const uint64_t* __restrict__ A, const uint32_t* __restrict__ Ac, size_t na, const uint64_t* __restrict__ B,
    const uint32_t* __restrict__ Bc, const uint16_t* __restrict__ Bchunk, size_t nb, const uint64_t* __restrict__ a_split,
    uint32_t* __restrict__ heads, const uint64_t* __restrict__ out_off, uint64_t* __restrict__ out_keys,
    uint32_t* __restrict__ out_counts, uint32_t* __restrict__ chunk_hist, int* __restrict__ err) {
  __shared__ uint64_t s_a[kMergeTile + 2];
  __shared__ uint64_t s_b[kMergeTile + 2];
  __shared__ uint32_t s_hist[kMergeHistMax];   // per-CTA histogram of first-chunk ids of new keys
  const bool use_hist = !WRITE && chunk_hist != nullptr;
  if (use_hist) for (int i = threadIdx.x; i < kMergeHistMax; i += blockDim.x) s_hist[i] = 0;
  const size_t t = blockIdx.x;
  size_t diag0 = t * kMergeTile, diag1 = diag0 + kMergeTile;
  if (diag1 > na + nb) diag1 = na + nb;
  const size_t a0 = a_split[t], a1 = a_split[t + 1];
  const size_t b0 = diag0 - a0, b1 = diag1 - a1;
  const int la = (int)(a1 - a0), lbn = (int)(b1 - b0);
  // s_a[0] = A[a0-1] (or sentinel), s_a[1..la], s_a[la+1] = unused; s_b[0..lbn), s_b[lbn] = B[b1] (or sentinel)
  const uint64_t kNone = ~0ULL;
  for (int i = threadIdx.x; i < la + 1; i += blockDim.x) s_a[i] = (i == 0) ? (a0 > 0 ? A[a0 - 1] : kNone) : A[a0 + i - 1];
  for (int i = threadIdx.x; i < lbn + 1; i += blockDim.x) s_b[i] = (i < lbn) ? B[b0 + i] : (b1 < nb ? B[b1] : kNone);
  __syncthreads();
  // thread-local merge path inside the tile
  const int total = la + lbn;
  int d0 = min(threadIdx.x * kMergeItems, total), d1 = min(d0 + kMergeItems, total);
  int lo = max(0, d0 - lbn), hi = min(d0, la);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (s_a[1 + mid] <= s_b[d0 - mid - 1]) lo = mid + 1; else hi = mid;
  }
  int ai = lo, bi = d0 - lo;
  uint32_t nheads = 0;
  uint64_t okey[kMergeItems]; uint32_t ocnt[kMergeItems];
  for (int d = d0; d < d1; d++) {
    bool takeA = (bi >= lbn) || (ai < la && s_a[1 + ai] <= s_b[bi]);
    if (takeA) {
      uint64_t k = s_a[1 + ai];
      uint64_t nextb = s_b[bi];  // bi <= lbn: s_b[lbn] is B[b1] or sentinel
      uint64_t c = Ac[a0 + ai];
      if (nextb == k) c += Bc[b0 + bi];
      if (WRITE) { if (c > 0xFFFFFFFFull) { *err = SHQ_ERR_COUNT_OVERFLOW; c = 0xFFFFFFFFull; } okey[nheads] = k; ocnt[nheads] = (uint32_t)c; }
      nheads++; ai++;
    } else {
      uint64_t k = s_b[bi];
      uint64_t preva = s_a[ai];  // s_a[ai] is A[a0+ai-1] (s_a[0] = A[a0-1] or sentinel)
      if (preva != k || (ai == 0 && a0 == 0)) {
        if (WRITE) { okey[nheads] = k; ocnt[nheads] = Bc[b0 + bi]; }
        else if (use_hist) {
          uint32_t ch = Bchunk ? Bchunk[b0 + bi] : 0;
          if (ch < (uint32_t)kMergeHistMax) atomicAdd(&s_hist[ch], 1u); else atomicAdd(&chunk_hist[ch], 1u);
        }
        nheads++;
      }
      bi++;
    }
  }
  uint32_t tot;
  uint32_t rank = block_exclusive_scan<SumU32>(nheads, &tot);
  if (!WRITE) {
    if (threadIdx.x == 0) heads[t] = tot;
    if (use_hist) {   // block_exclusive_scan ended with a barrier: every s_hist update is visible
      for (int i = threadIdx.x; i < kMergeHistMax; i += blockDim.x) { uint32_t v = s_hist[i]; if (v) atomicAdd(&chunk_hist[i], v); }
    }
  } else {
    size_t o = out_off[t] + rank;
    for (uint32_t i = 0; i < nheads; i++) { out_keys[o + i] = okey[i]; out_counts[o + i] = ocnt[i]; }
  }
}

// u32 -> u64 widening (scan input)
__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// ---------------------------------------------------------------- deNoise ---
// flags[k] = 1 if entry k survives the sweep: count != 1, or protected by the range-end rule.
__global__ void denoise_flag_kernel(const uint32_t* __restrict__ counts, uint64_t* __restrict__ flags, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = counts[i] != 1 ? 1 : 0;
}
__global__ void denoise_keep_kernel(const uint32_t* __restrict__ keep_idx, const uint32_t* __restrict__ n_keep,
                                    uint64_t* __restrict__ flags) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < *n_keep) flags[keep_idx[j]] = 1;
}
__global__ void denoise_compact_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                       const uint64_t* __restrict__ excl, size_t n, uint64_t total,
                                       uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t me = excl[i], nxt = (i + 1 < n) ? excl[i + 1] : total;
  if (nxt != me) { out_keys[me] = keys[i]; out_counts[me] = counts[i]; }
}

}  // namespace shq
