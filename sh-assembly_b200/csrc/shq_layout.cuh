// shq_layout.cuh -- K5 (canonical CQF image from the sorted state), the deNoise range walk (K4),
// the block-image decoder (load) and the batched rank/select query (K6).
//
// Image layout (reference cqf/gqf.c:63-86, 8-bit slots): block b = 89 bytes, packed:
//   [0] offset  [1..8] occupieds  [9..16] runends  [17..24] traveled  [25..88] 64 slot bytes.
#pragma once
#include "shq_common.cuh"
#include "shq_scan.cuh"

namespace shq {

constexpr int kBlockBytes = 89;

// ------------------------------------------------------------ positions -----
// Per entry the map f_k(e) = max(e + len, first ? q + len : -inf)  (MaxPlus); pos_k = first slot of entry k =
// (composition of all earlier maps)(0), raised to q_k for the first entry of a quotient.
// Two kernels around the tile-level scan; the maps are recomputed from (key, count) instead of being stored
// (32 instead of ~90 bytes of traffic per entry).
//   layout_reduce_kernel : tile_agg[t] = composition of the maps of tile t
//   (exclusive MaxPlus scan of tile_agg on the host side)
//   layout_apply_kernel  : pos[k] for every entry of tile t, seeded with the scanned tile_agg[t]
__device__ __forceinline__ MaxPlus layout_map(uint64_t key, uint64_t prev_key, bool has_prev, uint32_t count, uint64_t key_lo_q, bool* first, long long* q) {
  *q = (long long)((key >> 8) - key_lo_q);
  *first = !has_prev || ((prev_key >> 8) != (key >> 8));
  long long len = enc_len((uint32_t)(key & 0xFF), count);
  return MaxPlus{len, *first ? *q + len : kNegInf};
}
template <bool APPLY>
__global__ void __launch_bounds__(kScanThreads) layout_scan_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, size_t n,
                                                                  uint64_t key_lo_q, MaxPlus* __restrict__ tile_agg, uint64_t* __restrict__ pos,
                                                                  MaxPlus* __restrict__ grand_total, MaxPlus init /* everything before entry 0 (shards) */) {
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint64_t k[kScanItems + 1]; uint32_t c[kScanItems];
  k[0] = (base > 0 && base <= n) ? keys[base - 1] : 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) { k[i + 1] = base + i < n ? keys[base + i] : 0; c[i] = base + i < n ? counts[base + i] : 0; }
  MaxPlus acc = MaxPlusOp::id();
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) { bool first; long long q; acc = MaxPlusOp::op(acc, layout_map(k[i + 1], k[i], base + i > 0, c[i], key_lo_q, &first, &q)); }
  }
  MaxPlus tot;
  MaxPlus exc = block_exclusive_scan<MaxPlusOp>(acc, &tot);
  if (!APPLY) { if (threadIdx.x == 0) tile_agg[blockIdx.x] = tot; return; }
  MaxPlus run = MaxPlusOp::op(init, MaxPlusOp::op(tile_agg[blockIdx.x], exc));
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) {
      bool first; long long q;
      MaxPlus m = layout_map(k[i + 1], k[i], base + i > 0, c[i], key_lo_q, &first, &q);
      long long e = run.a > run.b ? run.a : run.b;      // f(0) of everything before this entry
      if (first && q > e) e = q;
      pos[base + i] = (uint64_t)e;
      run = MaxPlusOp::op(run, m);
    }
  }
  if (grand_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *grand_total = run;
}

__device__ __forceinline__ uint64_t entry_end(const uint64_t* keys, const uint32_t* counts, const uint64_t* pos, size_t k) {
  return pos[k] + enc_len((uint32_t)(keys[k] & 0xFF), counts[k]) - 1;
}

// first entry whose last slot is >= boundary (entries that reach into the next shard); one warp
__global__ void __launch_bounds__(32) tail_start_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                       const uint64_t* __restrict__ pos, size_t n, uint64_t boundary, uint64_t* __restrict__ out);

// ------------------------------------------------------- image tiles --------
constexpr int kTileBlocks = 64;                       // 64 blocks = 4096 slots per CTA
constexpr int kTileSlots = kTileBlocks * 64;
constexpr int kTileBytes = kTileBlocks * kBlockBytes;  // 5696, a multiple of 4
constexpr int kImgThreads = 128;

// Writes blocks [b0, b0+kTileBlocks) of the image (traveled words zero).
// q_base: quotient/slot index of block 0 of this image (shard offset); first_block: global index
// of block 0 (its offset byte is 0 only if it is global block 0 -- for shards > 0 the host fixes the
// carry separately).
__global__ void __launch_bounds__(kImgThreads) image_tiles_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                                 const uint64_t* __restrict__ pos, size_t n, uint64_t key_lo_q,
                                                                 uint64_t nblocks, uint8_t* __restrict__ image, uint64_t block_base = 0) {
  __shared__ __align__(16) uint8_t s_img[kTileBytes];
  __shared__ uint32_t s_occ[kTileBlocks * 2], s_run[kTileBlocks * 2];
  __shared__ size_t s_rng[4];
  // blocks [block_base, block_base + nblocks) of the filter are written to image[0 ...) (a shard's slice; block_base
  // is a multiple of kTileBlocks); slots and quotients are in the coordinates of (key >> 8) - key_lo_q
  const uint64_t b0 = block_base + (uint64_t)blockIdx.x * kTileBlocks;
  const uint64_t S0 = b0 * 64, S1 = S0 + kTileSlots;
  for (int i = threadIdx.x; i < kTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_img)[i] = 0;
  for (int i = threadIdx.x; i < kTileBlocks * 2; i += blockDim.x) { s_occ[i] = 0; s_run[i] = 0; }
  if (threadIdx.x < 4) {
    // 0: first entry whose last slot >= S0     1: first entry with pos >= S1
    // 2: first entry with quotient >= S0       3: first entry with quotient >= S1
    size_t lo = 0, hi = n;
    while (lo < hi) {
      size_t mid = lo + ((hi - lo) >> 1);
      bool before;
      if (threadIdx.x == 0) before = entry_end(keys, counts, pos, mid) < S0;
      else if (threadIdx.x == 1) before = pos[mid] < S1;
      else if (threadIdx.x == 2) before = ((keys[mid] >> 8) - key_lo_q) < S0;
      else before = ((keys[mid] >> 8) - key_lo_q) < S1;
      if (before) lo = mid + 1; else hi = mid;
    }
    s_rng[threadIdx.x] = lo;
  }
  __syncthreads();
  const size_t k_lo = s_rng[0], k_hi = s_rng[1], q_lo = s_rng[2], q_hi = s_rng[3];
  // slot bytes + runends
  for (size_t k = k_lo + threadIdx.x; k < k_hi; k += blockDim.x) {
    uint64_t key = keys[k];
    uint8_t enc[12];
    int len = enc_bytes((uint32_t)(key & 0xFF), counts[k], enc);
    uint64_t p = pos[k];
    for (int i = 0; i < len; i++) {
      uint64_t s = p + i;
      if (s >= S0 && s < S1) {
        uint32_t rel = (uint32_t)(s - S0);
        s_img[(rel >> 6) * kBlockBytes + 25 + (rel & 63)] = enc[i];
      }
    }
    bool last = (k + 1 == n) || ((keys[k + 1] >> 8) != (key >> 8));
    uint64_t e = p + len - 1;
    if (last && e >= S0 && e < S1) {
      uint32_t rel = (uint32_t)(e - S0);
      atomicOr(&s_run[rel >> 5], 1u << (rel & 31));
    }
  }
  // occupieds
  for (size_t k = q_lo + threadIdx.x; k < q_hi; k += blockDim.x) {
    uint64_t q = (keys[k] >> 8) - key_lo_q;
    bool first = (k == 0) || ((keys[k - 1] >> 8) != (keys[k] >> 8));
    if (first) { uint32_t rel = (uint32_t)(q - S0); atomicOr(&s_occ[rel >> 5], 1u << (rel & 31)); }
  }
  __syncthreads();
  // per block: offset byte + bitmap words
  for (int b = threadIdx.x; b < kTileBlocks; b += blockDim.x) {
    uint64_t gb = b0 + b;
    uint8_t* B = s_img + b * kBlockBytes;
    // offset = min(255, max(0, E - 64b + 1)), E = last slot of the last run with quotient < 64b
    uint32_t off = 0;
    if (gb > 0) {
      size_t lo = q_lo, hi = q_hi;   // last entry with quotient < 64*gb lies in [q_lo-1, q_hi)
      while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (((keys[mid] >> 8) - key_lo_q) < gb * 64) lo = mid + 1; else hi = mid;
      }
      if (lo > 0) {
        uint64_t E = entry_end(keys, counts, pos, lo - 1);
        if (E + 1 > gb * 64) { uint64_t o = E - gb * 64 + 1; off = o > 255 ? 255u : (uint32_t)o; }
      }
    }
    B[0] = (uint8_t)off;
    uint64_t occ = (uint64_t)s_occ[2 * b] | ((uint64_t)s_occ[2 * b + 1] << 32);
    uint64_t run = (uint64_t)s_run[2 * b] | ((uint64_t)s_run[2 * b + 1] << 32);
    for (int i = 0; i < 8; i++) { B[1 + i] = (uint8_t)(occ >> (8 * i)); B[9 + i] = (uint8_t)(run >> (8 * i)); }
  }
  __syncthreads();
  uint64_t nb_here = block_base + nblocks > b0 ? block_base + nblocks - b0 : 0;
  if (nb_here > kTileBlocks) nb_here = kTileBlocks;
  size_t bytes = (size_t)nb_here * kBlockBytes;
  uint8_t* dst = image + (b0 - block_base) * kBlockBytes;   // (b0 - block_base) % 4 == 0 -> 4-byte aligned
  size_t words = bytes / 4;
  for (size_t i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<uint32_t*>(s_img)[i];
  for (size_t i = words * 4 + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = s_img[i];
}

// ------------------------------------------------------ deNoise ranges ------
// The reference hands out sweep ranges of 2^20 slots stretched to a cluster end
// (cqf/CQF_mt.h:884-901) and its cleaners loop `while (start < end_bucket_id)`
// (cqf/gqf.c:2878-2886, CQF_mt.h:1024): a cluster that is one slot wide and sits exactly on a
// range end is never visited, so its count-1 key survives.  This walk finds those keys.
// One warp: the chain of ranges is inherently serial (nslots / 2^20 steps), but every step is a pair of
// searches and a walk to the end of a cluster, and those are done 32 entries at a time.
// warp_first_false: first index in [lo, hi) whose predicate is false (the predicate is true on a prefix).
template <class Pred>
__device__ __forceinline__ size_t warp_first_false(size_t lo, size_t hi, int lane, Pred&& pred) {
  while (lo < hi) {
    size_t span = hi - lo;
    size_t step = (span + 31) / 32;
    size_t idx = lo + (size_t)lane * step;
    bool t = idx < hi && pred(idx);
    uint32_t m = __ballot_sync(0xffffffffu, t);
    int c = __popc(m);                      // lanes 0..c-1 are true (monotone predicate)
    if (c == 0) return lo;
    size_t nlo = lo + (size_t)(c - 1) * step + 1;
    size_t nhi = lo + (size_t)c * step;
    if (nhi > hi) nhi = hi;
    lo = nlo; hi = nhi;
  }
  return lo;
}
// The same when the answer is expected close to `guess` (within `radius`): the chain walk knows where it is going -- the
// next range ends 2^20 slots ahead and the entries are spread evenly over the slots (ntHash), the next range starts
// right behind the last one's end -- so most searches need 1-4 dependent loads instead of log32(n) = 6.
template <class Pred>
__device__ __forceinline__ size_t warp_first_false_near(size_t lo, size_t hi, size_t guess, size_t radius, int lane, Pred&& pred) {
  if (hi - lo <= 2 * radius) return warp_first_false(lo, hi, lane, pred);
  size_t wlo = guess > lo + radius ? guess - radius : lo;
  if (wlo > hi - 2 * radius) wlo = hi - 2 * radius;
  const size_t whi = wlo + 2 * radius;
  // the answer lies in [wlo, whi] iff the predicate holds just below wlo and fails at whi
  bool ok = true;
  if (lane == 0) ok = wlo == lo || pred(wlo - 1);
  if (lane == 1) ok = whi == hi || !pred(whi);
  const uint32_t bad = __ballot_sync(0xffffffffu, !ok);
  if (bad == 0) return warp_first_false(wlo, whi, lane, pred);
  if (bad & 1u) return warp_first_false(lo, wlo, lane, pred);       // already false below the window
  return warp_first_false(whi, hi, lane, pred);
}
__global__ void __launch_bounds__(32) denoise_ranges_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                      const uint64_t* __restrict__ pos, size_t n, uint64_t nslots, uint64_t xnslots,
                                      uint32_t* __restrict__ keep_idx, uint32_t* __restrict__ n_keep, uint32_t keep_cap) {
  if (blockIdx.x) return;
  const int lane = threadIdx.x;
  const uint64_t min_len = 1ULL << 20;
  uint32_t nk = 0;
  if (n == 0) { if (lane == 0) *n_keep = 0; return; }
  uint64_t cur = keys[0] >> 8;  // find_first_nonempty_slot(qf, 0)
  size_t at = 0;                // first entry whose quotient is >= cur
  const double density = (double)n / (double)(pos[n - 1] - pos[0] + 1);
  while (cur < nslots) {
    uint64_t start = cur;
    uint64_t e0 = (start + min_len) > nslots ? nslots : (start + min_len);
    // lo = number of entries with pos <= e0: an entry takes at least one slot, so it is at most e0 - start + 1 entries past
    // the range's first entry (at) and, the entries being spread evenly, close to the interpolated index
    size_t lo;
    {
      const size_t hi = (size_t)std::min<uint64_t>((uint64_t)n, (uint64_t)at + (e0 - start) + 2);
      const size_t guess = at + (size_t)((double)(e0 - start) * density);
      lo = warp_first_false_near(at, hi, guess, 4096, lane, [&](size_t i) { return pos[i] <= e0; });
    }
    uint64_t end;
    size_t last = (size_t)-1;   // entry holding slot `end`, if any
    uint64_t prev_end = lo > 0 ? entry_end(keys, counts, pos, lo - 1) : 0;
    if (lo > 0 && prev_end >= e0) {
      // walk to the end of the cluster: first j >= lo-1 whose successor is not adjacent
      size_t j = lo - 1;
      for (;;) {
        size_t i = j + lane;
        bool adj = i + 1 < n && pos[i + 1] == entry_end(keys, counts, pos, i) + 1;
        uint32_t m = __ballot_sync(0xffffffffu, adj);
        int run = __ffs(~m) - 1;              // number of leading adjacent links (32 if all)
        if (run < 0) run = 32;
        j += run;
        if (run < 32) break;
      }
      end = entry_end(keys, counts, pos, j); last = j;   // find_first_empty_slot(e0) - 1
    } else {
      end = e0 - 1;
      if (lo > 0 && prev_end == end) last = lo - 1;
    }
    if (last != (size_t)-1 && end >= start) {
      // one-slot cluster [end,end]: entry starts and ends at `end`, previous entry not adjacent
      if (pos[last] == end && counts[last] == 1 && (last == 0 || entry_end(keys, counts, pos, last - 1) + 1 < end)) {
        if (nk < keep_cap && lane == 0) keep_idx[nk] = (uint32_t)last;
        nk++;
      }
    }
    // find_first_nonempty_slot(end + 1): next occupied quotient >= end+1 -- an entry or two past the one holding `end`
    const size_t from = last != (size_t)-1 ? last : (lo > 0 ? lo - 1 : 0);
    size_t nx = warp_first_false_near(from, n, from, 16, lane, [&](size_t i) { return (keys[i] >> 8) < end + 1; });
    cur = nx < n ? (keys[nx] >> 8) : xnslots;
    at = nx;
  }
  if (lane == 0) *n_keep = nk;
}

// ---- the same range walk over a SHARDED state -------------------------------------------------------------------
// The chain of sweep ranges runs over global slot positions, so with the state sharded by key range every rank walks its
// own entries and hands the chain over to the next rank.  The walk only ever moves forward through the entries, in three
// resumable phases:
//   A  seek the first entry whose quotient is >= need_q: it starts a range [start, e0 = min(start + 2^20, nslots)];
//   B  seek the first entry whose position is > e0; the entry before it decides: it covers slot e0 -> phase C,
//      otherwise the range ends at e0 - 1;
//   C  follow the cluster to its end (first entry not adjacent to its predecessor).
// At a range end the entry holding that slot is tested for the range-end rule (one-slot cluster, count 1).  What crosses
// a shard boundary is the phase, its operands and a summary of the last entry seen (position, end, count, gap before,
// owner), because the entry a decision is about may be the previous shard's last one (WalkState, 16 words).
struct WalkState {
  unsigned long long phase, need_q, e0, start, cur_end;
  unsigned long long prev_valid, prev_pos, prev_end, prev_count, prev_gap, prev_owner;
  unsigned long long protect_mask;     // bit r: rank r's LAST entry survives the sweep although its count is 1
  unsigned long long nslots, rank, reserved0, reserved1;
};
static_assert(sizeof(WalkState) == 128, "WalkState is 16 words");

struct WalkEntry { uint64_t pos, end; uint32_t count; bool gap; bool local; size_t idx; unsigned long long owner; };

__device__ __forceinline__ bool walk_rule(const WalkEntry& L, uint64_t end) {
  // cqf/gqf.c:2878-2886 + cqf/CQF_mt.h:1024: a one-slot cluster [end, end] is never visited by `while (start < end)`
  return L.pos == end && L.end == end && L.count == 1 && L.gap;
}

__global__ void __launch_bounds__(32) denoise_walk_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                         const uint64_t* __restrict__ pos, size_t n, WalkState* __restrict__ state,
                                                         uint32_t* __restrict__ keep_idx, uint32_t* __restrict__ n_keep, uint32_t keep_cap) {
  const int lane = threadIdx.x;
  WalkState S = *state;
  uint32_t nk = 0;
  const uint64_t min_len = 1ULL << 20;
  // summary of local entry i (its predecessor is local entry i - 1, or the entry handed over by the previous shards)
  auto entry = [&](size_t i) {
    WalkEntry e;
    e.pos = pos[i]; e.end = entry_end(keys, counts, pos, i); e.count = counts[i]; e.local = true; e.idx = i; e.owner = S.rank;
    if (i > 0) e.gap = entry_end(keys, counts, pos, i - 1) + 1 < e.pos;
    else e.gap = !S.prev_valid || S.prev_end + 1 < e.pos;
    return e;
  };
  auto handed = [&]() {
    WalkEntry e;
    e.pos = S.prev_pos; e.end = S.prev_end; e.count = (uint32_t)S.prev_count; e.gap = S.prev_gap != 0; e.local = false; e.idx = 0; e.owner = S.prev_owner;
    return e;
  };
  auto protect = [&](const WalkEntry& L) {
    if (L.local) { if (nk < keep_cap && lane == 0) keep_idx[nk] = (uint32_t)L.idx; nk++; }
    else S.protect_mask |= 1ULL << L.owner;
  };
  size_t i = 0;
  const double density = n > 1 ? (double)n / (double)(pos[n - 1] - pos[0] + 1) : 1.0;
  while (i < n || false) {
    if (S.phase == 0) {
      const unsigned long long need = S.need_q;
      const size_t j = warp_first_false_near(i, n, i, 16, lane, [&](size_t x) { return (keys[x] >> 8) < need; });
      if (j == n) { i = n; break; }
      S.start = keys[j] >> 8;
      S.e0 = S.start + min_len > S.nslots ? S.nslots : S.start + min_len;
      S.phase = 1; i = j;
    } else if (S.phase == 1) {
      const unsigned long long e0 = S.e0;
      size_t lo;
      {   // an entry takes at least one slot; the entries are spread evenly over the slots
        const uint64_t pi = pos[i];
        const size_t hi = pi <= e0 ? (size_t)std::min<uint64_t>((uint64_t)n, (uint64_t)i + (e0 - pi) + 2) : i + 1;
        const size_t guess = pi <= e0 ? i + (size_t)((double)(e0 - pi) * density) : i;
        lo = warp_first_false_near(i, hi > n ? n : hi, guess, 4096, lane, [&](size_t x) { return pos[x] <= e0; });
      }
      if (lo == n) { i = n; break; }
      WalkEntry P;
      bool have = true;
      if (lo > 0) P = entry(lo - 1); else if (S.prev_valid) P = handed(); else have = false;
      if (have && P.end >= e0) { S.phase = 2; S.cur_end = P.end; i = lo; }
      else {
        const uint64_t end = e0 - 1;
        if (have && P.end == end && walk_rule(P, end)) protect(P);
        S.need_q = end + 1; S.phase = 0; i = lo;
      }
    } else {
      // first entry from i on that is not adjacent to its predecessor (entry i's predecessor ends at cur_end)
      size_t j = i;
      for (;;) {
        const size_t x = j + lane;
        bool adj = false;
        if (x < n) {
          const uint64_t pe = x == i ? S.cur_end : entry_end(keys, counts, pos, x - 1);
          adj = pos[x] == pe + 1;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, adj);
        int run = __ffs(~m) - 1;
        if (run < 0) run = 32;
        j += run;
        if (run < 32 || j >= n) break;
      }
      if (j >= n) { S.cur_end = entry_end(keys, counts, pos, n - 1); i = n; break; }     // the cluster may go on in the next shard
      WalkEntry Lst;
      if (j > 0) Lst = entry(j - 1); else Lst = handed();       // j == 0: the cluster ended with the handed-over entry
      const uint64_t end = j > i ? Lst.end : S.cur_end;
      if (walk_rule(Lst, end)) protect(Lst);
      S.need_q = end + 1; S.phase = 0; i = j;
    }
  }
  // every local entry has been passed: the last one becomes the handed-over entry
  if (n > 0) {
    const WalkEntry e = entry(n - 1);
    S.prev_valid = 1; S.prev_pos = e.pos; S.prev_end = e.end; S.prev_count = e.count; S.prev_gap = e.gap ? 1 : 0; S.prev_owner = S.rank;
    if (S.phase == 2 && S.cur_end < e.end) S.cur_end = e.end;
  }
  if (lane == 0) { *state = S; *n_keep = nk; }
}

__global__ void __launch_bounds__(32) tail_start_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                       const uint64_t* __restrict__ pos, size_t n, uint64_t boundary, uint64_t* __restrict__ out) {
  size_t k = warp_first_false(0, n, threadIdx.x, [&](size_t i) { return entry_end(keys, counts, pos, i) < boundary; });
  if (threadIdx.x == 0) *out = (uint64_t)k;
}

// ------------------------------------------------------------ queries -------
struct ImageView {
  uint8_t* blocks; uint64_t nblocks; uint64_t xnslots; uint64_t q_base;  // q_base: quotient of slot 0 (shards)
  // sharded image: the blocks of the filter live in `nshards` arrays (blocks_per_shard blocks each, the last one also holds
  // the xnslots extension), possibly in other GPUs' memory (peer mappings); null = one array at `blocks`
  uint8_t* const* shard_blocks = nullptr; uint64_t blocks_per_shard = 0; uint32_t nshards = 0;
};
__device__ __forceinline__ uint64_t ld64u(const uint8_t* p) {  // unaligned little-endian load
  uint64_t v = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
  return v;
}
__device__ __forceinline__ uint8_t* blk_rw(const ImageView& im, uint64_t b) {
  if (!im.shard_blocks) return im.blocks + b * kBlockBytes;
  uint64_t sh = b / im.blocks_per_shard;
  if (sh >= im.nshards) sh = im.nshards - 1;
  return im.shard_blocks[sh] + (b - sh * im.blocks_per_shard) * kBlockBytes;
}
__device__ __forceinline__ const uint8_t* blk(const ImageView& im, uint64_t b) { return blk_rw(im, b); }
__device__ __forceinline__ uint8_t img_slot(const ImageView& im, uint64_t s) { return blk(im, s >> 6)[25 + (s & 63)]; }
__device__ __forceinline__ bool img_bit(const ImageView& im, uint64_t s, int field) {
  return (blk(im, s >> 6)[1 + 8 * field + ((s & 63) >> 3)] >> (s & 7)) & 1;
}
// position of the r-th (0-based) set bit of v, 64 if none
__device__ __forceinline__ int select64(uint64_t v, int r) {
  uint32_t lo = (uint32_t)v; int pl = __popc(lo);
  if (r < pl) return __fns(lo, 0, r + 1);
  uint32_t hi = (uint32_t)(v >> 32); r -= pl;
  if (r < __popc(hi)) return 32 + __fns(hi, 0, r + 1);
  return 64;
}
// run_end (cqf/gqf.c:655-704) on the packed image; a saturated offset byte is resolved like
// block_offset() does, by recursing on the previous quotient.
__device__ uint64_t img_run_end(const ImageView& im, uint64_t q);
__device__ __forceinline__ uint64_t img_block_offset(const ImageView& im, uint64_t b) {
  uint8_t o = blk(im, b)[0];
  if (o < 255) return o;
  return img_run_end(im, 64 * b - 1) - 64 * b + 1;
}
__device__ uint64_t img_run_end(const ImageView& im, uint64_t q) {
  // iterative form of the recursion run_end -> block_offset -> run_end(64b-1): walk back to a block
  // whose offset byte is not saturated, then resolve forward.
  uint64_t stack_q[40]; int sp = 0;
  uint64_t cur = q, known_off = 0; bool have_off = false;
  for (;;) {
    uint64_t b = cur >> 6;
    uint8_t o = blk(im, b)[0];
    if (o < 255 || b == 0) { known_off = o; have_off = true; }
    if (have_off) {
      // resolve run_end(cur) with block offset known_off
      for (;;) {
        uint64_t bb = cur >> 6; uint32_t io = (uint32_t)(cur & 63);
        uint64_t occ = ld64u(blk(im, bb) + 1);
        int rank = __popcll(occ & ((io == 63) ? ~0ULL : ((2ULL << io) - 1)));
        uint64_t res;
        if (rank == 0) {
          res = (known_off <= io) ? cur : (64 * bb + known_off - 1);
        } else {
          uint64_t rb = bb + known_off / 64; uint32_t ign = (uint32_t)(known_off % 64); int rr = rank - 1;
          for (;;) {
            uint64_t runs = ld64u(blk(im, rb) + 9);
            uint64_t masked = ign ? (runs & ~((1ULL << ign) - 1)) : runs;
            int s = select64(masked, rr);
            if (s < 64) { res = 64 * rb + s; break; }
            rr -= __popcll(masked); rb++; ign = 0;
            if (rb >= im.nblocks) { res = im.xnslots - 1; break; }
          }
          if (res < cur) res = cur;
        }
        if (sp == 0) return res;
        // res = run_end(64b-1) for the pending query on top of the stack
        uint64_t pq = stack_q[--sp];
        known_off = res - 64 * (pq >> 6) + 1;
        cur = pq;
      }
    }
    if (sp >= 40) return cur;  // cannot happen: offsets saturate only inside clusters > 255 slots
    stack_q[sp++] = cur;
    cur = 64 * b - 1;
  }
}

// decode_counter (cqf/gqf.c:1259-1299): entry starting at slot `idx` inside a run ending at run_end.
__device__ __forceinline__ uint64_t img_decode(const ImageView& im, uint64_t idx, uint64_t run_end, uint32_t* rem, uint64_t* count) {
  uint32_t r = img_slot(im, idx);
  *rem = r;
  if (idx == run_end) { *count = 1; return idx; }
  uint32_t d = img_slot(im, idx + 1);
  if (d > r) { *count = 1; return idx; }
  uint64_t e = idx + 1, c = 0;
  if (d == 0) { e++; d = img_slot(im, e); }
  while (d & 0x80) { c = c * 128 + (d & 0x7F); e++; d = img_slot(im, e); }
  c = c * 128 + d;
  *count = c + 1;
  return e;
}

// qf_count_key_value[_set_traveled] (cqf/gqf.c:2442-2469, 3092-3128), one thread per key.
template <bool SET>
__global__ void query_kernel(ImageView im, const uint64_t* __restrict__ qkeys, size_t n, uint64_t* __restrict__ counts,
                             uint8_t* __restrict__ was) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t key = qkeys[i];
  uint32_t rem = (uint32_t)(key & 0xFF);
  uint64_t q = (key >> 8) - im.q_base;
  uint64_t cnt = 0; uint8_t w = 0;
  if (q < im.nblocks * 64 && img_bit(im, q, 0)) {
    uint64_t rs = q == 0 ? 0 : img_run_end(im, q - 1) + 1;
    if (rs < q) rs = q;
    // the run of q ends at the first runend bit at or after rs
    uint64_t re = rs;
    while (!img_bit(im, re, 1)) re++;
    uint64_t idx = rs;
    while (idx <= re) {
      uint32_t r; uint64_t c;
      uint64_t e = img_decode(im, idx, re, &r, &c);
      if (r == rem) {
        cnt = c;
        if (SET) {
          // traveled bit of the entry's first slot; 32-bit atomic on the word holding that byte
          uintptr_t addr = reinterpret_cast<uintptr_t>(blk_rw(im, idx >> 6)) + 17 + ((idx & 63) >> 3);
          unsigned int* wp = reinterpret_cast<unsigned int*>(addr & ~(uintptr_t)3);
          unsigned int bit = 1u << (8 * (addr & 3) + (idx & 7));
          unsigned int old = atomicOr(wp, bit);
          w = (old & bit) ? 1 : 0;
        }
        break;
      }
      if (r > rem) break;
      idx = e + 1;
    }
  }
  counts[i] = cnt;
  if (SET) was[i] = w;
}

// ------------------------------------------------------------- decode -------
// Image -> sorted (key,count) list, one thread per block-aligned group of quotients.
// Pass 1 (WRITE=false) counts entries per block of quotients; pass 2 writes them at out_off[b].
template <bool WRITE>
__global__ void decode_blocks_kernel(ImageView im, uint32_t* __restrict__ n_entries, const uint64_t* __restrict__ out_off,
                                     uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts, int* __restrict__ err) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= im.nblocks) return;
  uint64_t occ = ld64u(blk(im, b) + 1);
  uint32_t cnt = 0;
  uint64_t o = WRITE ? out_off[b] : 0;
  if (occ) {
    uint64_t q0 = 64 * b + (__ffsll((long long)occ) - 1);
    uint64_t rs = q0 == 0 ? 0 : img_run_end(im, q0 - 1) + 1;
    if (rs < q0) rs = q0;
    while (occ) {
      int bitp = __ffsll((long long)occ) - 1; occ &= occ - 1;
      uint64_t q = 64 * b + bitp;
      if (rs < q) rs = q;
      uint64_t re = rs;
      while (re < im.xnslots && !img_bit(im, re, 1)) re++;
      uint64_t idx = rs;
      while (idx <= re) {
        uint32_t r; uint64_t c;
        uint64_t e = img_decode(im, idx, re, &r, &c);
        if (WRITE) {
          out_keys[o + cnt] = ((q + im.q_base) << 8) | r;
          if (c > 0xFFFFFFFFull) { *err = SHQ_ERR_COUNT_OVERFLOW; c = 0xFFFFFFFFull; }
          out_counts[o + cnt] = (uint32_t)c;
        }
        cnt++;
        idx = e + 1;
      }
      rs = re + 1;
    }
  }
  if (!WRITE) n_entries[b] = cnt;
}

}  // namespace shq
