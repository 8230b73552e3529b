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
// One CTA per bin (or per key sub-range of a bin) streams the bin's 32-bit words through an ORDER-PRESERVING
// open-addressing table in shared memory: the home slot of a key is a monotone function of the key (keys are ntHash
// values, i.e. uniform), probing is linear and never wraps.  A key with a smaller home is a smaller key, so the sorted
// position of a key is (number of keys with a smaller home) + (its rank among the keys that share its home).  The
// first term is an exclusive scan over per-home key counters (one byte per home, bumped by whoever inserts a new
// key); the second is 0 for most keys (homes are rarely shared at the load factors used) and otherwise costs a short
// walk from the home slot.  No insertion-order list, no separate sort: counting the duplicates and sorting the distinct
// keys are the same pass.  Two CTAs of 512 threads share an SM.
constexpr int kAggThreads = 512;
constexpr int kAggWarps = kAggThreads / 32;
constexpr int kAggHomes = 8192;          // home positions (a power of two)
constexpr int kAggPad = 256;             // slots behind the last home: probing never wraps around
constexpr int kAggTable = kAggHomes + kAggPad;
constexpr int kAggMaxProbe = kAggPad - 2;   // a longer probe sequence means skewed keys: restart with narrower key ranges
constexpr int kAggMaxDistinct = 5000;    // what the host sizes the bins for (load factor 0.6)
constexpr int kAggBuckets = 2048;        // upper bound of chunks per batch (first-chunk histograms)
constexpr int kAggU = 8;                 // words per thread and round
constexpr int kAggWarpTile = 32 * kAggU; // words of one warp per round (= capacity of its pending list)
constexpr int kAggSPT = (kAggTable + kAggThreads - 1) / kAggThreads;   // slots per thread in the emission
constexpr int kAggHPT = kAggHomes / kAggThreads;                       // homes per thread in the scan (16 = 4 counter words)
constexpr uint32_t kAggEmpty = 0xFFFFFFFFu;
// t_key, t_cnt (u32), the per-home key counters (u8) [, t_chk (u16)], then one region shared by the per-warp pending
// lists (insert phase) and the per-home output offsets (sort phase) [, the slots in key order (SPLIT only)]
constexpr size_t kAggOrdBytes = ((size_t)kAggTable * 2 + 15) & ~(size_t)15;
constexpr size_t kAggPendBytes = (size_t)kAggWarps * kAggWarpTile * 4;
constexpr size_t kAggShareBytes = kAggPendBytes > kAggOrdBytes ? kAggPendBytes : kAggOrdBytes;
constexpr size_t agg_smem_bytes(bool chunk, bool split) {
  return (size_t)kAggTable * 8 + (size_t)kAggHomes + (chunk ? kAggOrdBytes : 0) + kAggShareBytes + (split ? kAggOrdBytes : 0);
}

struct AggOut {
  uint64_t* keys; uint32_t* counts; uint16_t* chunk; uint32_t* n_out;   // run A (or the only run)
  uint64_t* keys2; uint32_t* counts2; uint32_t* n_out2;                 // run B (SPLIT only): occurrences with chunk > cstar
  uint32_t* occ_out; uint32_t* occ_out2;
  uint32_t* first_hist;     // optional: histogram over first-chunk ids of the distinct keys (state empty => all new)
  int* split_overflow;      // SPLIT: set when a 16-bit half counter would overflow
  const uint32_t* bin_end;  // optional: end of every bin (fixed-capacity bins); null = bin_off[p + 1]
  int fixed_bits;           // each bin is handled by 2^fixed_bits CTAs, one per key sub-range
  size_t sub_stride;        // outputs of sub-range s live at [s * sub_stride + bin_off[p] + i]
  int direct;               // most words are first occurrences: insert straight away instead of probing first
  int sample_bits;          // > 0: aggregate only the keys of the first 1/2^sample_bits of the bin's key range (a sample by KEY,
                            // so its distinct ratio is that of the whole bin) -- one pass over the bin by one CTA
  // dense != 0: the runs are written DENSE and in order straight away (no per-bin staging + compaction pass): CTAs take
  // tickets, publish their distinct counts and find their output offset by a decoupled look-back over the
  // predecessors (status words, zeroed; counts of run A in the low 31 bits, of run B above).  keys/counts/chunk are then
  // the dense run arrays; totals[0..3] receive n(A), n(B), occurrences(A), occurrences(B).  A CTA that needs several
  // rounds (key sub-ranges) stages its rounds in scratch_* first.
  int dense; uint32_t n_groups; unsigned long long* status; unsigned int* ticket; unsigned long long* totals;
  uint64_t* scratch_keys; uint32_t* scratch_counts; uint16_t* scratch_chunk; uint64_t* scratch_keys2; uint32_t* scratch_counts2;
};

struct AggTable {
  uint32_t* key; uint32_t* cnt; uint16_t* chk;
  uint32_t* hcnt;                     // keys per home, one byte each (four homes per word)
  volatile uint32_t* overflow; int* split_overflow;
  uint32_t relmask; int hshift;       // home of a key: (key & relmask) >> hshift (or << -hshift): monotone, all kAggHomes slots
};
__device__ __forceinline__ uint32_t agg_home(const AggTable& T, uint32_t kl) {
  const uint32_t rel = kl & T.relmask;
  return T.hshift >= 0 ? (rel >> T.hshift) : (rel << (-T.hshift));
}

// 16-bit minimum on a shared-memory half word (rarely taken: bins receive their words in chunk order)
__device__ __forceinline__ void agg_chunk_min(uint16_t* chk, uint32_t h, uint32_t ch) {
  if (ch >= chk[h]) return;
  uint32_t* wp = reinterpret_cast<uint32_t*>(chk) + (h >> 1);
  const int sh = (h & 1) * 16;
  uint32_t old = *wp;
  for (;;) {
    if (((old >> sh) & 0xFFFFu) <= ch) return;
    uint32_t nv = (old & ~(0xFFFFu << sh)) | (ch << sh);
    uint32_t seen = atomicCAS(wp, old, nv);
    if (seen == old) return;
    old = seen;
  }
}

// predicated shared-memory operations: pass 1 of the insert stays free of branches, so a warp pays for it once
// and not once per (hit, miss) side
__device__ __forceinline__ void smem_red_add_if(bool p, uint32_t* addr, uint32_t v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q red.shared.add.u32 [%1], %2; }"
               :: "r"((uint32_t)p), "r"((uint32_t)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
}
__device__ __forceinline__ void smem_st_if(bool p, uint32_t* addr, uint32_t v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.u32 [%1], %2; }"
               :: "r"((uint32_t)p), "r"((uint32_t)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
}

template <bool SPLIT, bool CHUNK>
__device__ __forceinline__ void agg_count(const AggTable& T, uint32_t h, uint32_t ch, bool second) {
  if (SPLIT) {
    uint32_t old = atomicAdd(&T.cnt[h], second ? 0x10000u : 1u);
    if ((second ? (old >> 16) : (old & 0xFFFFu)) >= 0xFFFEu) *T.split_overflow = 1;
  } else {
    atomicAdd(&T.cnt[h], 1u);
  }
  if (CHUNK && !second) agg_chunk_min(T.chk, h, ch);
}

// find-or-insert by linear probing from the key's home; false (and the overflow flag) if the probe sequence gets too long
__device__ __forceinline__ bool agg_place(const AggTable& T, uint32_t kl, uint32_t* slot) {
  const uint32_t h0 = agg_home(T, kl);
  uint32_t h = h0;
  for (int pr = 0; pr < kAggMaxProbe; pr++, h++) {
    const uint32_t cur = T.key[h];
    if (cur == kl) { *slot = h; return true; }
    if (cur == kAggEmpty) {
      const uint32_t old = atomicCAS(&T.key[h], kAggEmpty, kl);
      if (old == kAggEmpty) { atomicAdd(&T.hcnt[h0 >> 2], 1u << ((h0 & 3u) * 8)); *slot = h; return true; }   // a new key of home h0
      if (old == kl) { *slot = h; return true; }
    }
  }
  *T.overflow = 1;
  return false;
}

// Streams the bin through the table.  Every warp works on its own tiles of 256 words.
//  DIRECT (most words are first occurrences, e.g. a batch of a large genome): find-or-insert per word.
//  otherwise (30-100x coverage inside the batch):
//   pass 1 (all lanes, no loop): one probe of the home slot and its neighbour; a word whose key sits there is counted on the spot,
//           every other word is appended to the warp's pending list;
//   pass 2 (dense over the pending list): find-or-insert.
// Other warps may insert while a warp probes: a probe then sees either "empty" (-> pending -> found in pass 2)
// or the key itself, both correct.  Compile-time flags strip everything a plain batch does not need:
// FILT = chunk-range / key-sub-range filters, CHUNK = first-chunk minimum, SPLIT = two counters.
template <bool SPLIT, bool CHUNK, bool FILT, bool DIRECT>
__device__ __forceinline__ void agg_insert_pass(const uint32_t* __restrict__ packed, uint32_t beg, uint32_t end, int lb, uint32_t lowmask,
                                                uint32_t chunk_lo, uint32_t chunk_hi, uint32_t cstar, int sub_bits, uint32_t r,
                                                const AggTable& T, uint32_t* s_pend, uint32_t* occ1, uint32_t* occ2) {
  uint32_t my_occ = 0, my_occ2 = 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* pend = s_pend + wid * kAggWarpTile;
  // rounds of kAggThreads x kAggU words; inside a round word u of thread t is base + u * kAggThreads + t, so a short last
  // round is shared by all warps (a warp-contiguous split would leave most warps idle through it)
  for (uint32_t t0 = beg; t0 < end; t0 += kAggThreads * kAggU) {
    uint32_t wv[kAggU];
#pragma unroll
    for (int u = 0; u < kAggU; u++) { uint32_t i = t0 + u * kAggThreads + threadIdx.x; wv[u] = i < end ? __ldg(&packed[i]) : 0; }
    uint32_t np = 0;
#pragma unroll
    for (int u = 0; u < kAggU; u++) {
      const uint32_t w = wv[u];
      bool live = t0 + u * kAggThreads + threadIdx.x < end;
      const uint32_t kl = w & lowmask;
      uint32_t ch = 0;
      if (CHUNK || FILT || SPLIT) ch = lb >= 32 ? 0 : (w >> lb);
      if (FILT) {
        live = live & (ch >= chunk_lo) & (ch <= chunk_hi);
        if (sub_bits) live = live & ((kl >> (lb - sub_bits)) == r);
      }
      const bool second = SPLIT && ch > cstar;
      if (FILT || SPLIT) { my_occ2 += (live & second) ? 1u : 0u; my_occ += (live & !second) ? 1u : 0u; }
      if (DIRECT) {
        uint32_t h;
        if (live && agg_place(T, kl, &h)) agg_count<SPLIT, CHUNK>(T, h, ch, second);
        continue;
      }
      const uint32_t h0 = agg_home(T, kl);
      const uint32_t p0 = T.key[h0], p1 = T.key[h0 + 1];          // the home slot and its neighbour
      const bool hit = live & ((p0 == kl) | (p1 == kl));
      const uint32_t h = h0 + (p1 == kl ? 1u : 0u);
      if (SPLIT) {
        if (hit) agg_count<SPLIT, CHUNK>(T, h, ch, second);
      } else {
        smem_red_add_if(hit, &T.cnt[h], 1u);
        if (CHUNK) { if (hit & (ch < T.chk[h])) agg_chunk_min(T.chk, h, ch); }
      }
      const bool miss = live & !hit;
      const uint32_t m = __ballot_sync(0xffffffffu, miss);
      smem_st_if(miss, &pend[np + __popc(m & ((1u << lane) - 1))], w);
      np += __popc(m);
    }
    if (DIRECT) continue;
    __syncwarp();
    np = __shfl_sync(0xffffffffu, np, 0);
    // np is the same in every lane: a warp-uniform loop with one structured branch inside, so the warp
    // reconverges after every group of 32 pending words
    for (uint32_t i0 = 0; i0 < np; i0 += 32) {
      const uint32_t i = i0 + lane;
      if (i < np) {
        const uint32_t w = pend[i];
        const uint32_t kl = w & lowmask;
        uint32_t ch = 0;
        if (CHUNK || SPLIT) ch = lb >= 32 ? 0 : (w >> lb);
        const bool second = SPLIT && ch > cstar;
        uint32_t h;
        if (agg_place(T, kl, &h)) agg_count<SPLIT, CHUNK>(T, h, ch, second);
      }
      __syncwarp();
    }
  }
  if (FILT || SPLIT) { atomicAdd(occ1, my_occ); if (SPLIT) atomicAdd(occ2, my_occ2); }
  else if (threadIdx.x == 0) atomicAdd(occ1, end - beg);
}

// Decoupled look-back over (flag | value) words, one per ticket: flag 1 = the ticket's own value, 2 = inclusive prefix.
// Called by the 32 lanes of one warp; returns the exclusive prefix of ticket t (values may pack two counters).
__device__ __forceinline__ unsigned long long agg_lookback(unsigned long long* status, size_t t, unsigned long long val, int lane) {
  const unsigned long long kVal = (1ULL << 62) - 1;
  unsigned long long excl = 0;
  if (t == 0) {
    if (lane == 0) atomicExch(&status[0], (2ULL << 62) | val);
    return 0;
  }
  if (lane == 0) atomicExch(&status[t], (1ULL << 62) | val);
  long long p = (long long)t - 1;   // lane l looks at ticket p - l
  for (;;) {
    const long long q = p - lane;
    unsigned long long v = 2ULL << 62;              // tickets before 0: an inclusive prefix of 0
    if (q >= 0) { do { v = *reinterpret_cast<volatile unsigned long long*>(&status[q]); } while ((v >> 62) == 0); }
    const unsigned incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
    const int first = incl ? __ffs(incl) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
    unsigned long long part = lane <= first ? (v & kVal) : 0;
    for (int d = 16; d; d >>= 1) part += __shfl_down_sync(0xffffffffu, part, d);
    excl += __shfl_sync(0xffffffffu, part, 0);
    if (first < 32) break;
    p -= 32;
  }
  if (lane == 0) atomicExch(&status[t], (2ULL << 62) | (excl + val));
  return excl;
}

// One CTA per (bin, key sub-range).  packed[bin_off[p] .. bin_off[p+1]) holds (key_low | chunk << lb).
// Occurrences whose chunk is outside [chunk_lo, chunk_hi] are ignored.
// SPLIT: occurrences with chunk <= cstar are counted in the low half of the counter, the others in the high
// half, and two sorted runs come out of one pass (deNoise boundary inside a batch).
// Output (sorted by key) goes to out[sub * sub_stride + bin_off[p] + i].  A sub-range whose probe sequences get too long
// (more keys than the table is sized for, or skewed raw keys) is re-streamed in 2^k narrower key ranges.
template <bool SPLIT, bool CHUNK>
__global__ void __launch_bounds__(kAggThreads, 2) bin_aggregate_kernel(
    const uint32_t* __restrict__ packed, const uint32_t* __restrict__ bin_off, int lb, uint64_t key_lo,
    uint32_t chunk_lo, uint32_t chunk_hi, uint32_t cstar, AggOut o) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* t_key = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* t_cnt = t_key + kAggTable;
  uint32_t* t_hcnt = t_cnt + kAggTable;                                   // keys per home: kAggHomes bytes
  unsigned char* nxt = reinterpret_cast<unsigned char*>(t_hcnt + kAggHomes / 4);
  uint16_t* t_chk = reinterpret_cast<uint16_t*>(nxt);
  if (CHUNK) nxt += kAggOrdBytes;
  uint32_t* s_pend = reinterpret_cast<uint32_t*>(nxt);                    // insert phase: per-warp pending lists
  uint16_t* s_hpre = reinterpret_cast<uint16_t*>(nxt);                    // sort phase: keys with a smaller home, per home
  uint16_t* s_fin = reinterpret_cast<uint16_t*>(nxt + kAggShareBytes);    // SPLIT: slots in key order
  __shared__ uint32_t s_nocc, s_nocc2, s_overflow, s_na, s_nb;

  __shared__ uint32_t s_ticket;
  __shared__ unsigned long long s_excl;
  if (o.dense) { if (threadIdx.x == 0) s_ticket = atomicAdd(o.ticket, 1u); __syncthreads(); }
  const uint32_t g = o.dense ? s_ticket : blockIdx.x;      // (bin, sub-range) group handled by this CTA; tickets keep the look-back deadlock free
  const uint32_t p = g >> o.fixed_bits;
  const uint32_t fsub = g & ((1u << o.fixed_bits) - 1);
  const uint32_t beg = bin_off[p];
  uint32_t end = bin_off[p + 1];
  if (o.bin_end) { const uint32_t e = o.bin_end[p]; end = e < end ? e : end; }     // never past the bin's own range
  const size_t obase = (size_t)fsub * o.sub_stride + beg;       // staging position of this group (per-bin layout)
  const uint32_t lowmask = lb >= 32 ? 0xFFFFFFFFu : ((1u << lb) - 1);
  uint32_t written = 0, written2 = 0, occ_total = 0, occ_total2 = 0;
  bool did_direct = false;      // dense output: this CTA has already looked back and written its run in place
  int sub_bits = o.fixed_bits;   // key_low bits fixed per round: fixed_bits chosen by the grid, the rest by overflow restarts
  while (sub_bits < lb && ((end - beg) >> sub_bits) > 64u * kAggMaxDistinct) sub_bits++;
  if (o.sample_bits > sub_bits) sub_bits = o.sample_bits < lb ? o.sample_bits : lb - 1;

restart:
  written = 0; written2 = 0; occ_total = 0; occ_total2 = 0;
  for (uint32_t r = fsub << (sub_bits - o.fixed_bits); r < (o.sample_bits ? (fsub << (sub_bits - o.fixed_bits)) + 1 : ((fsub + 1) << (sub_bits - o.fixed_bits))); r++) {
    {
      uint4* k4 = reinterpret_cast<uint4*>(t_key);
      uint4* c4 = reinterpret_cast<uint4*>(t_cnt);
      const uint4 e4 = make_uint4(kAggEmpty, kAggEmpty, kAggEmpty, kAggEmpty), z4 = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < kAggTable / 4; i += blockDim.x) { k4[i] = e4; c4[i] = z4; }
      t_hcnt[threadIdx.x] = 0; t_hcnt[threadIdx.x + kAggThreads] = 0; t_hcnt[threadIdx.x + 2 * kAggThreads] = 0; t_hcnt[threadIdx.x + 3 * kAggThreads] = 0;
      if (CHUNK) { uint4* h4 = reinterpret_cast<uint4*>(t_chk); for (int i = threadIdx.x; i < kAggTable / 8; i += blockDim.x) h4[i] = e4; }
    }
    if (threadIdx.x == 0) { s_nocc = 0; s_nocc2 = 0; s_overflow = 0; s_na = 0; s_nb = 0; }
    __syncthreads();
    const int bits = lb - sub_bits;                    // key_low bits that vary inside this round
    const uint32_t relmask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1);
    {
      const uint32_t max_chunk = lb >= 32 ? 0u : (0xFFFFFFFFu >> lb);
      const bool filt = sub_bits != 0 || chunk_lo != 0 || chunk_hi < max_chunk;   // sub_bits includes the grid's fixed bits
      AggTable T{t_key, t_cnt, t_chk, t_hcnt, &s_overflow, o.split_overflow, relmask, bits - 13};
#define SHQ_AGG_CALL(F, D) agg_insert_pass<SPLIT, CHUNK, F, D>(packed, beg, end, lb, lowmask, chunk_lo, chunk_hi, cstar, sub_bits, r, T, s_pend, &s_nocc, &s_nocc2)
      if (o.direct) { if (filt) SHQ_AGG_CALL(true, true); else SHQ_AGG_CALL(false, true); }
      else { if (filt) SHQ_AGG_CALL(true, false); else SHQ_AGG_CALL(false, false); }
#undef SHQ_AGG_CALL
    }
    __syncthreads();
    if (*reinterpret_cast<volatile uint32_t*>(&s_overflow)) { sub_bits++; __syncthreads(); goto restart; }
    // ---- output offsets per home: every thread owns kAggHPT consecutive homes (four counter words); the exclusive scan of
    // the per-home key counts is the number of keys with a smaller home, i.e. where the keys of a home start in the run
    uint32_t nd;
    {
      const uint4 cw = reinterpret_cast<const uint4*>(t_hcnt)[threadIdx.x];
      const uint32_t w4[4] = {cw.x, cw.y, cw.z, cw.w};
      const uint32_t mine = __vsadu4(cw.x, 0u) + __vsadu4(cw.y, 0u) + __vsadu4(cw.z, 0u) + __vsadu4(cw.w, 0u);
      uint32_t tot;
      uint32_t ex = block_exclusive_scan<SumU32>(mine, &tot);     // ends with a barrier: the pending lists are dead
      nd = tot;
      uint32_t pk[kAggHPT / 2];
#pragma unroll
      for (int j = 0; j < kAggHPT; j++) {
        if (j & 1) pk[j >> 1] |= ex << 16; else pk[j >> 1] = ex;
        ex += (w4[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
      }
      uint4* dst = reinterpret_cast<uint4*>(s_hpre + threadIdx.x * kAggHPT);
      dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    // dense output: a CTA that is done after this round (the usual case) learns its offset now and writes the run in place;
    // one that needs more rounds stages them in the scratch arrays and copies them over at the end
    const bool last_round = r + 1 == (o.sample_bits ? (fsub << (sub_bits - o.fixed_bits)) + 1 : ((fsub + 1) << (sub_bits - o.fixed_bits)));
    const bool direct_out = o.dense && !SPLIT && last_round && written == 0;
    if (direct_out) {
      if (threadIdx.x < 32) {
        const unsigned long long excl = agg_lookback(o.status, g, (unsigned long long)nd, threadIdx.x);
        if (threadIdx.x == 0) { s_excl = excl; if (g == o.n_groups - 1) { o.totals[0] = excl + nd; o.totals[1] = 0; } }
      }
      did_direct = true;
    }
    __syncthreads();
    uint64_t* const wkeys = o.dense ? (direct_out ? o.keys + s_excl : o.scratch_keys + obase) : o.keys + obase;
    uint32_t* const wcounts = o.dense ? (direct_out ? o.counts + s_excl : o.scratch_counts + obase) : o.counts + obase;
    uint16_t* const wchunk = !o.chunk ? nullptr : (o.dense ? (direct_out ? o.chunk + s_excl : o.scratch_chunk + obase) : o.chunk + obase);
    // ---- emit in key order: slot by slot (lanes read neighbouring slots and, the order being preserved, write
    // neighbouring entries).  Keys that share their home are ranked by a walk from the home slot: they all sit at or
    // behind it with nothing but occupied slots in between, and the walk stops at the last of them.
    {
      const AggTable T{t_key, t_cnt, t_chk, t_hcnt, &s_overflow, o.split_overflow, relmask, bits - 13};
#pragma unroll 1
      for (int j = 0; j < kAggSPT; j++) {
        const uint32_t slot = threadIdx.x + j * kAggThreads;
        const uint32_t k = slot < (uint32_t)kAggTable ? t_key[slot] : kAggEmpty;
        if (k == kAggEmpty) continue;
        const uint32_t h = agg_home(T, k);
        const uint32_t same = (t_hcnt[h >> 2] >> ((h & 3u) * 8)) & 0xFFu;
        uint32_t pos = s_hpre[h];
        if (same > 1) {
          uint32_t seen = 0;
          for (uint32_t x = h; seen < same && x < (uint32_t)kAggTable; x++) {
            const uint32_t kx = t_key[x];
            if (agg_home(T, kx) == h) { seen++; pos += kx < k ? 1u : 0u; }
          }
        }
        if (SPLIT) { s_fin[pos] = (uint16_t)slot; continue; }
        const size_t oo = (size_t)written + pos;
        wkeys[oo] = key_lo + (((uint64_t)p << lb) | k);
        wcounts[oo] = t_cnt[slot];
        if (CHUNK) {
          const uint32_t fc = t_chk[slot];
          if (wchunk) wchunk[oo] = (uint16_t)fc;
          if (o.first_hist) {                       // one atomic per (warp, chunk id): chunk ids of neighbouring keys repeat
            const uint32_t peers = __match_any_sync(__activemask(), fc);
            if ((uint32_t)(threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&o.first_hist[fc & (kAggBuckets - 1)], (uint32_t)__popc(peers));
          }
        }
      }
    }
    if (SPLIT) {
      __syncthreads();
      // two ordered runs: every thread owns a contiguous slice of the sorted keys; rank = number of earlier keys with a
      // non-zero half counter (two block scans per round in total)
      constexpr int SL = (kAggTable + kAggThreads - 1) / kAggThreads;
      const uint32_t per = (nd + blockDim.x - 1) / blockDim.x;
      uint32_t cs[SL];
      uint32_t fa = 0, fb = 0;
      const uint32_t i0 = threadIdx.x * per;
      for (uint32_t j = 0; j < per && j < (uint32_t)SL; j++) {
        uint32_t i = i0 + j;
        cs[j] = i < nd ? t_cnt[s_fin[i]] : 0;
        fa += (cs[j] & 0xFFFFu) ? 1 : 0; fb += (cs[j] >> 16) ? 1 : 0;
      }
      uint32_t ta, tb;
      uint32_t ra = block_exclusive_scan<SumU32>(fa, &ta);
      uint32_t rb = block_exclusive_scan<SumU32>(fb, &tb);
      for (uint32_t j = 0; j < per && j < (uint32_t)SL; j++) {
        uint32_t i = i0 + j;
        if (i >= nd) break;
        uint32_t c = cs[j];
        uint64_t key = key_lo + (((uint64_t)p << lb) | t_key[s_fin[i]]);
        if (c & 0xFFFFu) { size_t oo = obase + written + ra; (o.dense ? o.scratch_keys : o.keys)[oo] = key; (o.dense ? o.scratch_counts : o.counts)[oo] = c & 0xFFFFu; ra++; }
        if (c >> 16) { size_t oo = obase + written2 + rb; (o.dense ? o.scratch_keys2 : o.keys2)[oo] = key; (o.dense ? o.scratch_counts2 : o.counts2)[oo] = c >> 16; rb++; }
      }
      if (threadIdx.x == 0) { s_na = ta; s_nb = tb; }
    }
    __syncthreads();
    if (!SPLIT) written += nd; else { written += s_na; written2 += s_nb; }
    occ_total += s_nocc; occ_total2 += s_nocc2;
    __syncthreads();
  }
  if (!o.dense) {
    if (threadIdx.x == 0) {
      o.n_out[blockIdx.x] = written; o.occ_out[blockIdx.x] = occ_total;
      if (SPLIT) { o.n_out2[blockIdx.x] = written2; o.occ_out2[blockIdx.x] = occ_total2; }
    }
    return;
  }
  // dense: occurrences into the global totals; staged rounds (several key sub-ranges, SPLIT) copied to their place
  if (threadIdx.x == 0) { if (occ_total) atomicAdd(&o.totals[2], (unsigned long long)occ_total); if (SPLIT && occ_total2) atomicAdd(&o.totals[3], (unsigned long long)occ_total2); }
  if (!did_direct) {
    // look back with both counts (A in the low 31 bits, B above)
    if (threadIdx.x < 32) {
      const unsigned long long val = (unsigned long long)written | ((unsigned long long)written2 << 31);
      const unsigned long long excl = agg_lookback(o.status, g, val, threadIdx.x);
      if (threadIdx.x == 0) { s_excl = excl; if (g == o.n_groups - 1) { o.totals[0] = (excl + val) & 0x7FFFFFFFull; o.totals[1] = (excl + val) >> 31; } }
    }
    __syncthreads();
    const size_t oa = (size_t)(s_excl & 0x7FFFFFFFull), ob = (size_t)(s_excl >> 31);
    for (uint32_t i = threadIdx.x; i < written; i += blockDim.x) {
      o.keys[oa + i] = o.scratch_keys[obase + i]; o.counts[oa + i] = o.scratch_counts[obase + i];
      if (o.chunk) o.chunk[oa + i] = SPLIT ? (uint16_t)0 : o.scratch_chunk[obase + i];
    }
    if (SPLIT) for (uint32_t i = threadIdx.x; i < written2; i += blockDim.x) { o.keys2[ob + i] = o.scratch_keys2[obase + i]; o.counts2[ob + i] = o.scratch_counts2[obase + i]; }
  }
}

// Dense gather: the n_out[g] entries of (bin, sub-range) g = p * 2^fixed_bits + s move from
// [s * sub_stride + bin_off[p], ...) to [dense_off[g], ...).
__global__ void __launch_bounds__(256) compact_bins_kernel(const uint32_t* __restrict__ bin_off, const uint32_t* __restrict__ dense_off,
                                                          const uint32_t* __restrict__ n_out, int fixed_bits, size_t sub_stride,
                                                          const uint64_t* __restrict__ in_keys, const uint32_t* __restrict__ in_counts,
                                                          const uint16_t* __restrict__ in_chunk, uint64_t* __restrict__ out_keys,
                                                          uint32_t* __restrict__ out_counts, uint16_t* __restrict__ out_chunk) {
  uint32_t g = blockIdx.x, p = g >> fixed_bits, sub = g & ((1u << fixed_bits) - 1);
  uint32_t n = n_out[g];
  size_t src = (size_t)sub * sub_stride + bin_off[p], dst = dense_off[g];
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
constexpr int kMergeItems = 6;
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
                                       uint64_t* __restrict__ a_split, size_t ntiles, int tile = kMergeTile) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  size_t diag = t * (size_t)tile;
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

// Single-pass variant of the committing merge: no counting pass and no scan.  Tiles take a ticket (so a tile's
// predecessors are always running or done), publish (flag | count) words and find their output offset by a
// warp-wide decoupled look-back over the predecessors' words: flag 1 = the tile's own count, 2 = inclusive prefix.
// All global traffic is coalesced: keys AND counts of both inputs are staged in shared memory, the merged tile is
// assembled there and written out linearly.
// HIST: B carries the first chunk of every key (Bchunk); chunk_hist[c] += keys of B that are NOT in A ("new" keys) whose
// first chunk is c -- the by-product that lets the host find the deNoise boundary chunk of an armed batch.
constexpr unsigned long long kLbValue = (1ULL << 62) - 1;
constexpr int kLbItems = 12;                         // merged elements per thread
constexpr int kLbTile = kMergeThreads * kLbItems;    // ... and per tile: the per-tile costs (merge-path search, block scan, look-back) are paid half as often as with 6
// one array of keys and one of counts: the tile's A part (with the key before it) followed by its B part (with the key
// after it) -- la + lbn = tile, so both inputs fit whatever the split
constexpr size_t kMergeLbSmem = (size_t)(kLbTile + 2) * 8 + (size_t)(kLbTile + 2) * 4;
template <bool HIST>
__global__ void __launch_bounds__(kMergeThreads) merge_tiles_lookback_kernel(
    const uint64_t* __restrict__ A, const uint32_t* __restrict__ Ac, size_t na, const uint64_t* __restrict__ B,
    const uint32_t* __restrict__ Bc, size_t nb, const uint64_t* __restrict__ a_split, size_t ntiles,
    unsigned long long* status /* ntiles words, zeroed */, unsigned int* ticket /* zeroed */, uint64_t* __restrict__ out_keys,
    uint32_t* __restrict__ out_counts, unsigned long long* __restrict__ n_out_total, int* __restrict__ err,
    const uint16_t* __restrict__ Bchunk = nullptr, uint32_t* __restrict__ chunk_hist = nullptr) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  uint64_t* s_k = reinterpret_cast<uint64_t*>(merge_smem);
  uint32_t* s_c = reinterpret_cast<uint32_t*>(s_k + kLbTile + 2);
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_hist[HIST ? kMergeHistMax : 1];
  if (HIST) for (int i = threadIdx.x; i < kMergeHistMax; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const size_t t = s_tile;
  size_t diag0 = t * kLbTile, diag1 = diag0 + kLbTile;
  if (diag1 > na + nb) diag1 = na + nb;
  const size_t a0 = a_split[t], a1 = a_split[t + 1];
  const size_t b0 = diag0 - a0, b1 = diag1 - a1;
  const int la = (int)(a1 - a0), lbn = (int)(b1 - b0);
  const uint64_t kNone = ~0ULL;
  const int total = la + lbn;
  {
    // staging: s_k[0] = A[a0 - 1] (or none), s_k[1 .. la] = the tile's A keys, s_k[la + 1 .. la + lbn] = its B keys,
    // s_k[la + 1 + lbn] = B[b1] (or none: an A key at the very end of the tile may match it); counts alike.  Fixed trip
    // count, loads first and shared-memory stores after, so that all loads of a half are in flight together.
    constexpr int kHalf = (kLbItems + 2) / 2;      // 2 x 7 rounds of 256 entries >= tile + 2
#pragma unroll
    for (int half = 0; half < 2; half++) {
      uint64_t vk[kHalf]; uint32_t vc[kHalf];
#pragma unroll
      for (int j = 0; j < kHalf; j++) {
        const int i = threadIdx.x + (half * kHalf + j) * kMergeThreads;
        uint64_t k = kNone; uint32_t c = 0;
        if (i <= la) {
          if (i > 0) { k = A[a0 + i - 1]; c = Ac[a0 + i - 1]; }
          else if (a0 > 0) k = A[a0 - 1];
        } else if (i < total + 2) {
          const size_t bj = b0 + (size_t)(i - la - 1);
          if (bj < nb) { k = B[bj]; c = Bc[bj]; }
        }
        vk[j] = k; vc[j] = c;
      }
#pragma unroll
      for (int j = 0; j < kHalf; j++) {
        const int i = threadIdx.x + (half * kHalf + j) * kMergeThreads;
        if (i < total + 2) { s_k[i] = vk[j]; s_c[i] = vc[j]; }
      }
    }
  }
  __syncthreads();
  const uint64_t* s_a = s_k;                 // s_a[1 + ai] = A key ai of the tile, s_a[0] = the key before the tile
  const uint64_t* s_b = s_k + la + 1;        // s_b[bi] = B key bi, s_b[lbn] = one key past the tile
  const uint32_t* s_ac = s_c + 1;
  const uint32_t* s_bc = s_c + la + 1;
  int d0 = min(threadIdx.x * kLbItems, total), d1 = min(d0 + kLbItems, total);
  int lo = max(0, d0 - lbn), hi = min(d0, la);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (s_a[1 + mid] <= s_b[d0 - mid - 1]) lo = mid + 1; else hi = mid;
  }
  int ai = lo, bi = d0 - lo;
  uint32_t nheads = 0;
  uint64_t okey[kLbItems]; uint32_t ocnt[kLbItems];
  for (int d = d0; d < d1; d++) {
    bool takeA = (bi >= lbn) || (ai < la && s_a[1 + ai] <= s_b[bi]);
    if (takeA) {
      uint64_t k = s_a[1 + ai];
      uint64_t nextb = s_b[bi];
      uint64_t c = s_ac[ai];
      if (nextb == k) c += s_bc[bi];
      if (c > 0xFFFFFFFFull) { *err = SHQ_ERR_COUNT_OVERFLOW; c = 0xFFFFFFFFull; }
      okey[nheads] = k; ocnt[nheads] = (uint32_t)c;
      nheads++; ai++;
    } else {
      uint64_t k = s_b[bi];
      uint64_t preva = s_a[ai];
      if (preva != k || (ai == 0 && a0 == 0)) {
        okey[nheads] = k; ocnt[nheads] = s_bc[bi]; nheads++;
        if (HIST) { const uint32_t ch = Bchunk[b0 + bi]; if (ch < (uint32_t)kMergeHistMax) atomicAdd(&s_hist[ch], 1u); else atomicAdd(&chunk_hist[ch], 1u); }
      }
      bi++;
    }
  }
  uint32_t tot;
  uint32_t rank = block_exclusive_scan<SumU32>(nheads, &tot);      // ends with a barrier: every thread is done with s_a / s_b, s_hist is complete
  if (HIST) for (int i = threadIdx.x; i < kMergeHistMax; i += blockDim.x) { const uint32_t v = s_hist[i]; if (v) atomicAdd(&chunk_hist[i], v); }
  // the merged tile, assembled in place of the inputs
  uint64_t* s_ok = s_k; uint32_t* s_oc = s_c;
  for (uint32_t i = 0; i < nheads; i++) { s_ok[rank + i] = okey[i]; s_oc[rank + i] = ocnt[i]; }
  if (threadIdx.x < 32) {               // warp 0: publish, then look back 32 predecessors at a time
    const int lane = threadIdx.x;
    unsigned long long excl = 0;
    if (t == 0) {
      if (lane == 0) atomicExch(&status[0], (2ULL << 62) | (unsigned long long)tot);
    } else {
      if (lane == 0) atomicExch(&status[t], (1ULL << 62) | (unsigned long long)tot);
      long long p = (long long)t - 1;   // lane l looks at tile p - l
      for (;;) {
        const long long q = p - lane;
        unsigned long long v = 2ULL << 62;              // tiles before 0: an inclusive prefix of 0
        if (q >= 0) { do { v = *reinterpret_cast<volatile unsigned long long*>(&status[q]); } while ((v >> 62) == 0); }
        const unsigned incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first = incl ? __ffs(incl) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
        unsigned long long part = lane <= first ? (v & kLbValue) : 0;
        for (int d = 16; d; d >>= 1) part += __shfl_down_sync(0xffffffffu, part, d);
        excl += __shfl_sync(0xffffffffu, part, 0);
        if (first < 32) break;
        p -= 32;
      }
      if (lane == 0) atomicExch(&status[t], (2ULL << 62) | (excl + tot));
    }
    if (lane == 0) {
      s_prefix = excl;
      if (t == ntiles - 1) *n_out_total = excl + tot;
    }
  }
  __syncthreads();
  const size_t o = (size_t)s_prefix;
  for (uint32_t i = threadIdx.x; i < tot; i += blockDim.x) { out_keys[o + i] = s_ok[i]; out_counts[o + i] = s_oc[i]; }
}

// ---- the same single-pass merge as a PERSISTENT kernel fed by the TMA engine ------------------------------------
// Each CTA loops over tickets; the four input slices of its NEXT tile (keys and counts of A and B, contiguous in HBM)
// are fetched by bulk asynchronous copies (cp.async.bulk.shared::cluster.global with mbarrier::complete_tx) into the other
// half of a two-stage shared-memory ring while the current tile is merged, so the loads cost no issue slots and their
// latency hides behind the merge of the previous tile.  Bulk copies move multiples of 16 bytes from 16-byte aligned
// addresses: every slice is fetched from the aligned element at or below its first one and indexed with that offset.
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes: a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
}  // namespace tma

struct MergeTileDesc {       // what the consumer needs to know about a staged tile
  unsigned long long t;      // ticket (ntiles = none)
  unsigned long long a0, b0;
  int la, lbn, da, db, dac, dbc;
  int a_first, b_last_real;  // a0 == 0 (no key before the tile) ; b1 < nb (B[b1] exists)
};
constexpr int kMergeTmaKeys = kMergeTile + 8;      // per input, incl. alignment slack
constexpr size_t kMergeTmaStage = (size_t)kMergeTmaKeys * 8 * 2 + (size_t)kMergeTmaKeys * 4 * 2;
constexpr size_t kMergeTmaSmem = 2 * kMergeTmaStage;
template <bool HIST>
__global__ void __launch_bounds__(kMergeThreads) merge_tiles_tma_kernel(
    const uint64_t* __restrict__ A, const uint32_t* __restrict__ Ac, size_t na, const uint64_t* __restrict__ B,
    const uint32_t* __restrict__ Bc, size_t nb, const uint64_t* __restrict__ a_split, size_t ntiles,
    unsigned long long* status /* ntiles words, zeroed */, unsigned int* ticket /* zeroed */, uint64_t* __restrict__ out_keys,
    uint32_t* __restrict__ out_counts, unsigned long long* __restrict__ n_out_total, int* __restrict__ err,
    const uint16_t* __restrict__ Bchunk = nullptr, uint32_t* __restrict__ chunk_hist = nullptr) {
  extern __shared__ __align__(128) unsigned char merge_smem[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ MergeTileDesc s_desc[2];
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_hist[HIST ? kMergeHistMax : 1];
  if (HIST) for (int i = threadIdx.x; i < kMergeHistMax; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x == 0) { tma::mbar_init(&s_bar[0], 1); tma::mbar_init(&s_bar[1], 1); tma::fence_mbar_init(); }
  __syncthreads();
  auto stage_ptr = [&](int st, int which) -> unsigned char* {   // which: 0 keys A, 1 keys B, 2 counts A, 3 counts B
    unsigned char* base = merge_smem + (size_t)st * kMergeTmaStage;
    if (which == 0) return base;
    if (which == 1) return base + (size_t)kMergeTmaKeys * 8;
    if (which == 2) return base + (size_t)kMergeTmaKeys * 16;
    return base + (size_t)kMergeTmaKeys * 16 + (size_t)kMergeTmaKeys * 4;
  };
  // thread 0: take a ticket and start the bulk copies of that tile into stage st
  auto fetch = [&](int st) {
    MergeTileDesc d;
    d.t = atomicAdd(ticket, 1u);
    if (d.t >= ntiles) { d.t = ntiles; s_desc[st] = d; return; }
    size_t diag0 = d.t * kMergeTile, diag1 = diag0 + kMergeTile;
    if (diag1 > na + nb) diag1 = na + nb;
    const size_t a0 = a_split[d.t], a1 = a_split[d.t + 1];
    const size_t b0 = diag0 - a0, b1 = diag1 - a1;
    d.a0 = a0; d.b0 = b0; d.la = (int)(a1 - a0); d.lbn = (int)(b1 - b0);
    d.a_first = a0 == 0; d.b_last_real = b1 < nb;
    // keys A: want A[a0 - 1 .. a1); fetched from the even element at or below a0 - 1 (a0 - 1 itself is patched in when a0 is even)
    const size_t ka = (a0 & ~(size_t)1);
    d.da = (int)(a0 - ka);                                   // keys of the tile start at element da of the fetched slice
    const uint32_t nA = (uint32_t)((d.la + d.da + 1) & ~1);
    const size_t kb = (b0 & ~(size_t)1);
    d.db = (int)(b0 - kb);
    const uint32_t nB = (uint32_t)((d.lbn + 1 + d.db + 1) & ~1);
    const size_t kac = (a0 & ~(size_t)3);
    d.dac = (int)(a0 - kac);
    const uint32_t nAc = (uint32_t)((d.la + d.dac + 3) & ~3);
    const size_t kbc = (b0 & ~(size_t)3);
    d.dbc = (int)(b0 - kbc);
    const uint32_t nBc = (uint32_t)((d.lbn + 1 + d.dbc + 3) & ~3);
    s_desc[st] = d;
    const uint32_t bytes = nA * 8 + nB * 8 + nAc * 4 + nBc * 4;
    tma::mbar_expect_tx(&s_bar[st], bytes);
    // element 2 of the key buffers / element 4.. keep 16-byte alignment: slices land at byte offset 16
    if (nA) tma::bulk_g2s(stage_ptr(st, 0) + 16, A + ka, nA * 8, &s_bar[st]);
    if (nB) tma::bulk_g2s(stage_ptr(st, 1) + 16, B + kb, nB * 8, &s_bar[st]);
    if (nAc) tma::bulk_g2s(stage_ptr(st, 2) + 16, Ac + kac, nAc * 4, &s_bar[st]);
    if (nBc) tma::bulk_g2s(stage_ptr(st, 3) + 16, Bc + kbc, nBc * 4, &s_bar[st]);
  };
  uint32_t phase[2] = {0, 0};
  if (threadIdx.x == 0) fetch(0);
  __syncthreads();
  const uint64_t kNone = ~0ULL;
  for (int cur = 0;; cur ^= 1) {
    const MergeTileDesc d = s_desc[cur];
    if (d.t >= ntiles) break;
    if (threadIdx.x == 0) fetch(cur ^ 1);                    // the next tile streams in while this one is merged
    tma::mbar_wait(&s_bar[cur], phase[cur]); phase[cur] ^= 1;
    // s_a[0] = A[a0 - 1] (or none), s_a[1 + i] = A[a0 + i];  s_b[i] = B[b0 + i], s_b[lbn] = B[b1] (or none)
    uint64_t* s_a = reinterpret_cast<uint64_t*>(stage_ptr(cur, 0) + 16) + d.da - 1;
    uint64_t* s_b = reinterpret_cast<uint64_t*>(stage_ptr(cur, 1) + 16) + d.db;
    uint32_t* s_ac = reinterpret_cast<uint32_t*>(stage_ptr(cur, 2) + 16) + d.dac;
    uint32_t* s_bc = reinterpret_cast<uint32_t*>(stage_ptr(cur, 3) + 16) + d.dbc;
    if (threadIdx.x == 0) {
      if (d.da == 0) s_a[0] = d.a_first ? kNone : A[d.a0 - 1];      // the element below an even a0 is not part of the slice
      else if (d.a_first) s_a[0] = kNone;                            // (cannot happen: a0 == 0 is even)
      if (!d.b_last_real) s_b[d.lbn] = kNone;
    }
    __syncthreads();
    const size_t a0 = d.a0, b0 = d.b0;
    const int la = d.la, lbn = d.lbn;
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
    for (int dd = d0; dd < d1; dd++) {
      bool takeA = (bi >= lbn) || (ai < la && s_a[1 + ai] <= s_b[bi]);
      if (takeA) {
        uint64_t k = s_a[1 + ai];
        uint64_t nextb = s_b[bi];
        uint64_t c = s_ac[ai];
        if (nextb == k) c += s_bc[bi];
        if (c > 0xFFFFFFFFull) { *err = SHQ_ERR_COUNT_OVERFLOW; c = 0xFFFFFFFFull; }
        okey[nheads] = k; ocnt[nheads] = (uint32_t)c;
        nheads++; ai++;
      } else {
        uint64_t k = s_b[bi];
        uint64_t preva = s_a[ai];
        if (preva != k || (ai == 0 && a0 == 0)) {
          okey[nheads] = k; ocnt[nheads] = s_bc[bi]; nheads++;
          if (HIST) { const uint32_t ch = Bchunk[b0 + bi]; if (ch < (uint32_t)kMergeHistMax) atomicAdd(&s_hist[ch], 1u); else atomicAdd(&chunk_hist[ch], 1u); }
        }
        bi++;
      }
    }
    uint32_t tot;
    uint32_t rank = block_exclusive_scan<SumU32>(nheads, &tot);      // ends with a barrier: every thread is done with the inputs
    uint64_t* s_ok = reinterpret_cast<uint64_t*>(stage_ptr(cur, 0) + 16);
    uint32_t* s_oc = reinterpret_cast<uint32_t*>(stage_ptr(cur, 2) + 16);
    for (uint32_t i = 0; i < nheads; i++) { s_ok[rank + i] = okey[i]; s_oc[rank + i] = ocnt[i]; }
    const size_t t = d.t;
    if (threadIdx.x < 32) {               // warp 0: publish, then look back 32 predecessors at a time
      const int lane = threadIdx.x;
      unsigned long long excl = 0;
      if (t == 0) {
        if (lane == 0) atomicExch(&status[0], (2ULL << 62) | (unsigned long long)tot);
      } else {
        if (lane == 0) atomicExch(&status[t], (1ULL << 62) | (unsigned long long)tot);
        long long p = (long long)t - 1;   // lane l looks at tile p - l
        for (;;) {
          const long long q = p - lane;
          unsigned long long v = 2ULL << 62;              // tiles before 0: an inclusive prefix of 0
          if (q >= 0) { do { v = *reinterpret_cast<volatile unsigned long long*>(&status[q]); } while ((v >> 62) == 0); }
          const unsigned incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
          const int first = incl ? __ffs(incl) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
          unsigned long long part = lane <= first ? (v & kLbValue) : 0;
          for (int s2 = 16; s2; s2 >>= 1) part += __shfl_down_sync(0xffffffffu, part, s2);
          excl += __shfl_sync(0xffffffffu, part, 0);
          if (first < 32) break;
          p -= 32;
        }
        if (lane == 0) atomicExch(&status[t], (2ULL << 62) | (excl + tot));
      }
      if (lane == 0) {
        s_prefix = excl;
        if (t == ntiles - 1) *n_out_total = excl + tot;
      }
    }
    __syncthreads();
    const size_t o = (size_t)s_prefix;
    for (uint32_t i = threadIdx.x; i < tot; i += blockDim.x) { out_keys[o + i] = s_ok[i]; out_counts[o + i] = s_oc[i]; }
    __syncthreads();                                         // the stage is free again ...
    if (threadIdx.x == 0) tma::fence_proxy_async();           // ... and the generic-proxy accesses to it are ordered before the next bulk copy
  }
  if (HIST) {
    __syncthreads();
    for (int i = threadIdx.x; i < kMergeHistMax; i += blockDim.x) { const uint32_t v = s_hist[i]; if (v) atomicAdd(&chunk_hist[i], v); }
  }
}

// u32 -> u64 widening (scan input)
__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// ---------------------------------------------------------------- deNoise ---
// An entry survives the sweep unless its count is exactly 1.  The few count-1 entries protected by the range-end
// rule are marked by the host side with count 0 (a value the state never holds) and get their 1 back here.
// No flag array: the tile totals are scanned, then every tile recomputes its flags and writes its survivors.
__global__ void denoise_keep_kernel(const uint32_t* __restrict__ keep_idx, const uint32_t* __restrict__ n_keep, uint32_t keep_cap,
                                    uint32_t* __restrict__ counts) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nk = *n_keep; if (nk > keep_cap) nk = keep_cap;
  if (j < nk) counts[keep_idx[j]] = 0;
}
template <bool WRITE>
__global__ void __launch_bounds__(kScanThreads) denoise_tiles_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, size_t n,
                                                                    uint64_t* __restrict__ tile_cnt, uint64_t* __restrict__ out_keys,
                                                                    uint32_t* __restrict__ out_counts) {
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t c[kScanItems]; uint32_t kept = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) { c[i] = base + i < n ? counts[base + i] : 1; kept += c[i] != 1; }
  uint32_t tot;
  uint32_t rank = block_exclusive_scan<SumU32>(kept, &tot);
  if (!WRITE) { if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot; return; }
  size_t o = tile_cnt[blockIdx.x] + rank;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (c[i] != 1) { out_keys[o] = keys[base + i]; out_counts[o] = c[i] ? c[i] : 1u; o++; }
  }
}

}  // namespace shq
