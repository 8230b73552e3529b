// shq_roll.cuh -- K2, second generation: one THREAD per read, rolling canonical ntHash.
//
// The first generation (shq_hash.cuh, warp per read + prefix-XOR scan) executed ~10 warp instructions per
// window and scattered 4-byte stores into 296 x 16384 open segments (21 GB of DRAM traffic for 2.4 GB of
// algorithmic bytes, profiles/r1_ncu_full_v1_raw.csv).  Here:
//   * hashing is the reference's own recurrence (base/nthash.hpp:305-309) per thread: 4 shared-memory table
//     lookups, two 1-bit rotates and four XORs per window; text is fetched as aligned 32-bit words;
//     the 'N' restart rule (cqf/CQF_mt.h:624-718) is part of the same loop, so there is no slow path;
//   * the scatter pass stages one round of keys per CTA in shared memory, reserves space in every bin with ONE
//     global atomicAdd per (round, bin) on cursors shared by all CTAs, and writes the keys of a bin next to
//     each other, so L2 merges them into full sectors (one open line per bin instead of one per CTA x bin).
#pragma once
#include "shq_common.cuh"
#include "shq_hash.cuh"

namespace shq {

struct RollTables {            // in shared memory, indexed by the text byte
  uint64_t in_f[256];          // seed(c)
  uint64_t out_f[256];         // rol(seed(c), K)
  uint64_t in_r[256];          // rol(seed'(c & 7), K-1)
  uint64_t out_r[256];         // rol(seed'(c & 7), 63)
};
__device__ __forceinline__ void roll_tables_init(RollTables* t, int K) {
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    uint64_t f = nt_seed_fwd((unsigned char)c), r = nt_seed_rev((unsigned char)c);
    t->in_f[c] = f; t->out_f[c] = rotl64(f, K);
    t->in_r[c] = rotl64(r, K - 1); t->out_r[c] = rotl64(r, 63);
  }
}

// Four consecutive text bytes starting at an arbitrary address, from two aligned 32-bit loads.
__device__ __forceinline__ uint32_t load4(const char* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t w0 = __ldg(wp);
  if (sh == 0) return w0;
  return __funnelshift_r(w0, __ldg(wp + 1), sh);
}

// Per-thread generator of the key stream of one read after another (reads first, first+stride, ...).
// step() advances by one unit -- open+prime a read (1 key), a group of 4 windows (4 keys), or a single
// window near the end of a read / around an 'N' (<= 1 key) -- and hands every key to emit(key, chunk).
struct Roller {
  const HashParams* hp; const RollTables* tab;
  uint32_t read, read_end, stride;
  uint32_t s, len, pos;        // current read: start offset, length, index of the NEXT char to roll in
  uint64_t fh, rh; uint32_t chunk;
  bool active;                 // a read is open and primed (fh/rh valid for the window ending at pos-1)

  __device__ __forceinline__ void start(const HashParams* hp_, const RollTables* tab_, uint32_t first, uint32_t end, uint32_t stride_) {
    hp = hp_; tab = tab_; read = first; read_end = end; stride = stride_; active = false;
  }
  // prime at sub-read start s0; false if fewer than K chars remain (cqf/CQF_mt.h:624-625)
  __device__ __forceinline__ bool prime(uint32_t s0) {
    const uint32_t K = hp->K;
    if (s0 > len || len - s0 < K) return false;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(hp->text) + s + s0;
    uint64_t f = 0, r = 0;
    for (uint32_t i = 0; i < K; i++) {
      unsigned char c = __ldg(p + i);
      f = rotl64(f, 1) ^ tab->in_f[c];                 // == xor of rol(seed, K-1-i)
      r ^= rotr64(tab->in_r[c], K - 1 - i);            // in_r = rol(seed', K-1)  ->  rol(seed', i)
    }
    fh = f; rh = r; pos = s0 + K;
    return true;
  }
  __device__ __forceinline__ void roll(unsigned char cout, unsigned char cin) {   // base/nthash.hpp:305-309
    fh = rotl64(fh, 1) ^ tab->out_f[cout] ^ tab->in_f[cin];
    rh = rotr64(rh, 1) ^ tab->out_r[cout] ^ tab->in_r[cin];
  }
  // returns false when this thread has no reads left; never emits more than 4 keys per call
  template <class Emit>
  __device__ __forceinline__ bool step(const uint64_t mask, Emit&& emit) {
    if (!active) {
      if (read >= read_end) return false;
      s = hp->nl_pos[4 * (size_t)read] + 1;
      len = hp->nl_pos[4 * (size_t)read + 1] - s;
      chunk = chunk_of_line(*hp, 4 * read + 1);
      read += stride;
      if (prime(0)) { active = true; emit((fh < rh ? fh : rh) & mask, chunk); }
      return true;
    }
    if (pos >= len) { active = false; return true; }
    const char* base = hp->text + s;
    if (len - pos >= 4) {
      uint32_t vin = load4(base + pos);
      if (__vcmpeq4(vin, 0x4E4E4E4Eu) == 0) {            // no 'N' among the next four characters
        uint32_t vout = load4(base + pos - hp->K);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          roll((unsigned char)(vout >> (8 * j)), (unsigned char)(vin >> (8 * j)));
          emit((fh < rh ? fh : rh) & mask, chunk);
        }
        pos += 4;
        return true;
      }
    }
    unsigned char cin = __ldg(reinterpret_cast<const unsigned char*>(base) + pos);
    if (cin == 'N') {                                     // restart after the N (CQF_mt.h:680-683)
      if (prime(pos + 1)) emit((fh < rh ? fh : rh) & mask, chunk); else active = false;
      return true;
    }
    roll(__ldg(reinterpret_cast<const unsigned char*>(base) + pos - hp->K), cin);
    pos++;
    emit((fh < rh ? fh : rh) & mask, chunk);
    return true;
  }
};

constexpr int kRollThreads = 512;

// Pass 1: global histogram of bins (+ per-chunk key counts when armed).
// dynamic smem: RollTables + n_bins * 4
__global__ void __launch_bounds__(kRollThreads) roll_count_kernel(HashParams hp, uint32_t* __restrict__ g_hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RollTables* tab = reinterpret_cast<RollTables*>(smem_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + sizeof(RollTables));
  roll_tables_init(tab, hp.K);
  for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_hist[i] = 0;
  if (hp.chunk_line_end) {                   // chunk -> line map next to the histogram: the per-read search stays on chip
    uint32_t* s_cle = s_hist + hp.n_bins;
    for (int i = threadIdx.x; i < hp.nchunks; i += blockDim.x) s_cle[i] = hp.chunk_line_end[i];
    hp.chunk_line_end = s_cle;
  }
  __syncthreads();
  // reads are dealt round-robin over the whole grid, so all CTAs walk the text front to back together
  Roller R;
  R.start(&hp, tab, blockIdx.x * blockDim.x + threadIdx.x, hp.n_reads, gridDim.x * blockDim.x);
  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  const uint64_t key_lo = hp.key_lo, key_hi = hp.key_hi; const int lb = hp.lb;
  const bool whole = key_lo == 0 && key_hi > mask;      // unsharded: every key is ours
  uint32_t last_chunk = 0xFFFFFFFFu, run = 0;
  uint32_t* chunk_keys = hp.chunk_keys;
  auto emit = [&](uint64_t key, uint32_t ch) {
    if (chunk_keys) {
      if (ch != last_chunk) { if (run) atomicAdd(&chunk_keys[last_chunk], run); last_chunk = ch; run = 0; }
      run++;
    }
    if (whole) { atomicAdd(&s_hist[(uint32_t)(key >> lb)], 1u); return; }
    if (key < key_lo || key >= key_hi) return;
    atomicAdd(&s_hist[(uint32_t)((key - key_lo) >> lb)], 1u);
  };
  while (R.step(mask, emit)) {}
  if (chunk_keys && run) atomicAdd(&chunk_keys[last_chunk], run);
  __syncthreads();
  for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) { uint32_t v = s_hist[i]; if (v) atomicAdd(&g_hist[i], v); }
}

// Pass 2: staged scatter.  g_cursor[bin] starts at the bin's offset and is shared by all CTAs.
// Per round a CTA (1) rolls up to W windows per thread into shared memory and histograms their bins,
// (2) prefix-sums the histogram and reserves one global range per bin (one atomicAdd per bin and round),
// (3) counting-sorts the round by bin inside shared memory, (4) writes every bin's keys with 8 adjacent
// lanes, i.e. as one 32-byte sector request instead of eight scattered 4-byte stores.
// dynamic smem: RollTables + (3*n_bins + 1) * 4 + kRollThreads * W * (4 + 4 + 2)
template <int THREADS, int W>
__global__ void __launch_bounds__(THREADS) roll_scatter_kernel(HashParams hp, uint32_t* __restrict__ g_cursor, uint32_t* __restrict__ packed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RollTables* tab = reinterpret_cast<RollTables*>(smem_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + sizeof(RollTables));
  uint32_t* s_off = s_hist + hp.n_bins;              // n_bins + 1 entries
  uint32_t* s_base = s_off + hp.n_bins + 1;
  uint32_t* s_key = s_base + hp.n_bins;
  uint32_t* s_sorted = s_key + THREADS * W;
  uint16_t* s_bin = reinterpret_cast<uint16_t*>(s_sorted + THREADS * W);
  roll_tables_init(tab, hp.K);
  if (hp.chunk_line_end) {
    uint32_t* s_cle = reinterpret_cast<uint32_t*>(s_bin + THREADS * W);
    for (int i = threadIdx.x; i < hp.nchunks; i += blockDim.x) s_cle[i] = hp.chunk_line_end[i];
    hp.chunk_line_end = s_cle;
  }
  __syncthreads();
  // reads are dealt round-robin over the whole grid: every bin then receives its words in (roughly) chunk
  // order, which keeps the aggregator's first-chunk minimum almost free of updates
  Roller R;
  R.start(&hp, tab, blockIdx.x * blockDim.x + threadIdx.x, hp.n_reads, gridDim.x * blockDim.x);
  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  const uint64_t key_lo = hp.key_lo, key_hi = hp.key_hi; const int lb = hp.lb;
  const bool whole = key_lo == 0 && key_hi > mask;
  const uint32_t lowmask = lb >= 32 ? 0xFFFFFFFFu : ((1u << lb) - 1);
  const int per = (hp.n_bins + THREADS - 1) / THREADS;      // bins per thread in the scan
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool more = true;
  int produced = 0;
  auto emit = [&](uint64_t key, uint32_t ch) {
    if (!whole && (key < key_lo || key >= key_hi)) return;
    uint64_t rel = key - key_lo;
    uint32_t bin = (uint32_t)(rel >> lb);
    int idx = produced * THREADS + threadIdx.x;
    s_key[idx] = ((uint32_t)rel & lowmask) | (lb >= 32 ? 0 : (ch << lb));
    s_bin[idx] = (uint16_t)bin;
    atomicAdd(&s_hist[bin], 1u);
    produced++;
  };
  for (;;) {
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    produced = 0;
    while (more && produced + 4 <= W) more = R.step(mask, emit);
    for (int j = produced; j < W; j++) s_bin[j * THREADS + threadIdx.x] = 0xFFFF;
    int any = __syncthreads_or(produced > 0);
    if (!any) break;
    {   // exclusive scan over the bins + global reservation
      uint32_t sum = 0;
      int b0 = threadIdx.x * per;
      for (int j = 0; j < per; j++) { int b = b0 + j; if (b < hp.n_bins) sum += s_hist[b]; }
      uint32_t tot;
      uint32_t ex = block_exclusive_scan<SumU32>(sum, &tot);
      for (int j = 0; j < per; j++) {
        int b = b0 + j;
        if (b < hp.n_bins) {
          uint32_t c = s_hist[b];
          s_off[b] = ex; ex += c;
          s_base[b] = c ? atomicAdd(&g_cursor[b], c) : 0;
        }
      }
      if (threadIdx.x == 0) s_off[hp.n_bins] = tot;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < W; j++) {      // counting sort of the round by bin
      int idx = j * THREADS + threadIdx.x;
      uint32_t bin = s_bin[idx];
      if (bin != 0xFFFF) {
        uint32_t rank = atomicSub(&s_hist[bin], 1u) - 1;
        s_sorted[s_off[bin] + rank] = s_key[idx];
      }
    }
    __syncthreads();
    for (int b = wid * 4 + (lane >> 3); b < hp.n_bins; b += (THREADS / 32) * 4) {
      uint32_t off = s_off[b], cnt = s_off[b + 1] - off, base = s_base[b];
      for (uint32_t l = lane & 7; l < cnt; l += 8) packed[base + l] = s_sorted[off + l];
    }
    __syncthreads();
  }
}

// Variant without the in-round sort: keys go straight from production order to base[bin] + rank.  Needs only
// 6 bytes of shared memory per staged key, so two CTAs of 1024 threads fit on an SM and overlap each other's
// barrier phases (reserve / write-out) with hashing.
// dynamic smem: RollTables + 2 * n_bins * 4 + THREADS * W * 6 (+ chunk map)
template <int THREADS, int W>
__global__ void __launch_bounds__(THREADS, 2) roll_scatter_direct_kernel(HashParams hp, uint32_t* __restrict__ g_cursor, uint32_t* __restrict__ packed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RollTables* tab = reinterpret_cast<RollTables*>(smem_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + sizeof(RollTables));
  uint32_t* s_base = s_hist + hp.n_bins;
  uint32_t* s_key = s_base + hp.n_bins;
  uint16_t* s_bin = reinterpret_cast<uint16_t*>(s_key + THREADS * W);
  roll_tables_init(tab, hp.K);
  if (hp.chunk_line_end) {
    uint32_t* s_cle = reinterpret_cast<uint32_t*>(s_bin + THREADS * W);
    for (int i = threadIdx.x; i < hp.nchunks; i += blockDim.x) s_cle[i] = hp.chunk_line_end[i];
    hp.chunk_line_end = s_cle;
  }
  __syncthreads();
  Roller R;
  R.start(&hp, tab, blockIdx.x * blockDim.x + threadIdx.x, hp.n_reads, gridDim.x * blockDim.x);
  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  const uint64_t key_lo = hp.key_lo, key_hi = hp.key_hi; const int lb = hp.lb;
  const bool whole = key_lo == 0 && key_hi > mask;
  const uint32_t lowmask = lb >= 32 ? 0xFFFFFFFFu : ((1u << lb) - 1);
  bool more = true;
  int produced = 0;
  auto emit = [&](uint64_t key, uint32_t ch) {
    if (!whole && (key < key_lo || key >= key_hi)) return;
    uint64_t rel = key - key_lo;
    uint32_t bin = (uint32_t)(rel >> lb);
    int idx = produced * THREADS + threadIdx.x;
    s_key[idx] = ((uint32_t)rel & lowmask) | (lb >= 32 ? 0 : (ch << lb));
    s_bin[idx] = (uint16_t)bin;
    atomicAdd(&s_hist[bin], 1u);
    produced++;
  };
  for (;;) {
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    produced = 0;
    while (more && produced + 4 <= W) more = R.step(mask, emit);
    int any = __syncthreads_or(produced > 0);
    if (!any) break;
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) {
      uint32_t c = s_hist[i];
      s_base[i] = c ? atomicAdd(&g_cursor[i], c) : 0;
    }
    __syncthreads();
    for (int j = 0; j < produced; j++) {
      int idx = j * THREADS + threadIdx.x;
      uint32_t bin = s_bin[idx];
      uint32_t rank = atomicSub(&s_hist[bin], 1u) - 1;
      packed[s_base[bin] + rank] = s_key[idx];
    }
    __syncthreads();
  }
}

// Key list (test hook / multi-GPU exchange): read r writes its keys at out[off[r] ...]; off = exclusive scan of
// max(len-K+1, 0).  Slots that the N rule leaves unused hold the sentinel ~0.
__global__ void read_window_counts_kernel(const uint32_t* __restrict__ nl_pos, uint32_t n_reads, int K, uint64_t* __restrict__ cnt) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  uint32_t len = nl_pos[4 * (size_t)r + 1] - (nl_pos[4 * (size_t)r] + 1);
  cnt[r] = len >= (uint32_t)K ? len - K + 1 : 0;
}
constexpr uint64_t kNoKey = ~0ULL;
__global__ void __launch_bounds__(256) roll_keys_kernel(HashParams hp, const uint64_t* __restrict__ off, uint64_t* __restrict__ out) {
  __shared__ RollTables tab;
  roll_tables_init(&tab, hp.K);
  __syncthreads();
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= hp.n_reads) return;
  uint64_t o = off[r], e = off[r + 1];
  Roller R;
  R.start(&hp, &tab, r, r + 1, 1);
  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  auto emit = [&](uint64_t key, uint32_t) { if (o < e) out[o++] = key; };
  while (R.step(mask, emit)) {}
  while (o < e) out[o++] = kNoKey;
}

}  // namespace shq
