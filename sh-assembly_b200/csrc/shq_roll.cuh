// shq_roll.cuh -- K2, third generation: one THREAD per read, rolling canonical ntHash, 8 windows per step.
//
// History (profiles/): the first generation (shq_hash.cuh, warp per read + prefix-XOR scan) spent ~10 warp
// instructions per window; the second (thread per read, 4-byte text loads, four byte-indexed table lookups per
// window) was bound by L1 tag throughput -- every lane of a load touches its own cache line, one tag lookup per
// window (l1tex 84 % busy in roll_count_kernel) -- and the staged scatter spent more instructions on the
// in-round sort than on hashing.  Here:
//   * the reference's recurrence (base/nthash.hpp:305-309) runs on 2-bit base codes: a step loads ONE aligned
//     64-bit word of incoming characters and one of outgoing characters (8 windows per L1 lookup pair), checks
//     with three SIMD byte operations that all eight are A/C/G/T (either case), packs them to 16 bits, and then
//     needs one 16-byte shared-memory load per window: entry (out << 2 | in) of a 16-entry table holds
//     rol(seed(out),K) ^ seed(in)  and  rol(seed'(out),63) ^ rol(seed'(in),K-1);
//   * anything else -- 'N' (restart rule, cqf/CQF_mt.h:624-718), other bytes (which hash as 0 forward but through
//     c & 7 on the reverse strand), the unaligned head and the tail of a read -- takes the byte-exact slow step
//     on the 256-entry tables, for as long as such a byte is inside the window;
//   * the scatter pass ranks every key inside its (round, bin) with the histogram atomic itself, places the
//     round in bin order in shared memory with one pass, and writes it out as one linear sweep in which
//     neighbouring lanes hold neighbouring words of the same bin (full sectors, no per-bin loop).
#pragma once
#include "shq_common.cuh"
#include "shq_hash.cuh"

namespace shq {

struct RollTables {            // in shared memory
  uint4 fast[16];              // [out << 2 | in] over 2-bit codes (c >> 1) & 3: {f.lo, f.hi, r.lo, r.hi}
  uint64_t in_f[256];          // byte-indexed: seed(c)
  uint64_t in_r[256];          // rol(seed'(c & 7), K-1)
  // the outgoing terms rol(seed(c), K) and rol(seed'(c & 7), 63) of the byte-exact step are computed on the fly: that
  // step is rare, and 4 KiB less per CTA lets two scatter CTAs share an SM
  uint8_t bad[256];            // 1 unless c is one of ACGTacgt
};
__device__ __forceinline__ void roll_tables_init(RollTables* t, int K) {
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    uint64_t f = nt_seed_fwd((unsigned char)c), r = nt_seed_rev((unsigned char)c);
    t->in_f[c] = f;
    t->in_r[c] = rotl64(r, K - 1);
    unsigned char u = (unsigned char)(c & 0xDF);
    t->bad[c] = (u == 'A' || u == 'C' || u == 'G' || u == 'T') ? 0 : 1;
  }
  if (threadIdx.x < 16) {
    const unsigned char letter[4] = {'A', 'C', 'T', 'G'};      // code = (c >> 1) & 3
    unsigned char co = letter[threadIdx.x >> 2], ci = letter[threadIdx.x & 3];
    uint64_t f = rotl64(nt_seed_fwd(co), K) ^ nt_seed_fwd(ci);
    uint64_t r = rotl64(nt_seed_rev(co), 63) ^ rotl64(nt_seed_rev(ci), K - 1);
    t->fast[threadIdx.x] = make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
  }
}

// Four text bytes -> four 2-bit codes in one byte (first character in the low bits); *ok is cleared unless all
// four are ACGT in either case.
__device__ __forceinline__ uint32_t pack4(uint32_t w, bool* ok) {
  uint32_t x = (w >> 1) & 0x03030303u;
  uint32_t t = (x >> 1) & ~x & 0x01010101u;                        // code 2 = 'T'
  uint32_t letter = 0x41414141u + (x << 1) + t * 15u;              // A 41, C 43, G 47, T 54
  *ok = *ok && ((w & 0xDFDFDFDFu) == letter);
  return ((x * 0x41041u) >> 18) & 0xFFu;
}
__device__ __forceinline__ uint32_t pack4_nocheck(uint32_t w) {
  uint32_t x = (w >> 1) & 0x03030303u;
  return ((x * 0x41041u) >> 18) & 0xFFu;
}

__device__ __forceinline__ uint64_t rol1(uint64_t v) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  return ((uint64_t)__funnelshift_l(lo, hi, 1) << 32) | __funnelshift_l(hi, lo, 1);
}
__device__ __forceinline__ uint64_t ror1(uint64_t v) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  return ((uint64_t)__funnelshift_r(hi, lo, 1) << 32) | __funnelshift_r(lo, hi, 1);
}

// Per-thread generator of the key stream of one read after another (reads first, first+stride, ...).
// step() advances by one unit -- open+prime a read (1 key), a group of up to 8 windows, or a single
// byte-exact window (<= 1 key) -- and hands every key to emit(key, chunk).  Never more than 8 keys per call.
// Lanes whose reads have the same length run the same sequence of units, i.e. stay converged.
struct Roller {
  const HashParams* hp; const RollTables* tab;
  uint32_t read, read_end, stride;
  uint32_t s, len, pos;        // current read: start offset, length, index of the NEXT char to roll in
  uint32_t slow_until;         // windows whose incoming char has index < slow_until take the byte-exact step
  uint64_t fh, rh; uint32_t chunk;
  // two streams of aligned 64-bit text words: incoming characters (at pos) and outgoing ones (at pos - K);
  // lo holds the word with the current character, hi the next one (loaded one group ahead)
  uint64_t ilo, ihi, olo, ohi;
  const uint64_t* inext; uint32_t odelta;      // next incoming word to load; the outgoing stream trails by odelta words
  const uint64_t* last_word;                   // word holding the last character of the read: nothing beyond is touched
  bool active;                 // a read is open and primed (fh/rh valid for the window ending at pos-1)
  bool cont;                   // the stream registers are valid for pos

  __device__ __forceinline__ void start(const HashParams* hp_, const RollTables* tab_, uint32_t first, uint32_t end, uint32_t stride_) {
    hp = hp_; tab = tab_; read = first; read_end = end; stride = stride_; active = false; cont = false; slow_until = 0;
  }
  // prime at sub-read start s0; false if fewer than K chars remain (cqf/CQF_mt.h:624-625)
  __device__ __forceinline__ bool prime(uint32_t s0) {
    const uint32_t K = hp->K;
    if (s0 > len || len - s0 < K) return false;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(hp->text) + s + s0;
    uint64_t f = 0, r = 0;
    uint32_t anybad = 0;
    // aligned 64-bit words covering [p, p + K)
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(p);
    const uint64_t* wp = reinterpret_cast<const uint64_t*>(a0 & ~(uintptr_t)7);
    uint32_t skip = (uint32_t)(a0 & 7);
    uint32_t i = 0;
    while (i < K) {
      uint64_t w = __ldg(wp++) >> (8 * skip);
      uint32_t nb = 8 - skip; skip = 0;
      if (nb > K - i) nb = K - i;
      for (uint32_t b = 0; b < nb; b++, i++) {
        unsigned char c = (unsigned char)w; w >>= 8;
        f = rol1(f) ^ tab->in_f[c];                      // ends as the xor of rol(seed(c_i), K-1-i)
        r = ror1(r) ^ tab->in_r[c];                      // in_r = rol(seed', K-1): ends as the xor of rol(seed'(c_i), i)
        anybad |= tab->bad[c];
      }
    }
    fh = f; rh = r; pos = s0 + K;
    if (anybad) slow_until = s0 + 2 * K;                 // until every byte of the primed window has left
    cont = false;
    return true;
  }
  __device__ __forceinline__ void roll(unsigned char cout, unsigned char cin) {   // base/nthash.hpp:305-309
    fh = rol1(fh) ^ rotl64(nt_seed_fwd(cout), hp->K) ^ tab->in_f[cin];
    rh = ror1(rh) ^ rotl64(nt_seed_rev(cout), 63) ^ tab->in_r[cin];
  }
  __device__ __forceinline__ uint64_t guarded(const uint64_t* p) const { return p <= last_word ? __ldg(p) : 0; }
  __device__ __forceinline__ void open_streams(const unsigned char* base) {
    const uintptr_t ia = reinterpret_cast<uintptr_t>(base + pos), oa = ia - hp->K;
    const uint64_t* ip = reinterpret_cast<const uint64_t*>(ia & ~(uintptr_t)7);
    const uint64_t* op = reinterpret_cast<const uint64_t*>(oa & ~(uintptr_t)7);
    last_word = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(base + len - 1) & ~(uintptr_t)7);
    odelta = (uint32_t)(ip - op);
    ilo = __ldg(ip); ihi = guarded(ip + 1);
    olo = __ldg(op); ohi = guarded(op + 1);
    inext = ip + 2;
    cont = true;
  }
  // returns false when this thread has no reads left
  // emit(key, chunk): one key; emit8(keys[8], chunk): the eight keys of a full group
  template <class Emit, class Emit8>
  __device__ __forceinline__ bool step(const uint64_t mask, Emit&& emit, Emit8&& emit8) {
    if (!active) {
      if (read >= read_end) return false;
      s = hp->nl_pos[4 * (size_t)read] + 1;
      len = hp->nl_pos[4 * (size_t)read + 1] - s;
      chunk = chunk_of_line(*hp, 4 * read + 1);
      read += stride;
      slow_until = 0;
      if (prime(0)) { active = true; emit((fh < rh ? fh : rh) & mask, chunk); }
      return true;
    }
    if (pos >= len) { active = false; return true; }
    const unsigned char* base = reinterpret_cast<const unsigned char*>(hp->text) + s;
    const uint32_t K = hp->K;
    if (pos >= slow_until) {
      if (!cont) open_streams(base);
      const uint32_t r = len - pos < 8 ? len - pos : 8;
      const uint32_t ish = (uint32_t)(reinterpret_cast<uintptr_t>(base + pos) & 7) * 8;
      const uint32_t osh = (uint32_t)(reinterpret_cast<uintptr_t>(base + pos - K) & 7) * 8;
      uint64_t in8 = ish ? (ilo >> ish) | (ihi << (64 - ish)) : ilo;
      const uint64_t out8 = osh ? (olo >> osh) | (ohi << (64 - osh)) : olo;
      if (r < 8) { const uint64_t m = (1ULL << (8 * r)) - 1; in8 = (in8 & m) | (0x4141414141414141ULL & ~m); }   // past the read: 'A'
      bool ok = true;
      const uint32_t ic = pack4((uint32_t)in8, &ok) | (pack4((uint32_t)(in8 >> 32), &ok) << 8);
      if (ok) {
        // next group's words, one group ahead of their use
        ilo = ihi; ihi = guarded(inext);
        olo = ohi; ohi = guarded(inext - odelta);
        inext++;
        const uint32_t oc = pack4_nocheck((uint32_t)out8) | (pack4_nocheck((uint32_t)(out8 >> 32)) << 8);
        // nibble n of ze / zo = (out << 2 | in) of window 2n / 2n+1
        const uint32_t ze = (ic & 0x3333u) | ((oc & 0x3333u) << 2);
        const uint32_t zo = ((ic >> 2) & 0x3333u) | (oc & 0xCCCCu);
        const uint4* fast = tab->fast;
        if (r == 8) {
          // table entries first (they do not depend on the hashes), then the serial hash chain, then ONE call that
          // hands over all eight keys, so the sink can overlap its shared-memory round trips
          uint64_t k8[8];
#pragma unroll
          for (int half = 0; half < 2; half++) {
            uint4 t[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const int jj = half * 4 + j; t[j] = fast[(((jj & 1) ? zo : ze) >> (4 * (jj >> 1))) & 15u]; }
#pragma unroll
            for (int j = 0; j < 4; j++) {
              fh = rol1(fh) ^ ((uint64_t)t[j].x | ((uint64_t)t[j].y << 32));
              rh = ror1(rh) ^ ((uint64_t)t[j].z | ((uint64_t)t[j].w << 32));
              k8[half * 4 + j] = (fh < rh ? fh : rh) & mask;
            }
          }
          emit8(k8, chunk);
        } else {
          for (uint32_t j = 0; j < r; j++) {               // tail of the read
            const uint4 t = fast[(((j & 1) ? zo : ze) >> (4 * (j >> 1))) & 15u];
            fh = rol1(fh) ^ ((uint64_t)t.x | ((uint64_t)t.y << 32));
            rh = ror1(rh) ^ ((uint64_t)t.z | ((uint64_t)t.w << 32));
            emit((fh < rh ? fh : rh) & mask, chunk);
          }
        }
        pos += r;
        return true;
      }
      slow_until = pos + 8 + K;                           // some byte of this group is not ACGT
    }
    cont = false;
    unsigned char cin = __ldg(base + pos);
    if (cin == 'N') {                                     // restart after the N (CQF_mt.h:680-683)
      const uint32_t su = slow_until;
      if (prime(pos + 1)) { if (su > slow_until) slow_until = su; emit((fh < rh ? fh : rh) & mask, chunk); } else active = false;
      return true;
    }
    if (tab->bad[cin] && pos + 1 + K > slow_until) slow_until = pos + 1 + K;
    roll(__ldg(base + pos - K), cin);
    pos++;
    emit((fh < rh ? fh : rh) & mask, chunk);
    return true;
  }
};

constexpr int kRollThreads = 512;

// Pass 1: global histogram of bins.
// dynamic smem: RollTables + n_bins * 4 (+ chunk map)
template <bool WHOLE>          // WHOLE: unsharded handle, every key is ours
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
  auto emit = [&](uint64_t key, uint32_t) {
    if (WHOLE) { atomicAdd(&s_hist[(uint32_t)(key >> lb)], 1u); return; }
    if (key < key_lo || key >= key_hi) return;
    atomicAdd(&s_hist[(uint32_t)((key - key_lo) >> lb)], 1u);
  };
  auto emit8 = [&](const uint64_t* k, uint32_t ch) {
#pragma unroll
    for (int j = 0; j < 8; j++) emit(k[j], ch);
  };
  while (R.step(mask, emit, emit8)) {}
  __syncthreads();
  for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) { uint32_t v = s_hist[i]; if (v) atomicAdd(&g_hist[i], v); }
}

// Pass 2: staged scatter.  g_cursor[bin] starts at the bin's offset and is shared by all CTAs.
// Per round a CTA
//   A. rolls up to W windows per thread; every key takes its rank inside (round, bin) from the histogram
//      atomic and is parked in shared memory as (packed word, bin << 16 | rank);
//   B. prefix-sums the histogram and reserves one global range per bin (one atomicAdd per bin and round);
//   C. moves every key to slot off[bin] + rank of the sorted round, with its bin beside it;
//   D. writes the sorted round out linearly: word i goes to (base[bin] - off[bin]) + i, so lanes that hold
//      neighbouring words of a bin write neighbouring addresses.
// dynamic smem: RollTables + (3 * n_bins + 1) * 4 + THREADS * W * 14 (+ chunk map)
constexpr int kScatThreads = 768;
constexpr int kScatW = 16;
// capbin > 0: no counting pass was run -- bin b owns the fixed range [b * capbin, (b + 1) * capbin); a reservation that
// does not fit raises *ovf and its words go to a scratch area at slack_base (one round long), so nothing is overwritten;
// the host then redoes the batch with exact offsets.
template <int THREADS, int W, bool WHOLE>
__global__ void __launch_bounds__(THREADS) roll_scatter_kernel(HashParams hp, uint32_t* __restrict__ g_cursor, uint32_t* __restrict__ packed,
                                                               uint32_t capbin = 0, uint32_t slack_base = 0, int* __restrict__ ovf = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RollTables* tab = reinterpret_cast<RollTables*>(smem_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + sizeof(RollTables));
  uint32_t* s_off = s_hist + hp.n_bins;              // n_bins + 1 entries
  uint32_t* s_delta = s_off + hp.n_bins + 1;         // base[bin] - off[bin]
  uint32_t* s_key = s_delta + hp.n_bins;
  uint32_t* s_br = s_key + THREADS * W;
  uint32_t* s_sorted = s_br + THREADS * W;
  uint16_t* s_sbin = reinterpret_cast<uint16_t*>(s_sorted + THREADS * W);
  roll_tables_init(tab, hp.K);
  if (hp.chunk_line_end) {
    uint32_t* s_cle = reinterpret_cast<uint32_t*>(s_sbin + THREADS * W);
    for (int i = threadIdx.x; i < hp.nchunks; i += blockDim.x) s_cle[i] = hp.chunk_line_end[i];
    hp.chunk_line_end = s_cle;
  }
  for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  // reads are dealt round-robin over the whole grid: every bin then receives its words in (roughly) chunk
  // order, which keeps the aggregator's first-chunk minimum almost free of updates
  Roller R;
  R.start(&hp, tab, blockIdx.x * blockDim.x + threadIdx.x, hp.n_reads, gridDim.x * blockDim.x);
  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  const uint64_t key_lo = hp.key_lo, key_hi = hp.key_hi; const int lb = hp.lb;
  const uint32_t lowmask = lb >= 32 ? 0xFFFFFFFFu : ((1u << lb) - 1);
  const int per = (hp.n_bins + THREADS - 1) / THREADS;      // bins per thread in the scan
  bool more = true;
  int produced = 0;
  uint32_t* my_key = s_key + threadIdx.x;                    // slot of this thread's next key: stride THREADS
  const uint32_t br_off = (uint32_t)(s_br - s_key);
  auto emit = [&](uint64_t key, uint32_t ch) {
    if (!WHOLE && (key < key_lo || key >= key_hi)) return;
    uint64_t rel = WHOLE ? key : key - key_lo;
    uint32_t bin = (uint32_t)(rel >> lb);
    my_key[0] = ((uint32_t)rel & lowmask) | (lb >= 32 ? 0 : (ch << lb));
    my_key[br_off] = (bin << 16) | atomicAdd(&s_hist[bin], 1u);
    my_key += THREADS;
    produced++;
  };
  auto emit8 = [&](const uint64_t* k, uint32_t ch) {
    if (!WHOLE) {
#pragma unroll
      for (int j = 0; j < 8; j++) emit(k[j], ch);
      return;
    }
    uint32_t bin[8], rank[8];
#pragma unroll
    for (int j = 0; j < 8; j++) bin[j] = (uint32_t)(k[j] >> lb);
#pragma unroll
    for (int j = 0; j < 8; j++) rank[j] = atomicAdd(&s_hist[bin[j]], 1u);      // eight independent round trips
#pragma unroll
    for (int j = 0; j < 8; j++) {
      my_key[j * THREADS] = ((uint32_t)k[j] & lowmask) | (lb >= 32 ? 0 : (ch << lb));
      my_key[br_off + j * THREADS] = (bin[j] << 16) | rank[j];
    }
    my_key += 8 * THREADS;
    produced += 8;
  };
  for (;;) {
    produced = 0; my_key = s_key + threadIdx.x;
    while (more && produced + 8 <= W) more = R.step(mask, emit, emit8);
    int any = __syncthreads_or(produced > 0);
    if (!any) break;
    {   // B: global reservations first (their round trip runs under the block scan), then the exclusive scan over the
        // bins; the histogram is cleared for the next round.  PER bins per thread (n_bins <= PER * THREADS).
      constexpr int PER = W == 16 ? 6 : 11;     // 4096 bins with 16-key rounds, 8192 with 8-key rounds
      uint32_t cnt[PER], base[PER], sum = 0;
      const int b0 = threadIdx.x * per;
#pragma unroll
      for (int j = 0; j < PER; j++) {
        const int b = b0 + j;
        cnt[j] = (j < per && b < hp.n_bins) ? s_hist[b] : 0;
        base[j] = cnt[j] ? atomicAdd(&g_cursor[b], cnt[j]) : 0;
        sum += cnt[j];
      }
      uint32_t tot;
      uint32_t ex = block_exclusive_scan<SumU32>(sum, &tot);
#pragma unroll
      for (int j = 0; j < PER; j++) {
        const int b = b0 + j;
        if (j < per && b < hp.n_bins) {
          const uint32_t c = cnt[j];
          uint32_t bs = base[j];
          if (capbin && c && (uint64_t)bs + c > (uint64_t)(b + 1) * capbin) { *ovf = 1; bs = slack_base + ex; }
          s_hist[b] = 0;
          s_off[b] = ex;
          s_delta[b] = bs - ex;
          ex += c;
        }
      }
      if (threadIdx.x == 0) s_off[hp.n_bins] = tot;
    }
    __syncthreads();
    const uint32_t total = s_off[hp.n_bins];
    for (int j0 = 0; j0 < produced; j0 += 4) {      // C: four keys at a time, loads before the dependent loads before the stores
      uint32_t br[4], kk[4], off[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int idx = (j0 + u) * THREADS + threadIdx.x; const bool on = j0 + u < produced; br[u] = on ? s_br[idx] : 0; kk[u] = on ? s_key[idx] : 0; }
#pragma unroll
      for (int u = 0; u < 4; u++) off[u] = s_off[br[u] >> 16];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (j0 + u < produced) { const uint32_t dst = off[u] + (br[u] & 0xFFFFu); s_sorted[dst] = kk[u]; s_sbin[dst] = (uint16_t)(br[u] >> 16); }
      }
    }
    __syncthreads();
    for (uint32_t i0 = threadIdx.x; i0 < total; i0 += 4 * THREADS) {   // D
      uint32_t sb[4], dl[4], kk[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const uint32_t i = i0 + u * THREADS; const bool on = i < total; sb[u] = on ? s_sbin[i] : 0; kk[u] = on ? s_sorted[i] : 0; }
#pragma unroll
      for (int u = 0; u < 4; u++) dl[u] = s_delta[sb[u]];
#pragma unroll
      for (int u = 0; u < 4; u++) { const uint32_t i = i0 + u * THREADS; if (i < total) packed[dl[u] + i] = kk[u]; }
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
// upper bound of the number of keys in a batch: sum over the reads of max(len - K + 1, 0) (the N rule only removes keys)
__global__ void __launch_bounds__(256) window_total_kernel(const uint32_t* __restrict__ nl_pos, uint32_t n_reads, int K,
                                                          unsigned long long* __restrict__ total) {
  unsigned long long acc = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
    uint32_t len = nl_pos[4 * (size_t)r + 1] - (nl_pos[4 * (size_t)r] + 1);
    acc += len >= (uint32_t)K ? len - K + 1 : 0;
  }
  for (int d = 16; d; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}
__global__ void fixed_bin_offsets_kernel(uint32_t* __restrict__ bin_off, uint32_t* __restrict__ cursor, int n_bins, uint32_t capbin) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p <= n_bins) { bin_off[p] = (uint32_t)p * capbin; if (p < n_bins) cursor[p] = (uint32_t)p * capbin; }
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
  auto emit8 = [&](const uint64_t* k, uint32_t ch) {
#pragma unroll
    for (int j = 0; j < 8; j++) emit(k[j], ch);
  };
  while (R.step(mask, emit, emit8)) {}
  while (o < e) out[o++] = kNoKey;
}

}  // namespace shq
