// shq_hash.cuh -- K1 (FASTQ line index) and K2 (canonical ntHash of every window) kernels.
//
// Replaces the record scan and the NTPC64 calls of reads_to_kmers
// (reference cqf/CQF_mt.h:610-731, base/nthash.hpp:295-309).
//
// K2 formulation (not the reference's serial roll): with A_j = ror(seed(c_j), j) and
// B_j = rol(seed'(c_j), j), prefix XORs P, Q give every window in O(1):
//   fh_i = rol(P[i+K] ^ P[i], i+K-1),   rh_i = ror(Q[i+K] ^ Q[i], i)
// so one warp hashes a read with a warp XOR-scan and no serial dependency between windows.
#pragma once
#include "shq_common.cuh"
#include "shq_scan.cuh"

namespace shq {

// ------------------------------------------------------------------ K1 ------
constexpr int kNlThreads = 256;
constexpr int kNlBytesPerThread = 64;
constexpr int kNlTile = kNlThreads * kNlBytesPerThread;  // 16 KiB of text per CTA

__device__ __forceinline__ int count_nl_u32(uint32_t w) {
  return __popc(__vcmpeq4(w, 0x0A0A0A0Au)) >> 3;
}

// text must be 16-byte aligned; bytes past n inside the last vector are masked out.
__global__ void __launch_bounds__(kNlThreads) nl_count_kernel(const char* __restrict__ text, size_t n,
                                                             uint32_t* __restrict__ tile_counts) {
  size_t base = (size_t)blockIdx.x * kNlTile + (size_t)threadIdx.x * kNlBytesPerThread;
  int cnt = 0;
  if (base < n) {
    const uint4* v = reinterpret_cast<const uint4*>(text + base);
#pragma unroll
    for (int i = 0; i < kNlBytesPerThread / 16; i++) {
      size_t off = base + 16 * i;
      if (off + 16 <= n) {
        uint4 x = __ldg(v + i);
        cnt += count_nl_u32(x.x) + count_nl_u32(x.y) + count_nl_u32(x.z) + count_nl_u32(x.w);
      } else {
        for (size_t b = off; b < n; b++) cnt += text[b] == '\n';
      }
    }
  }
  uint32_t tot;
  block_exclusive_scan<SumU32>((uint32_t)cnt, &tot);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kNlThreads) nl_write_kernel(const char* __restrict__ text, size_t n,
                                                             const uint32_t* __restrict__ tile_offs,
                                                             uint32_t* __restrict__ nl_pos) {
  size_t base = (size_t)blockIdx.x * kNlTile + (size_t)threadIdx.x * kNlBytesPerThread;
  uint32_t words[kNlBytesPerThread / 4];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < kNlBytesPerThread / 4; i++) words[i] = 0;
  if (base < n) {
    const uint4* v = reinterpret_cast<const uint4*>(text + base);
#pragma unroll
    for (int i = 0; i < kNlBytesPerThread / 16; i++) {
      size_t off = base + 16 * i;
      if (off + 16 <= n) {
        uint4 x = __ldg(v + i);
        words[4 * i] = x.x; words[4 * i + 1] = x.y; words[4 * i + 2] = x.z; words[4 * i + 3] = x.w;
      } else {
        for (size_t b = off; b < n; b++) {
          int k = (int)(b - base);
          words[k >> 2] |= (uint32_t)(unsigned char)text[b] << (8 * (k & 3));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kNlBytesPerThread / 4; i++) cnt += count_nl_u32(words[i]);
  }
  uint32_t tot;
  uint32_t rank = block_exclusive_scan<SumU32>((uint32_t)cnt, &tot);
  if (cnt) {
    uint32_t o = tile_offs[blockIdx.x] + rank;
#pragma unroll
    for (int i = 0; i < kNlBytesPerThread / 4; i++) {
      uint32_t m = __vcmpeq4(words[i], 0x0A0A0A0Au);
      while (m) {
        int byte = (__ffs(m) - 1) >> 3;
        nl_pos[o++] = (uint32_t)(base + 4 * i + byte);
        m &= ~(0xFFu << (8 * byte));
      }
    }
  }
}

// One thread per chunk: chunk c = [starts[c], ends[c]) must be empty or end in '\n' and hold 4m lines.
// line_end[c] = number of newlines with position < ends[c].
__global__ void chunk_validate_kernel(const char* __restrict__ text, const uint32_t* __restrict__ nl_pos,
                                      uint32_t n_lines, const uint64_t* __restrict__ chunk_ends, int nchunks,
                                      uint32_t* __restrict__ line_end, int* __restrict__ err) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint64_t e = chunk_ends[c], s = c ? chunk_ends[c - 1] : 0;
  uint32_t lo = 0, hi = n_lines;
  while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (nl_pos[mid] < e) lo = mid + 1; else hi = mid; }
  line_end[c] = lo;
  uint32_t lo2 = 0; hi = n_lines;
  while (lo2 < hi) { uint32_t mid = lo2 + ((hi - lo2) >> 1); if (nl_pos[mid] < s) lo2 = mid + 1; else hi = mid; }
  if (e > s) {
    if (text[e - 1] != '\n' || ((lo - lo2) & 3u)) atomicExch(err, SHQ_ERR_MALFORMED);
  }
}

// ------------------------------------------------------------------ K2 ------
struct HashParams {
  const char* text;            // batch text
  const uint32_t* nl_pos;      // newline positions
  uint32_t n_reads;            // n_lines / 4
  int K;
  int hb;                      // key bits (key = min(fh,rh) & (2^hb-1))
  int lb;                      // low bits kept in the packed word; bin = (key - key_lo) >> lb
  uint64_t key_lo;             // first key owned by this shard
  uint64_t key_hi;             // one past the last key owned (keys outside are dropped)
  int n_bins;
  const uint32_t* chunk_line_end;  // per chunk, may be null when not armed
  int nchunks;
  uint32_t* chunk_keys;        // per chunk number of keys (count pass, armed only)
  uint32_t* slow_list;         // reads with an 'N' at index >= K
  uint32_t* slow_count;
};

enum { kSinkCount = 0, kSinkScatter = 1, kSinkKeys = 2 };

struct Sink {
  uint32_t* bins;      // count: histogram; scatter: running offsets (shared or global)
  uint32_t* packed;    // scatter output
  uint64_t* keys;      // key-list output
  unsigned long long* nkeys;
  unsigned long long cap;
};

template <int MODE>
__device__ __forceinline__ void sink_key(const HashParams& hp, const Sink& sk, uint64_t key, uint32_t chunk) {
  if (MODE == kSinkKeys) {
    unsigned long long p = atomicAdd(sk.nkeys, 1ULL);
    if (p < sk.cap) sk.keys[p] = key;
    return;
  }
  if (key < hp.key_lo || key >= hp.key_hi) return;
  uint64_t rel = key - hp.key_lo;
  uint32_t bin = (uint32_t)(rel >> hp.lb);
  if (MODE == kSinkCount) {
    atomicAdd(&sk.bins[bin], 1u);
  } else {
    uint32_t pos = atomicAdd(&sk.bins[bin], 1u);
    sk.packed[pos] = (uint32_t)(rel & ((1ULL << hp.lb) - 1)) | (chunk << hp.lb);
  }
}

__device__ __forceinline__ uint32_t chunk_of_line(const HashParams& hp, uint32_t line) {
  if (!hp.chunk_line_end) return 0;
  int lo = 0, hi = hp.nchunks - 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (hp.chunk_line_end[mid] > line) hi = mid; else lo = mid + 1; }
  return (uint32_t)lo;
}

constexpr int kHashWarps = 8;
constexpr int kHashThreads = kHashWarps * 32;
constexpr int kPosPerLane = 8;
constexpr int kTilePos = 32 * kPosPerLane;  // 256 positions per warp tile (K must be <= 224)

// grid = G persistent CTAs; CTA g owns reads [g*R/G, (g+1)*R/G) in BOTH passes so that the
// per-CTA bin offsets computed from the count pass are exact for the scatter pass.
// dynamic smem: n_bins * 4 (hist / offsets) + kHashWarps * (kTilePos+1) * 16 (P,Q prefix arrays)
template <int MODE>
__global__ void __launch_bounds__(kHashThreads) hash_reads_kernel(HashParams hp, Sink sk, uint32_t* __restrict__ hist_matrix,
                                                                 const uint32_t* __restrict__ bin_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_bins = reinterpret_cast<uint32_t*>(smem_raw);
  size_t bins_bytes = MODE == kSinkKeys ? 0 : (((size_t)hp.n_bins * 4 + 15) & ~(size_t)15);
  uint64_t* s_pq = reinterpret_cast<uint64_t*>(smem_raw + bins_bytes);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t* P = s_pq + (size_t)wid * 2 * (kTilePos + 1);
  uint64_t* Q = P + (kTilePos + 1);

  if (MODE == kSinkCount) {
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_bins[i] = 0;
  } else if (MODE == kSinkScatter) {
    const uint32_t* row = hist_matrix + (size_t)blockIdx.x * hp.n_bins;
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_bins[i] = bin_off[i] + row[i];
  }
  __syncthreads();
  Sink lsk = sk;
  if (MODE != kSinkKeys) lsk.bins = s_bins;

  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  const int K = hp.K;
  uint32_t r0 = (uint32_t)(((uint64_t)hp.n_reads * blockIdx.x) / gridDim.x);
  uint32_t r1 = (uint32_t)(((uint64_t)hp.n_reads * (blockIdx.x + 1)) / gridDim.x);
  for (uint32_t r = r0 + wid; r < r1; r += kHashWarps) {
    uint32_t s = hp.nl_pos[4 * (size_t)r] + 1, e = hp.nl_pos[4 * (size_t)r + 1];
    uint32_t len = e - s;
    if (len < (uint32_t)K) continue;
    const unsigned char* seq = reinterpret_cast<const unsigned char*>(hp.text) + s;
    // reads with an upper-case 'N' at index >= K take the serial path (restart rule)
    bool hasN = false;
    for (uint32_t p = K + lane; p < len; p += 32) hasN |= (seq[p] == 'N');
    if (__any_sync(0xffffffffu, hasN)) {
      if (MODE == kSinkCount && lane == 0) hp.slow_list[atomicAdd(hp.slow_count, 1u)] = r;
      continue;
    }
    uint32_t nwin = len - K + 1;
    uint32_t chunk = chunk_of_line(hp, 4 * r + 1);
    if (MODE == kSinkCount && hp.chunk_keys && lane == 0) atomicAdd(&hp.chunk_keys[chunk], nwin);
    const uint32_t TW = kTilePos - K + 1;
    for (uint32_t w0 = 0; w0 < nwin; w0 += TW) {
      uint32_t npos = min((uint32_t)kTilePos, len - w0);
      uint64_t a[kPosPerLane], b[kPosPerLane];
      uint64_t accA = 0, accB = 0;
#pragma unroll
      for (int i = 0; i < kPosPerLane; i++) {
        uint32_t j = lane * kPosPerLane + i;
        uint64_t A = 0, B = 0;
        if (j < npos) {
          unsigned char c = seq[w0 + j];
          A = rotr64(nt_seed_fwd(c), j);
          B = rotl64(nt_seed_rev(c), j);
        }
        accA ^= A; accB ^= B;
        a[i] = accA; b[i] = accB;
      }
      // warp exclusive XOR scan of the lane totals
      uint64_t incA = accA, incB = accB;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint64_t oA = __shfl_up_sync(0xffffffffu, incA, d), oB = __shfl_up_sync(0xffffffffu, incB, d);
        if (lane >= d) { incA ^= oA; incB ^= oB; }
      }
      uint64_t excA = incA ^ accA, excB = incB ^ accB;
      if (lane == 0) { P[0] = 0; Q[0] = 0; }
#pragma unroll
      for (int i = 0; i < kPosPerLane; i++) {
        uint32_t j = lane * kPosPerLane + i;
        P[j + 1] = excA ^ a[i];
        Q[j + 1] = excB ^ b[i];
      }
      __syncwarp();
      uint32_t nw = min(TW, nwin - w0);
      for (uint32_t i = lane; i < nw; i += 32) {
        uint64_t fh = rotl64(P[i + K] ^ P[i], i + K - 1);
        uint64_t rh = rotr64(Q[i + K] ^ Q[i], i);
        uint64_t key = (fh < rh ? fh : rh) & mask;
        sink_key<MODE>(hp, lsk, key, chunk);
      }
      __syncwarp();
    }
  }
  if (MODE == kSinkCount) {
    __syncthreads();
    uint32_t* row = hist_matrix + (size_t)blockIdx.x * hp.n_bins;
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) row[i] = s_bins[i];
  }
}

// Serial restatement of the restart rule for the (rare) reads that contain 'N' past the first K
// bases (cqf/CQF_mt.h:624-718).  One thread per such read; bins live in global memory (row G of the
// histogram matrix in the count pass, absolute running offsets in the scatter pass).
template <int MODE>
__global__ void hash_slow_reads_kernel(HashParams hp, Sink sk, uint32_t n_slow) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_slow) return;
  uint32_t r = hp.slow_list[t];
  uint32_t s = hp.nl_pos[4 * (size_t)r] + 1, e = hp.nl_pos[4 * (size_t)r + 1];
  uint32_t len = e - s;
  const unsigned char* read = reinterpret_cast<const unsigned char*>(hp.text) + s;
  const uint64_t mask = hp.hb >= 64 ? ~0ULL : ((1ULL << hp.hb) - 1);
  const uint32_t K = hp.K;
  uint32_t chunk = chunk_of_line(hp, 4 * r + 1);
  uint32_t emitted = 0;
  uint32_t s0 = 0;
  while (s0 <= len && len - s0 >= K) {
    uint64_t fh = 0, rh = 0;
    for (uint32_t i = 0; i < K; i++) {
      unsigned char c = read[s0 + i];
      fh ^= rotl64(nt_seed_fwd(c), K - 1 - i);
      rh ^= rotl64(nt_seed_rev(c), i);
    }
    sink_key<MODE>(hp, sk, (fh < rh ? fh : rh) & mask, chunk); emitted++;
    bool restart = false;
    for (uint32_t i = s0 + K; i < len; i++) {
      unsigned char cin = read[i], cout = read[i - K];
      if (cin == 'N') { s0 = i + 1; restart = true; break; }
      fh = rotl64(fh, 1) ^ rotl64(nt_seed_fwd(cout), K) ^ nt_seed_fwd(cin);
      rh = rotr64(rh, 1) ^ rotl64(nt_seed_rev(cout), 63) ^ rotl64(nt_seed_rev(cin), K - 1);
      sink_key<MODE>(hp, sk, (fh < rh ? fh : rh) & mask, chunk); emitted++;
    }
    if (!restart) break;
  }
  if (MODE == kSinkCount && hp.chunk_keys) atomicAdd(&hp.chunk_keys[chunk], emitted);
}

// Raw key stream source (shq_insert_keys*): the same two passes over a u64 array.
// count: global histogram.  scatter: each CTA re-counts its strip, reserves one range per bin on the shared
// cursors and writes its keys there.  dynamic smem: 2 * n_bins * 4.
template <int MODE>
__global__ void __launch_bounds__(256) hash_keys_kernel(const uint64_t* __restrict__ keys, size_t n, HashParams hp, Sink sk,
                                                       uint32_t* __restrict__ g_bins, const uint32_t* __restrict__ unused) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_base = s_hist + hp.n_bins;
  for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  size_t i0 = (n * blockIdx.x) / gridDim.x, i1 = (n * (blockIdx.x + 1)) / gridDim.x;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    uint64_t key = keys[i];
    if (key < hp.key_lo || key >= hp.key_hi) continue;
    atomicAdd(&s_hist[(uint32_t)((key - hp.key_lo) >> hp.lb)], 1u);
  }
  __syncthreads();
  if (MODE == kSinkCount) {
    for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) { uint32_t v = s_hist[i]; if (v) atomicAdd(&g_bins[i], v); }
    return;
  }
  for (int i = threadIdx.x; i < hp.n_bins; i += blockDim.x) {
    uint32_t c = s_hist[i];
    s_base[i] = c ? atomicAdd(&g_bins[i], c) : 0;
    s_hist[i] = 0;
  }
  __syncthreads();
  const uint64_t lowmask = (1ULL << hp.lb) - 1;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    uint64_t key = keys[i];
    if (key < hp.key_lo || key >= hp.key_hi) continue;
    uint64_t rel = key - hp.key_lo;
    uint32_t bin = (uint32_t)(rel >> hp.lb);
    uint32_t rank = atomicAdd(&s_hist[bin], 1u);
    sk.packed[s_base[bin] + rank] = (uint32_t)(rel & lowmask);
  }
  (void)unused;
}

// Column-wise exclusive scan of the (rows x n_bins) histogram matrix, in place; tot[p] = column sum.
__global__ void hist_column_scan_kernel(uint32_t* __restrict__ hist, int rows, int n_bins, uint32_t* __restrict__ tot) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_bins) return;
  uint32_t acc = 0;
  for (int g = 0; g < rows; g++) {
    uint32_t v = hist[(size_t)g * n_bins + p];
    hist[(size_t)g * n_bins + p] = acc;
    acc += v;
  }
  tot[p] = acc;
}

// row += bin_off (turns the slow-path row of relative offsets into absolute running offsets)
__global__ void add_bin_off_kernel(uint32_t* __restrict__ row, const uint32_t* __restrict__ bin_off, int n_bins) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_bins) row[p] += bin_off[p];
}


// ---------------------------------------------------- partition by owner ----
// Thread t owns the strip keys[t*S, (t+1)*S).  Pass 1 counts per owner into cnt[owner][t] (owner-major so
// that one exclusive scan over the whole matrix yields the output offsets); pass 2 scatters.
constexpr int kOwnerStrip = 128;
template <bool SCATTER>
__global__ void owner_partition_kernel(const uint64_t* __restrict__ keys, size_t n, int shift, int nparts, size_t nthreads,
                                       uint64_t* __restrict__ mat /* [nparts][nthreads] */, uint64_t* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  size_t b = t * kOwnerStrip, e = b + kOwnerStrip < n ? b + kOwnerStrip : n;
  uint32_t loc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) loc[i] = 0;
  if (!SCATTER) {
    for (size_t i = b; i < e; i++) { uint64_t k = keys[i]; if (k == ~0ULL) continue; int o = (int)(k >> shift); loc[o & 15]++; }
    for (int o = 0; o < nparts; o++) mat[(size_t)o * nthreads + t] = loc[o];
  } else {
    for (size_t i = b; i < e; i++) {
      uint64_t k = keys[i];
      if (k == ~0ULL) continue;      // unused slot of the key list (N restart rule)
      int o = (int)(k >> shift);
      out[mat[(size_t)o * nthreads + t] + loc[o & 15]++] = k;
    }
  }
}

}  // namespace shq
