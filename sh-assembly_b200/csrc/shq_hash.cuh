// shq_hash.cuh -- K1 (FASTQ line index), the parameters shared by the hashing kernels (shq_roll.cuh), the binning
// of raw key streams and the partition-by-owner kernels of the 64-bit key exchange.
//
// Replaces the record scan of reads_to_kmers (reference cqf/CQF_mt.h:610-731).  The first-generation K2 (warp per
// read, prefix-XOR scan: fh_i = rol(P[i+K] ^ P[i], i+K-1)) lived here until the rolling kernels replaced it
// (profiles/r1_ncu_full_v1_raw.csv is its capture).
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

// text must be 16-byte aligned; the batch proper is text[skip, n): the first `skip` (< 16) bytes belong to whatever lies
// before it in the caller's buffer (a batch may start at any byte, its vector loads start at the aligned address below
// it) and bytes past n inside the last vector are masked out.  Positions are reported relative to text + skip.
__global__ void __launch_bounds__(kNlThreads) nl_count_kernel(const char* __restrict__ text, size_t n, uint32_t skip,
                                                             uint32_t* __restrict__ tile_counts) {
  size_t base = (size_t)blockIdx.x * kNlTile + (size_t)threadIdx.x * kNlBytesPerThread;
  int cnt = 0;
  if (base < n) {
    const uint4* v = reinterpret_cast<const uint4*>(text + base);
#pragma unroll
    for (int i = 0; i < kNlBytesPerThread / 16; i++) {
      size_t off = base + 16 * i;
      if (off + 16 <= n && off >= skip) {
        uint4 x = __ldg(v + i);
        cnt += count_nl_u32(x.x) + count_nl_u32(x.y) + count_nl_u32(x.z) + count_nl_u32(x.w);
      } else {
        for (size_t b = off < skip ? skip : off; b < n && b < off + 16; b++) cnt += text[b] == '\n';
      }
    }
  }
  uint32_t tot;
  block_exclusive_scan<SumU32>((uint32_t)cnt, &tot);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kNlThreads) nl_write_kernel(const char* __restrict__ text, size_t n, uint32_t skip,
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
      if (off + 16 <= n && off >= skip) {
        uint4 x = __ldg(v + i);
        words[4 * i] = x.x; words[4 * i + 1] = x.y; words[4 * i + 2] = x.z; words[4 * i + 3] = x.w;
      } else {
        for (size_t b = off < skip ? skip : off; b < n && b < off + 16; b++) {
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
        nl_pos[o++] = (uint32_t)(base + 4 * i + byte) - skip;
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

__device__ __forceinline__ uint32_t chunk_of_line(const HashParams& hp, uint32_t line) {
  if (!hp.chunk_line_end) return 0;
  int lo = 0, hi = hp.nchunks - 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (hp.chunk_line_end[mid] > line) hi = mid; else lo = mid + 1; }
  return (uint32_t)lo;
}

// warp tile of the prefix-XOR window hashing that the unitig enumeration still uses (shq_unitig.cuh):
// fh_i = rol(P[i+K] ^ P[i], i+K-1) over 256 positions per warp
constexpr int kPosPerLane = 8;
constexpr int kTilePos = 32 * kPosPerLane;  // 256 positions per warp tile (K must be <= 224)

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
