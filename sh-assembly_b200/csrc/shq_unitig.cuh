// shq_unitig.cuh -- Contiger device path (K6-K8): unitig contraction over the CQF image.
//
// Replaces find_unitigs_mt_master / get_unitig_forward / check_unitig / track_kmer_worker /
// build_graph_worker (reference src/contig_assembly.cpp:2034-2172, 3028-3218, 935-1084).
// The reference walks one unitig per thread with 7 dependent CQF probes per base.  Here:
//   U1  every read window whose key is present (count >= s) becomes a node: canonical 2-bit k-mer in a
//       device hash table (exact strings, so two k-mers sharing a CQF key stay two nodes);
//   U3  one thread per node issues the 8 neighbour queries (4 successors per strand) on the image,
//       resolves neighbour ids, and closes over false-positive neighbours that no read contains;
//   U4  simple links (out-degree 1 meeting in-degree 1) form chains; pointer jumping ranks them;
//   U5  chains become unitig sequences in parallel; one strand per unitig is kept;
//   U6  links between unitig ends, connected components (union-find) and the seed rule
//       (a component is reported iff it holds a read-middle k-mer with x <= count <= X).
#pragma once
#include <functional>
#include "shq_common.cuh"
#include "shq_layout.cuh"
#include <chrono>
#include <thread>
#include <unordered_map>

#include "shq_scan.cuh"

namespace shq {

struct Kmer { uint64_t hi, lo; };
__host__ __device__ __forceinline__ bool kmer_less(Kmer a, Kmer b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__host__ __device__ __forceinline__ bool kmer_eq(Kmer a, Kmer b) { return a.hi == b.hi && a.lo == b.lo; }
__host__ __device__ __forceinline__ Kmer kmer_mask(int K) {
  Kmer m;
  int bits = 2 * K;
  m.lo = bits >= 64 ? ~0ULL : ((1ULL << bits) - 1);
  m.hi = bits <= 64 ? 0 : (bits >= 128 ? ~0ULL : ((1ULL << (bits - 64)) - 1));
  return m;
}
__host__ __device__ __forceinline__ Kmer kmer_push_right(Kmer k, uint32_t code, Kmer mask) {   // k[1:] + code
  Kmer r;
  r.hi = ((k.hi << 2) | (k.lo >> 62)) & mask.hi;
  r.lo = ((k.lo << 2) | code) & mask.lo;
  return r;
}
__host__ __device__ __forceinline__ uint32_t kmer_base(Kmer k, int K, int i) {   // code of base i (0 = first)
  int sh = 2 * (K - 1 - i);
  return sh >= 64 ? (uint32_t)((k.hi >> (sh - 64)) & 3) : (uint32_t)((k.lo >> sh) & 3);
}
__host__ __device__ __forceinline__ uint64_t rev_pairs64(uint64_t x) {
  // reverse the order of the 32 two-bit groups
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
  return (x >> 32) | (x << 32);
}
__host__ __device__ __forceinline__ Kmer kmer_rc(Kmer k, int K) {
  // complement (3 - code == ~code on 2 bits), reverse group order over 128 bits, shift down to 2K bits
  uint64_t a = rev_pairs64(~k.lo), b = rev_pairs64(~k.hi);   // 128-bit value (hi=a, lo=b)
  int sh = 128 - 2 * K;
  Kmer r;
  if (sh == 0) { r.hi = a; r.lo = b; }
  else if (sh < 64) { r.lo = (b >> sh) | (a << (64 - sh)); r.hi = a >> sh; }
  else if (sh == 64) { r.lo = a; r.hi = 0; }
  else { r.lo = a >> (sh - 64); r.hi = 0; }
  return r;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__host__ __device__ __forceinline__ uint64_t kmer_hash(Kmer k) { return mix64(k.lo ^ mix64(k.hi + 0x9E3779B97F4A7C15ULL)); }

__device__ __forceinline__ uint64_t seed_of_code(uint32_t c) {
  return c == 0 ? SHQ_SEED_A : c == 1 ? SHQ_SEED_C : c == 2 ? SHQ_SEED_G : SHQ_SEED_T;
}
// ntHash of a packed k-mer (forward strand fh, reverse strand rh), base/nthash.hpp:295-302
__device__ __forceinline__ void kmer_nthash(Kmer k, int K, uint64_t* fh, uint64_t* rh) {
  uint64_t f = 0, r = 0;
  for (int i = 0; i < K; i++) {
    uint32_t c = kmer_base(k, K, i);
    f ^= rotl64(seed_of_code(c), K - 1 - i);
    r ^= rotl64(seed_of_code(3 - c), i);
  }
  *fh = f; *rh = r;
}
// hashes of k[1:] + in, given those of k and out = k[0]  (base/nthash.hpp:305-309)
__device__ __forceinline__ void nthash_roll(uint64_t fh, uint64_t rh, uint32_t out, uint32_t in, int K, uint64_t* nfh, uint64_t* nrh) {
  *nfh = rotl64(fh, 1) ^ rotl64(seed_of_code(out), K) ^ seed_of_code(in);
  *nrh = rotr64(rh, 1) ^ rotl64(seed_of_code(3 - out), 63) ^ rotl64(seed_of_code(3 - in), K - 1);
}

// image lookup without side effects: count of `key`, and the slot index of its entry
__device__ __forceinline__ uint64_t img_lookup(const ImageView& im, uint64_t key, uint64_t* slot) {
  uint32_t rem = (uint32_t)(key & 0xFF);
  uint64_t q = (key >> 8) - im.q_base;
  if (q >= im.nblocks * 64 || !img_bit(im, q, 0)) return 0;
  uint64_t rs = q == 0 ? 0 : img_run_end(im, q - 1) + 1;
  if (rs < q) rs = q;
  uint64_t re = rs;
  while (!img_bit(im, re, 1)) re++;
  uint64_t idx = rs;
  while (idx <= re) {
    uint32_t r; uint64_t c;
    uint64_t e = img_decode(im, idx, re, &r, &c);
    if (r == rem) { if (slot) *slot = idx; return c; }
    if (r > rem) break;
    idx = e + 1;
  }
  return 0;
}

// ------------------------------------------------------------- node table ---
constexpr uint32_t kNil = 0xFFFFFFFFu;
struct NodeTable {
  uint64_t* lo; uint64_t* hi; uint32_t* state; uint32_t* count; uint64_t* fh; uint64_t* rh;
  uint32_t cap_mask;            // capacity - 1 (power of two)
  uint32_t* list;               // node slots in insertion order
  uint32_t* n_nodes;            // counter
  uint32_t max_nodes;
  int* err;
};
// insert-if-absent; returns the slot.  State: 0 empty, 1 being written, 2 ready.
__device__ __forceinline__ uint32_t node_insert(const NodeTable& T, Kmer k, uint32_t count, uint64_t fh, uint64_t rh, bool* fresh) {
  uint32_t h = (uint32_t)kmer_hash(k) & T.cap_mask;
  *fresh = false;
  for (uint32_t probe = 0; probe <= T.cap_mask; probe++) {
    uint32_t s = *reinterpret_cast<volatile uint32_t*>(&T.state[h]);
    if (s == 0) {
      uint32_t old = atomicCAS(&T.state[h], 0u, 1u);
      if (old == 0) {
        T.lo[h] = k.lo; T.hi[h] = k.hi; T.count[h] = count; T.fh[h] = fh; T.rh[h] = rh;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(&T.state[h]) = 2u;
        uint32_t id = atomicAdd(T.n_nodes, 1u);
        if (id < T.max_nodes) T.list[id] = h; else *T.err = SHQ_ERR_CAPACITY;
        *fresh = true;
        return h;
      }
      s = old;
    }
    while (s == 1) s = *reinterpret_cast<volatile uint32_t*>(&T.state[h]);   // writer is between CAS and publish
    __threadfence();
    if (*reinterpret_cast<volatile uint64_t*>(&T.lo[h]) == k.lo && *reinterpret_cast<volatile uint64_t*>(&T.hi[h]) == k.hi) return h;
    h = (h + 1) & T.cap_mask;
  }
  *T.err = SHQ_ERR_CAPACITY;
  return kNil;
}
__device__ __forceinline__ uint32_t node_find(const NodeTable& T, Kmer k) {
  uint32_t h = (uint32_t)kmer_hash(k) & T.cap_mask;
  for (uint32_t probe = 0; probe <= T.cap_mask; probe++) {
    uint32_t s = T.state[h];
    if (s == 0) return kNil;
    if (T.lo[h] == k.lo && T.hi[h] == k.hi) return h;
    h = (h + 1) & T.cap_mask;
  }
  return kNil;
}

// ------------------------------------------------------------------- U1 -----
// One warp per read.  Per tile of 256 positions: 2-bit codes + validity prefix + ntHash prefix XORs in
// shared memory, then every window: key -> image lookup -> first-visitor test on aux[slot] -> node insert.
struct EnumParams {
  const char* text; const uint32_t* nl_pos; uint32_t n_reads; int K; int hb; uint32_t s_min;
};
constexpr int kEnumWarps = 8;

__device__ __forceinline__ uint32_t base_code(unsigned char c, bool* ok) {
  switch (c) {
    case 'A': case 'a': *ok = true; return 0;
    case 'C': case 'c': *ok = true; return 1;
    case 'G': case 'g': *ok = true; return 2;
    case 'T': case 't': *ok = true; return 3;
  }
  *ok = false; return 0;
}

// aux[cqf slot]  = fingerprint of the first k-mer string that touched the entry
// aux2[cqf slot] = node slot of that first visitor (written after its insertion)
// victims        = (node slot, cqf slot) of every further string that shares the entry's key
struct VictimList { uint32_t* node; uint64_t* cqf_slot; uint32_t* n; uint32_t cap; };
__global__ void __launch_bounds__(kEnumWarps * 32) enum_nodes_kernel(EnumParams ep, ImageView im, uint64_t* __restrict__ aux, uint32_t* __restrict__ aux2,
                                                                     VictimList V, NodeTable T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // per warp: P[257], Q[257] (u64), codes[257] (u8), bad[257] (u16 prefix of invalid chars)
  unsigned char* base = smem_raw + (size_t)wid * ((kTilePos + 1) * 16 + 264 + 520);
  uint64_t* P = reinterpret_cast<uint64_t*>(base);
  uint64_t* Q = P + (kTilePos + 1);
  uint16_t* bad = reinterpret_cast<uint16_t*>(Q + (kTilePos + 1));
  uint8_t* codes = reinterpret_cast<uint8_t*>(bad + 260);
  const int K = ep.K;
  const uint64_t mask = ep.hb >= 64 ? ~0ULL : ((1ULL << ep.hb) - 1);
  const Kmer km = kmer_mask(K);
  uint32_t warp_global = blockIdx.x * kEnumWarps + wid, n_warps = gridDim.x * kEnumWarps;
  for (uint32_t r = warp_global; r < ep.n_reads; r += n_warps) {
    uint32_t s = ep.nl_pos[4 * (size_t)r] + 1, e = ep.nl_pos[4 * (size_t)r + 1];
    uint32_t len = e - s;
    if (len < (uint32_t)K) continue;
    const unsigned char* seq = reinterpret_cast<const unsigned char*>(ep.text) + s;
    uint32_t nwin = len - K + 1;
    const uint32_t TW = kTilePos - K + 1;
    for (uint32_t w0 = 0; w0 < nwin; w0 += TW) {
      uint32_t npos = min((uint32_t)kTilePos, len - w0);
      uint64_t a[kPosPerLane], b[kPosPerLane];
      uint64_t accA = 0, accB = 0; uint32_t nbad = 0; uint32_t badloc[kPosPerLane];
#pragma unroll
      for (int i = 0; i < kPosPerLane; i++) {
        uint32_t j = lane * kPosPerLane + i;
        uint64_t A = 0, B = 0; bool ok = true; uint32_t code = 0;
        if (j < npos) {
          unsigned char c = seq[w0 + j];
          code = base_code(c, &ok);
          if (ok) { A = rotr64(seed_of_code(code), j); B = rotl64(seed_of_code(3 - code), j); }
          codes[j] = (uint8_t)code;
        }
        nbad += ok ? 0 : 1;
        badloc[i] = nbad;
        accA ^= A; accB ^= B; a[i] = accA; b[i] = accB;
      }
      uint64_t incA = accA, incB = accB; uint32_t incBad = nbad;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint64_t oA = __shfl_up_sync(0xffffffffu, incA, d), oB = __shfl_up_sync(0xffffffffu, incB, d);
        uint32_t oC = __shfl_up_sync(0xffffffffu, incBad, d);
        if (lane >= d) { incA ^= oA; incB ^= oB; incBad += oC; }
      }
      uint64_t excA = incA ^ accA, excB = incB ^ accB; uint32_t excBad = incBad - nbad;
      if (lane == 0) { P[0] = 0; Q[0] = 0; bad[0] = 0; }
#pragma unroll
      for (int i = 0; i < kPosPerLane; i++) {
        uint32_t j = lane * kPosPerLane + i;
        P[j + 1] = excA ^ a[i]; Q[j + 1] = excB ^ b[i]; bad[j + 1] = (uint16_t)(excBad + badloc[i]);
      }
      __syncwarp();
      uint32_t nw = min(TW, nwin - w0);
      for (uint32_t i = lane; i < nw; i += 32) {
        if (bad[i + K] != bad[i]) continue;                       // window holds a non-ACGT character
        uint64_t fh = rotl64(P[i + K] ^ P[i], i + K - 1);
        uint64_t rh = rotr64(Q[i + K] ^ Q[i], i);
        uint64_t key = (fh < rh ? fh : rh) & mask;
        uint64_t slot = 0;
        uint64_t cnt = img_lookup(im, key, &slot);
        if (cnt < ep.s_min) continue;
        Kmer k{0, 0};
        for (int t = 0; t < K; t++) k = kmer_push_right(k, codes[i + t], km);
        Kmer rc = kmer_rc(k, K);
        bool fwd_canon = !kmer_less(rc, k);
        Kmer cano = fwd_canon ? k : rc;
        uint64_t fp = kmer_hash(cano) | 1ULL;
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&aux[slot]), 0ULL, (unsigned long long)fp);
        if (old == 0ULL || old != fp) {     // first visitor of this key, or a different k-mer sharing the key
          bool fresh;
          uint32_t ns = node_insert(T, cano, (uint32_t)(cnt > 0xFFFFFFFFull ? 0xFFFFFFFFull : cnt), fwd_canon ? fh : rh, fwd_canon ? rh : fh, &fresh);
          if (old == 0ULL) aux2[slot] = ns;
          else if (fresh && ns != kNil) {
            uint32_t vi = atomicAdd(V.n, 1u);
            if (vi < V.cap) { V.node[vi] = ns; V.cqf_slot[vi] = slot; }
          }
        }
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------- U3 -----
// Oriented node id = 2*slot + o; o = 1 means "reverse complement of the stored canonical k-mer".
// outdeg[oid] = number of present successors; succ[oid] = oriented id of the last present one.
struct LinkArrays { uint8_t* outdeg; uint32_t* succ; uint8_t* succ_mask; };

__device__ __forceinline__ void oriented_kmer(const NodeTable& T, uint32_t oid, int K, Kmer* k, uint64_t* fh, uint64_t* rh) {
  uint32_t slot = oid >> 1;
  Kmer c{T.hi[slot], T.lo[slot]};
  if (oid & 1) { *k = kmer_rc(c, K); *fh = T.rh[slot]; *rh = T.fh[slot]; }
  else { *k = c; *fh = T.fh[slot]; *rh = T.rh[slot]; }
}

// processes list[begin, end): both orientations of each node.  New (false-positive) neighbours are
// inserted and will be processed by the next round.
__global__ void link_nodes_kernel(NodeTable T, ImageView im, LinkArrays L, VictimList V, int K, int hb, uint32_t s_min, uint32_t begin, uint32_t end) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t idx = begin + (t >> 1);
  if (idx >= end) return;
  uint32_t oid = 2 * T.list[idx] + (t & 1);
  const uint64_t mask = hb >= 64 ? ~0ULL : ((1ULL << hb) - 1);
  const Kmer km = kmer_mask(K);
  Kmer k; uint64_t fh, rh;
  oriented_kmer(T, oid, K, &k, &fh, &rh);
  uint32_t out = kmer_base(k, K, 0);
  uint32_t deg = 0, last = kNil, pmask = 0;
  for (uint32_t x = 0; x < 4; x++) {
    uint64_t nfh, nrh;
    nthash_roll(fh, rh, out, x, K, &nfh, &nrh);
    uint64_t key = (nfh < nrh ? nfh : nrh) & mask;
    uint64_t cslot = 0;
    uint64_t cnt = img_lookup(im, key, &cslot);
    if (cnt < s_min) continue;
    Kmer n = kmer_push_right(k, x, km);
    Kmer rc = kmer_rc(n, K);
    bool fwd_canon = !kmer_less(rc, n);
    bool fresh;
    uint32_t slot = node_insert(T, fwd_canon ? n : rc, (uint32_t)(cnt > 0xFFFFFFFFull ? 0xFFFFFFFFull : cnt), fwd_canon ? nfh : nrh,
                                fwd_canon ? nrh : nfh, &fresh);
    if (slot == kNil) return;
    if (fresh) {   // a k-mer no read contains: it shares its key with the entry's first visitor (seed blocking rule)
      uint32_t vi = atomicAdd(V.n, 1u);
      if (vi < V.cap) { V.node[vi] = slot; V.cqf_slot[vi] = cslot; }
    }
    deg++; last = 2 * slot + (fwd_canon ? 0 : 1); pmask |= 1u << x;
  }
  L.outdeg[oid] = (uint8_t)deg; L.succ[oid] = last; L.succ_mask[oid] = (uint8_t)pmask;
  if (deg == 0) {
    // dead end: the reference still probes the 3 sibling k-mers y + k[1:] (they would be alternative
    // predecessors of a successor) and spawns the present ones (contig_assembly.cpp:3094-3165), so a sibling
    // that exists only as a false positive of the filter must become a node as well.
    for (uint32_t y = 0; y < 4; y++) {
      if (y == out) continue;
      int sh = 2 * (K - 1);
      Kmer sib = k;
      if (sh >= 64) sib.hi = (sib.hi & ~(3ULL << (sh - 64))) | ((uint64_t)y << (sh - 64));
      else sib.lo = (sib.lo & ~(3ULL << sh)) | ((uint64_t)y << sh);
      uint64_t sfh, srh;
      kmer_nthash(sib, K, &sfh, &srh);
      uint64_t key = (sfh < srh ? sfh : srh) & mask;
      uint64_t cslot = 0;
      uint64_t cnt = img_lookup(im, key, &cslot);
      if (cnt < s_min) continue;
      Kmer rc = kmer_rc(sib, K);
      bool fwd_canon = !kmer_less(rc, sib);
      bool fresh;
      uint32_t ns = node_insert(T, fwd_canon ? sib : rc, (uint32_t)(cnt > 0xFFFFFFFFull ? 0xFFFFFFFFull : cnt), fwd_canon ? sfh : srh, fwd_canon ? srh : sfh, &fresh);
      if (fresh && ns != kNil) {
        uint32_t vi = atomicAdd(V.n, 1u);
        if (vi < V.cap) { V.node[vi] = ns; V.cqf_slot[vi] = cslot; }
      }
    }
  }
}

// ------------------------------------------------------------------- U4 -----
// next/prev over oriented ids.  c -> n is simple iff outdeg(c) == 1 and outdeg(RC(n)) == 1.
__global__ void simple_links_kernel(NodeTable T, LinkArrays L, uint32_t n_nodes, uint32_t* __restrict__ next, uint32_t* __restrict__ prev) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  if (L.outdeg[oid] != 1) return;
  uint32_t n = L.succ[oid];
  if (L.outdeg[n ^ 1] != 1) return;
  next[oid] = n; prev[n] = oid;
}
// jd[oid] = (jump << 32) | dist; jump = some ancestor (towards the head), dist = distance to it.
__global__ void rank_init_kernel(NodeTable T, uint32_t n_nodes, const uint32_t* __restrict__ prev, unsigned long long* __restrict__ jd) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  uint32_t p = prev[oid];
  jd[oid] = p == kNil ? ((unsigned long long)oid << 32) : (((unsigned long long)p << 32) | 1ULL);
}
__global__ void rank_jump_kernel(NodeTable T, uint32_t n_nodes, unsigned long long* __restrict__ jd, uint32_t* __restrict__ changed) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  unsigned long long me = *reinterpret_cast<volatile unsigned long long*>(&jd[oid]);
  uint32_t j = (uint32_t)(me >> 32);
  if (j == oid) return;
  unsigned long long pj = *reinterpret_cast<volatile unsigned long long*>(&jd[j]);
  uint32_t jj = (uint32_t)(pj >> 32);
  if (jj == j) return;                       // j is a head: done
  unsigned long long nd = (unsigned long long)(uint32_t)me + (unsigned long long)(uint32_t)pj;
  *reinterpret_cast<volatile unsigned long long*>(&jd[oid]) = ((unsigned long long)jj << 32) | (nd & 0xFFFFFFFFull);
  *changed = 1;
}
// cycles: nodes whose jump target is not a head after ranking.  minid propagates the smallest oriented id.
__global__ void cycle_flag_kernel(NodeTable T, uint32_t n_nodes, const unsigned long long* __restrict__ jd, const uint32_t* __restrict__ prev,
                                  uint32_t* __restrict__ minid, uint32_t* __restrict__ jmp, uint32_t* __restrict__ n_cycle) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  uint32_t j = (uint32_t)(jd[oid] >> 32);
  bool in_cycle = prev[j] != kNil;          // the final jump target of a chain node is its head (prev == nil)
  minid[oid] = in_cycle ? oid : kNil;
  jmp[oid] = in_cycle ? prev[oid] : kNil;
  if (in_cycle) atomicAdd(n_cycle, 1u);
}
__global__ void cycle_min_kernel(NodeTable T, uint32_t n_nodes, uint32_t* __restrict__ minid, uint32_t* __restrict__ jmp) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  uint32_t j = jmp[oid];
  if (j == kNil) return;
  uint32_t m = min(*reinterpret_cast<volatile uint32_t*>(&minid[oid]), *reinterpret_cast<volatile uint32_t*>(&minid[j]));
  minid[oid] = m;
  jmp[oid] = *reinterpret_cast<volatile uint32_t*>(&jmp[j]);
}
// Cut every cycle so that the two strands of a circular unitig become reverse-complementary chains:
// the strand holding the globally smallest oriented id m starts at m; the other strand ends at m^1.
__global__ void cycle_cut_kernel(NodeTable T, uint32_t n_nodes, const uint32_t* __restrict__ minid, uint32_t* __restrict__ next, uint32_t* __restrict__ prev) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  uint32_t M = minid[oid];
  if (M == kNil) return;
  uint32_t Mr = minid[oid ^ 1];
  if (M <= Mr) {
    if (oid == M) { uint32_t p = prev[oid]; prev[oid] = kNil; if (p != kNil) next[p] = kNil; }
  } else if (oid == (Mr ^ 1)) {
    uint32_t n = next[oid]; next[oid] = kNil; if (n != kNil) prev[n] = kNil;
  }
}

// ------------------------------------------------------------------- U5 -----
// tails report (tail, length) to their head; heads decide which strand is kept.
__global__ void tails_kernel(NodeTable T, uint32_t n_nodes, const uint32_t* __restrict__ next, const unsigned long long* __restrict__ jd,
                             uint32_t* __restrict__ tail_of, uint32_t* __restrict__ len_of) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  if (next[oid] != kNil) return;
  unsigned long long me = jd[oid];
  uint32_t head = (uint32_t)(me >> 32);
  tail_of[head] = oid; len_of[head] = (uint32_t)me + 1;
}
// unitig ids for kept heads: head <= RC(tail).  uid_of_head[oid] = id or kNil.
__global__ void heads_kernel(NodeTable T, uint32_t n_nodes, const uint32_t* __restrict__ prev, const uint32_t* __restrict__ tail_of,
                             const uint32_t* __restrict__ len_of, int K, uint32_t* __restrict__ uid_of_head, uint32_t* __restrict__ n_unitigs,
                             uint32_t* __restrict__ u_head, uint64_t* __restrict__ u_seqlen, uint64_t* __restrict__ u_klen, uint32_t max_unitigs) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  if (prev[oid] != kNil) return;
  uint32_t tail = tail_of[oid];
  if (oid > (tail ^ 1)) return;              // the reverse strand's chain is the kept one
  uint32_t id = atomicAdd(n_unitigs, 1u);
  if (id >= max_unitigs) return;
  uid_of_head[oid] = id; u_head[id] = oid;
  u_seqlen[id] = (uint64_t)len_of[oid] + K - 1; u_klen[id] = len_of[oid];
}
// every node of a kept chain writes its last base (the head also its first K-1) and its count
__global__ void assemble_kernel(NodeTable T, uint32_t n_nodes, const unsigned long long* __restrict__ jd, const uint32_t* __restrict__ uid_of_head,
                                const uint64_t* __restrict__ seq_off, const uint64_t* __restrict__ k_off, int K, char* __restrict__ seq,
                                uint32_t* __restrict__ kcounts, uint32_t* __restrict__ unitig_of_oid) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 1) >= n_nodes) return;
  uint32_t oid = 2 * T.list[t >> 1] + (t & 1);
  unsigned long long me = jd[oid];
  uint32_t head = (uint32_t)(me >> 32), dist = (uint32_t)me;
  uint32_t uid = uid_of_head[head];
  unitig_of_oid[oid] = uid;
  if (uid == kNil) return;
  Kmer k; uint64_t fh, rh;
  oriented_kmer(T, oid, K, &k, &fh, &rh);
  const char B[4] = {'A', 'C', 'G', 'T'};
  uint64_t so = seq_off[uid];
  seq[so + K - 1 + dist] = B[kmer_base(k, K, K - 1)];
  if (dist == 0) for (int i = 0; i < K - 1; i++) seq[so + i] = B[kmer_base(k, K, i)];
  kcounts[k_off[uid] + dist] = T.count[oid >> 1];
}

// ------------------------------------------------------------------- U6 -----
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  for (;;) {
    uint32_t p = *reinterpret_cast<volatile uint32_t*>(&parent[x]);
    if (p == x) return x;
    uint32_t gp = *reinterpret_cast<volatile uint32_t*>(&parent[p]);
    if (gp != p) parent[x] = gp;
    x = p;
  }
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { uint32_t t = a; a = b; b = t; }
    if (atomicCAS(&parent[b], b, a) == b) return;
  }
}
// signed unitig reference of the chain headed by oriented node n: +(id+1) if that strand is kept,
// -(id+1) if the kept strand is its reverse (whose head is RC(tail)).
__device__ __forceinline__ int unitig_ref(uint32_t n, const uint32_t* uid_of_head, const uint32_t* tail_of) {
  uint32_t u = uid_of_head[n];
  if (u != kNil) return (int)(u + 1);
  uint32_t tl = tail_of[n];
  if (tl == kNil) return 0;                 // n is not a chain head (cannot happen for a non-simple link)
  uint32_t rh = tl ^ 1;
  u = uid_of_head[rh];
  return u == kNil ? 0 : -(int)(u + 1);
}
// per unitig: links after (successors of the tail, order A,C,G,T) and before (successors of RC(head),
// order T,G,C,A -- build_graph_worker, contig_assembly.cpp:1030-1060); union with every linked unitig and,
// for dead ends, with the unitigs of sibling k-mers (the reference spawns those, :3094-3165).
__global__ void unitig_links_kernel(NodeTable T, ImageView im, LinkArrays L, int K, int hb, uint32_t s_min, uint32_t n_unitigs,
                                    const uint32_t* __restrict__ u_head, const uint32_t* __restrict__ tail_of,
                                    const uint32_t* __restrict__ uid_of_head, const uint32_t* __restrict__ unitig_of_oid,
                                    int* __restrict__ links /* [n_unitigs][8] */, uint32_t* __restrict__ parent) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t u = t >> 1;
  if (u >= n_unitigs) return;
  int side = t & 1;                                 // 0: after (tail), 1: before (RC(head))
  uint32_t head = u_head[u];
  uint32_t e = side == 0 ? tail_of[head] : (head ^ 1);
  const uint64_t mask = hb >= 64 ? ~0ULL : ((1ULL << hb) - 1);
  const Kmer km = kmer_mask(K);
  Kmer k; uint64_t fh, rh;
  oriented_kmer(T, e, K, &k, &fh, &rh);
  uint32_t out = kmer_base(k, K, 0);
  int nl = 0;
  for (int xi = 0; xi < 4; xi++) {
    uint32_t x = side == 0 ? (uint32_t)xi : (uint32_t)(3 - xi);
    links[(size_t)u * 8 + side * 4 + xi] = 0;
    if (!((L.succ_mask[e] >> x) & 1)) continue;
    Kmer n = kmer_push_right(k, x, km);
    Kmer rc = kmer_rc(n, K);
    bool fwd_canon = !kmer_less(rc, n);
    uint32_t slot = node_find(T, fwd_canon ? n : rc);
    if (slot == kNil) continue;
    uint32_t noid = 2 * slot + (fwd_canon ? 0 : 1);
    int ref = unitig_ref(noid, uid_of_head, tail_of);
    links[(size_t)u * 8 + side * 4 + nl] = ref; nl++;
    if (ref) uf_union(parent, u, (uint32_t)(abs(ref) - 1));
  }
  if (L.outdeg[e] == 0) {                           // dead end: siblings y + e[1:], y != e[0]
    // sibling = predecessor-side variant: replace the first base.  hashes: recompute from the packed k-mer
    for (uint32_t y = 0; y < 4; y++) {
      if (y == out) continue;
      int sh = 2 * (K - 1);
      Kmer sib = k;
      if (sh >= 64) sib.hi = (sib.hi & ~(3ULL << (sh - 64))) | ((uint64_t)y << (sh - 64));
      else sib.lo = (sib.lo & ~(3ULL << sh)) | ((uint64_t)y << sh);
      Kmer rc = kmer_rc(sib, K);
      bool fwd_canon = !kmer_less(rc, sib);
      uint32_t slot = node_find(T, fwd_canon ? sib : rc);
      if (slot == kNil) continue;                    // not a node: either absent or never reached -> check presence
      uint32_t soid = 2 * slot + (fwd_canon ? 0 : 1);
      uint32_t su = unitig_of_oid[soid];
      if (su == kNil) su = unitig_of_oid[soid ^ 1];
      if (su != kNil) uf_union(parent, u, su);
    }
  }
  (void)mask; (void)im; (void)s_min; (void)fh; (void)rh;
}

// seeds: K ASCII bytes each (already upper-cased, no 'N').  Per seed: the node slot and the component root of
// its unitig if x <= count <= X, else kNil.  The t=1 "already traveled" rule is applied on the host in read order.
__global__ void seeds_kernel(NodeTable T, const char* __restrict__ seeds, uint32_t n_seeds, int K, uint32_t x_min, uint32_t x_max,
                             const uint32_t* __restrict__ unitig_of_oid, uint32_t* __restrict__ parent, uint32_t* __restrict__ seed_node,
                             uint32_t* __restrict__ seed_root) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  seed_node[i] = kNil; seed_root[i] = kNil;
  const Kmer km = kmer_mask(K);
  Kmer k{0, 0};
  for (int t = 0; t < K; t++) {
    bool ok; uint32_t c = base_code((unsigned char)seeds[(size_t)i * K + t], &ok);
    if (!ok) return;
    k = kmer_push_right(k, c, km);
  }
  Kmer rc = kmer_rc(k, K);
  bool fwd_canon = !kmer_less(rc, k);
  uint32_t slot = node_find(T, fwd_canon ? k : rc);
  if (slot == kNil) return;
  uint32_t cnt = T.count[slot];
  if (cnt < x_min || cnt > x_max) return;
  uint32_t u = unitig_of_oid[2 * slot];
  if (u == kNil) u = unitig_of_oid[2 * slot + 1];
  if (u == kNil) return;
  seed_node[i] = slot; seed_root[i] = uf_find(parent, u);
}
// victims: resolve (victim node, first-visitor node) -> component roots of both
__global__ void victims_kernel(VictimList V, uint32_t nv, const uint32_t* __restrict__ aux2, const uint32_t* __restrict__ unitig_of_oid,
                               uint32_t* __restrict__ parent, uint32_t* __restrict__ out /* [nv][4]: vnode, vroot, pnode, proot */) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  uint32_t vn = V.node[i], pn = aux2[V.cqf_slot[i]];
  auto root_of = [&](uint32_t slot) -> uint32_t {
    if (slot == kNil) return kNil;
    uint32_t u = unitig_of_oid[2 * slot];
    if (u == kNil) u = unitig_of_oid[2 * slot + 1];
    return u == kNil ? kNil : uf_find(parent, u);
  };
  out[4 * i] = vn; out[4 * i + 1] = root_of(vn); out[4 * i + 2] = pn; out[4 * i + 3] = root_of(pn);
}
__global__ void roots_kernel(uint32_t* parent, uint32_t n, uint32_t* root) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) root[i] = uf_find(parent, i);
}
__global__ void init_iota_kernel(uint32_t* p, uint32_t n) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

}  // namespace shq

// Seed candidates on the device: the middle k-mer (worker rule, len/2 - K/2) of every record of a REGULAR chunk
// (every record is 4 lines, header starts with '@').  One slot of K bytes per read, slot[0] = 0 when the read has no
// seed (too short, or an 'N' in the seed).  *irregular is raised when the reference's line cursor
// (contig_assembly.cpp:1843-1921, whose `continue`s skip skipLines(2)) could parse the chunk differently:
// a header not starting with '@', or an N-seed whose quality line starts with '@'.
__global__ void seed_extract_kernel(const char* __restrict__ text, const uint32_t* __restrict__ nl_pos, uint32_t n_reads, int K,
                                    char* __restrict__ seeds, int* __restrict__ irregular, int master_rule) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  uint32_t h0 = r ? nl_pos[4 * (size_t)r - 1] + 1 : 0;
  uint32_t s = nl_pos[4 * (size_t)r] + 1, e = nl_pos[4 * (size_t)r + 1];
  uint32_t q0 = nl_pos[4 * (size_t)r + 2] + 1, q1 = nl_pos[4 * (size_t)r + 3];
  char* out = seeds + (size_t)r * K;
  out[0] = 0;
  if (nl_pos[4 * (size_t)r] == h0 || text[h0] != '@') { *irregular = 1; return; }     // empty or foreign header line
  int len = (int)(e - s);
  if (len < K) return;
  int middle = master_rule ? len / 2 : len / 2 - K / 2;      // contig_assembly.cpp:2067-2071 (master) vs 1859-1863 (workers)
  if (middle > len - K) return;
  bool hasN = false;
  for (int i = 0; i < K; i++) {
    signed char c = (signed char)text[s + middle + i];
    if (c > 0x60) c -= 32;                                   // to_upper_DNA, base/Utility.h:120-155
    if (c == 'N') hasN = true;
    out[i] = (char)c;
  }
  if (hasN) { out[0] = 0; if (q1 > q0 && text[q0] == '@') *irregular = 1; }
}

// ============================================================ host side ======
struct shq_unitigs_s {
  shq_t h = nullptr;
  shq_unitig_opts opts{};
  uint32_t cap = 0, max_nodes = 0;
  shq::DevBuf lo, hi, state, count, fh, rh, list, nn, aux, outdeg, succ, smask, next, prev, jd, minid, jmp, tail_of, len_of, uid_of_head,
      unitig_of_oid, u_head, u_seqlen, u_klen, seq, kcounts, links, parent, reached, keep, d_seeds, d_text, misc, aux2, v_node, v_slot, v_n, v_out,
      seed_node, seed_root, roots;
  uint32_t victim_cap = 0;
  std::vector<char> seeds;        // host scratch (irregular chunks only)
  shq::DevBuf d_seed_slots; size_t n_seed_slots = 0;   // device: K bytes per seed candidate, in read order
  shq::DevBuf d_stage; size_t st_fill = 0; std::vector<size_t> st_ends;   // read chunks staged on the device
  size_t st_cap_target = (size_t)128 << 20;
  bool built = false;
  uint64_t global_ndistinct = 0;  // sharded filter: distinct keys over all shards (sizes the node table)
  bool master_rule = false;       // seed k-mer of a read: len/2 - K/2 (workers, the default) or len/2 (the master thread)
  // results (host)
  std::vector<std::string> seqs; std::vector<int> km; std::vector<std::vector<int>> after, before;
  std::string fasta;
  uint64_t n_queries = 0;
};

namespace {

using namespace shq;

NodeTable node_table(shq_unitigs_s* u) {
  NodeTable T;
  T.lo = u->lo.as<uint64_t>(); T.hi = u->hi.as<uint64_t>(); T.state = u->state.as<uint32_t>(); T.count = u->count.as<uint32_t>();
  T.fh = u->fh.as<uint64_t>(); T.rh = u->rh.as<uint64_t>(); T.cap_mask = u->cap - 1; T.list = u->list.as<uint32_t>();
  T.n_nodes = u->nn.as<uint32_t>(); T.max_nodes = u->max_nodes; T.err = u->h->d_err;
  return T;
}

// processDataChunk's line cursor (contig_assembly.cpp:1839-1924), restated literally: every `continue`
// skips the two skipLines(), so the quality line is examined as a would-be header.  Collects the
// upper-cased middle k-mer (worker rule: len/2 - K/2) of every record that passes the 'N' test.
void collect_seeds_range(int K, const char* text, size_t p, size_t n, std::vector<char>* out, bool master_rule) {
  auto read_line = [&](const char** s, size_t* len) -> bool {
    if (p >= n) return false;
    const char* nl = (const char*)memchr(text + p, '\n', n - p);
    size_t e = nl ? (size_t)(nl - text) : n;
    *s = text + p; *len = e - p;
    p = nl ? e + 1 : n;
    return true;
  };
  const char* line; size_t ll; const char* seq; size_t sl;
  char kmer[1024];
  while (read_line(&line, &ll)) {
    if (ll == 0 || line[0] != '@') continue;
    if (!read_line(&seq, &sl)) break;
    if ((int)sl < K) { const char* t; size_t tl; read_line(&t, &tl); read_line(&t, &tl); continue; }
    int seq_len = (int)sl, middle = master_rule ? seq_len / 2 : seq_len / 2 - K / 2;
    if (middle <= seq_len - K) {
      bool hasN = false;
      for (int i = 0; i < K; i++) { char c = seq[middle + i]; if (c > 0x60) c -= 32; if (c == 'N') hasN = true; kmer[i] = c; }
      if (hasN) continue;
      // the count / traveled tests happen on the device; rejected seeds `continue` in the reference too,
      // i.e. without skipLines(2) -- but so does every accepted-then-skipped path except full processing.
      // A seed that is processed ends with skipLines(2); one that is rejected does not.  Which of the two
      // applies is only known on the device, so the cursor follows the processed path here (regular
      // 4-line records make both paths parse identically unless a quality line starts with '@').
      out->insert(out->end(), kmer, kmer + K);
    }
    const char* t; size_t tl; read_line(&t, &tl); read_line(&t, &tl);
  }
}

// Record start at or after `from`: a line starting with '@' whose line+2 starts with '+', the test
// fastq_read_parts itself uses to cut chunks (cqf/CQF_mt.h:796-806).
size_t find_record_start(const char* text, size_t from, size_t n) {
  size_t p = from;
  if (p > 0) { const char* nl = (const char*)memchr(text + p - 1, '\n', n - p + 1); if (!nl) return n; p = (size_t)(nl - text) + 1; }
  for (int tries = 0; tries < 8 && p < n; tries++) {
    size_t l0 = p;
    const char* a = (const char*)memchr(text + l0, '\n', n - l0); if (!a) return n;
    size_t l1 = (size_t)(a - text) + 1;
    const char* b = l1 < n ? (const char*)memchr(text + l1, '\n', n - l1) : nullptr; if (!b) return n;
    size_t l2 = (size_t)(b - text) + 1;
    if (text[l0] == '@' && l2 < n && text[l2] == '+') return l0;
    p = l1;
  }
  return n;
}

void collect_seeds(shq_unitigs_s* u, const char* text, size_t n) {
  const int K = u->opts.K;
  unsigned hw = std::thread::hardware_concurrency();
  int T = (int)std::min<unsigned>(hw ? hw : 1, 8);
  if (n < ((size_t)1 << 20) || T <= 1) { collect_seeds_range(K, text, 0, n, &u->seeds, u->master_rule); return; }
  std::vector<size_t> starts(T + 1, n);
  starts[0] = 0;
  for (int i = 1; i < T; i++) starts[i] = find_record_start(text, n / T * i, n);
  for (int i = 1; i <= T; i++) if (starts[i] < starts[i - 1]) starts[i] = starts[i - 1];
  std::vector<std::vector<char>> parts(T);
  std::vector<std::thread> th;
  for (int i = 0; i < T; i++) th.emplace_back(collect_seeds_range, K, text, starts[i], starts[i + 1], &parts[i], u->master_rule);
  for (auto& t : th) t.join();
  for (auto& v : parts) u->seeds.insert(u->seeds.end(), v.begin(), v.end());
}

template <class T> void fill_u32(DevBuf& b, size_t n, uint32_t v, cudaStream_t st) {
  b.reserve(n * 4);
  fill_u32_kernel<<<cdiv(n, 256), 256, 0, st>>>(b.as<uint32_t>(), v, n);
}

struct DbgTimer {
  bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0;
  DbgTimer(cudaStream_t s) : on(getenv("SHQ_DEBUG_TIMING") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

void unitigs_build_impl(shq_unitigs_s* u) {
  shq_handle_s* h = u->h;
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  const int K = u->opts.K;
  ImageView im = image_view(h);
  NodeTable T = node_table(u);
  const size_t O = (size_t)2 * u->cap;      // oriented ids
  uint32_t n_nodes = read_scalar(h, u->nn.as<uint32_t>());
  check_err_flag(h);
  DbgTimer tm(st);
  // U3: links + false-positive closure
  u->outdeg.reserve(O); u->smask.reserve(O); u->succ.reserve(O * 4);
  SHQ_CUDA(cudaMemsetAsync(u->outdeg.p, 0, O, st)); SHQ_CUDA(cudaMemsetAsync(u->smask.p, 0, O, st));
  LinkArrays LA{u->outdeg.as<uint8_t>(), u->succ.as<uint32_t>(), u->smask.as<uint8_t>()};
  uint32_t begin = 0;
  VictimList VL{u->v_node.as<uint32_t>(), u->v_slot.as<uint64_t>(), u->v_n.as<uint32_t>(), u->victim_cap};
  for (int round = 0; round < 64 && begin < n_nodes; round++) {
    uint32_t cnt = n_nodes - begin;
    link_nodes_kernel<<<cdiv((size_t)cnt * 2, 128), 128, 0, st>>>(T, im, LA, VL, K, h->hb, (uint32_t)u->opts.abundance_min, begin, n_nodes); L++;
    u->n_queries += (uint64_t)cnt * 8;
    begin = n_nodes;
    n_nodes = read_scalar(h, u->nn.as<uint32_t>());
    check_err_flag(h);
    if (n_nodes > u->max_nodes) SHQ_FAIL(SHQ_ERR_CAPACITY, "unitig node table overflow (%u nodes)", n_nodes);
  }
  if (begin < n_nodes) SHQ_FAIL(SHQ_ERR_STATE, "false-positive closure did not converge in 64 rounds (%u nodes left)", n_nodes - begin);
  if (n_nodes == 0) { u->built = true; return; }
  const unsigned g2 = cdiv((size_t)n_nodes * 2, 256);
  tm.lap("U3 links+closure");
  // U4: chains
  fill_u32<uint32_t>(u->next, O, kNil, st); fill_u32<uint32_t>(u->prev, O, kNil, st); L += 2;
  simple_links_kernel<<<g2, 256, 0, st>>>(T, LA, n_nodes, u->next.as<uint32_t>(), u->prev.as<uint32_t>()); L++;
  u->jd.reserve(O * 8); u->misc.reserve(256);
  uint32_t* d_flag = u->misc.as<uint32_t>();
  auto rank = [&]() {
    rank_init_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->prev.as<uint32_t>(), u->jd.as<unsigned long long>()); L++;
    for (int r = 0; r < 34; r++) {
      SHQ_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
      rank_jump_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->jd.as<unsigned long long>(), d_flag); L++;
      if (!read_scalar(h, d_flag)) break;
    }
  };
  rank();
  fill_u32<uint32_t>(u->minid, O, kNil, st); fill_u32<uint32_t>(u->jmp, O, kNil, st); L += 2;
  SHQ_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
  cycle_flag_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->jd.as<unsigned long long>(), u->prev.as<uint32_t>(), u->minid.as<uint32_t>(),
                                        u->jmp.as<uint32_t>(), d_flag); L++;
  if (read_scalar(h, d_flag)) {
    for (int r = 0; r < 34; r++) { cycle_min_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->minid.as<uint32_t>(), u->jmp.as<uint32_t>()); L++; }
    cycle_cut_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->minid.as<uint32_t>(), u->next.as<uint32_t>(), u->prev.as<uint32_t>()); L++;
    rank();
  }
  tm.lap("U4 chains+ranking");
  // U5: unitigs
  fill_u32<uint32_t>(u->tail_of, O, kNil, st); fill_u32<uint32_t>(u->len_of, O, 0, st); fill_u32<uint32_t>(u->uid_of_head, O, kNil, st); L += 3;
  fill_u32<uint32_t>(u->unitig_of_oid, O, kNil, st); L++;
  tails_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->next.as<uint32_t>(), u->jd.as<unsigned long long>(), u->tail_of.as<uint32_t>(), u->len_of.as<uint32_t>()); L++;
  uint32_t max_unitigs = n_nodes + 1;
  u->u_head.reserve((size_t)max_unitigs * 4); u->u_seqlen.reserve((size_t)(max_unitigs + 2) * 8); u->u_klen.reserve((size_t)(max_unitigs + 2) * 8);
  uint32_t* d_nu = d_flag + 2;
  SHQ_CUDA(cudaMemsetAsync(d_nu, 0, 4, st));
  heads_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->prev.as<uint32_t>(), u->tail_of.as<uint32_t>(), u->len_of.as<uint32_t>(), K,
                                   u->uid_of_head.as<uint32_t>(), d_nu, u->u_head.as<uint32_t>(), u->u_seqlen.as<uint64_t>(),
                                   u->u_klen.as<uint64_t>(), max_unitigs); L++;
  uint32_t nu = read_scalar(h, d_nu);
  if (nu == 0) { u->built = true; return; }
  std::vector<uint64_t> seqlen(nu), klen(nu);
  SHQ_CUDA(cudaMemcpyAsync(seqlen.data(), u->u_seqlen.p, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
  SHQ_CUDA(cudaMemcpyAsync(klen.data(), u->u_klen.p, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
  exclusive_scan_inplace<SumU64>(u->u_seqlen.as<uint64_t>(), nu, h->d_scan_scratch, st, u->u_seqlen.as<uint64_t>() + nu, &L);
  exclusive_scan_inplace<SumU64>(u->u_klen.as<uint64_t>(), nu, h->d_scan_scratch, st, u->u_klen.as<uint64_t>() + nu, &L);
  uint64_t total_seq = read_scalar(h, u->u_seqlen.as<uint64_t>() + nu), total_k = read_scalar(h, u->u_klen.as<uint64_t>() + nu);
  u->seq.reserve(total_seq + 16); u->kcounts.reserve((total_k + 4) * 4);
  assemble_kernel<<<g2, 256, 0, st>>>(T, n_nodes, u->jd.as<unsigned long long>(), u->uid_of_head.as<uint32_t>(), u->u_seqlen.as<uint64_t>(),
                                      u->u_klen.as<uint64_t>(), K, u->seq.as<char>(), u->kcounts.as<uint32_t>(), u->unitig_of_oid.as<uint32_t>()); L++;
  tm.lap("U5 assemble");
  // U6: links, components, seeds
  u->links.reserve((size_t)nu * 8 * 4); u->parent.reserve((size_t)nu * 4); u->reached.reserve(nu); u->keep.reserve(nu);
  init_iota_kernel<<<cdiv(nu, 256), 256, 0, st>>>(u->parent.as<uint32_t>(), nu); L++;
  unitig_links_kernel<<<cdiv((size_t)nu * 2, 128), 128, 0, st>>>(T, im, LA, K, h->hb, (uint32_t)u->opts.abundance_min, nu, u->u_head.as<uint32_t>(),
                                                               u->tail_of.as<uint32_t>(), u->uid_of_head.as<uint32_t>(),
                                                               u->unitig_of_oid.as<uint32_t>(), u->links.as<int>(), u->parent.as<uint32_t>()); L++;
  uint32_t n_seeds = (uint32_t)u->n_seed_slots;
  std::vector<uint32_t> seed_node(n_seeds), seed_root(n_seeds);
  if (n_seeds) {
    u->seed_node.reserve((size_t)n_seeds * 4); u->seed_root.reserve((size_t)n_seeds * 4);
    seeds_kernel<<<cdiv(n_seeds, 128), 128, 0, st>>>(T, u->d_seed_slots.as<char>(), n_seeds, K, (uint32_t)u->opts.solid_abundance_min,
                                                    (uint32_t)u->opts.solid_abundance_max, u->unitig_of_oid.as<uint32_t>(), u->parent.as<uint32_t>(),
                                                    u->seed_node.as<uint32_t>(), u->seed_root.as<uint32_t>()); L++;
    SHQ_CUDA(cudaMemcpyAsync(seed_node.data(), u->seed_node.p, (size_t)n_seeds * 4, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaMemcpyAsync(seed_root.data(), u->seed_root.p, (size_t)n_seeds * 4, cudaMemcpyDeviceToHost, st));
  }
  uint32_t nv = read_scalar(h, u->v_n.as<uint32_t>());
  if (nv > u->victim_cap) SHQ_FAIL(SHQ_ERR_CAPACITY, "%u same-key k-mer records exceed the victim list capacity %u", nv, u->victim_cap);
  std::vector<uint32_t> vout((size_t)nv * 4);
  if (nv) {
    u->v_out.reserve((size_t)nv * 16);
    VictimList V{u->v_node.as<uint32_t>(), u->v_slot.as<uint64_t>(), u->v_n.as<uint32_t>(), u->victim_cap};
    victims_kernel<<<cdiv(nv, 128), 128, 0, st>>>(V, nv, u->aux2.as<uint32_t>(), u->unitig_of_oid.as<uint32_t>(), u->parent.as<uint32_t>(),
                                                 u->v_out.as<uint32_t>()); L++;
    SHQ_CUDA(cudaMemcpyAsync(vout.data(), u->v_out.p, (size_t)nv * 16, cudaMemcpyDeviceToHost, st));
  }
  u->roots.reserve((size_t)nu * 4);
  roots_kernel<<<cdiv(nu, 256), 256, 0, st>>>(u->parent.as<uint32_t>(), nu, u->roots.as<uint32_t>()); L++;
  std::vector<uint32_t> roots(nu);
  SHQ_CUDA(cudaMemcpyAsync(roots.data(), u->roots.p, (size_t)nu * 4, cudaMemcpyDeviceToHost, st));
  SHQ_CUDA(cudaGetLastError());
  tm.lap("U6 links+seeds");
  // D2H
  std::vector<char> seq(total_seq); std::vector<uint32_t> kc(total_k); std::vector<int> links((size_t)nu * 8); std::vector<uint8_t> keep(nu, 0);
  SHQ_CUDA(cudaMemcpyAsync(seq.data(), u->seq.p, total_seq, cudaMemcpyDeviceToHost, st));
  SHQ_CUDA(cudaMemcpyAsync(kc.data(), u->kcounts.p, total_k * 4, cudaMemcpyDeviceToHost, st));
  SHQ_CUDA(cudaMemcpyAsync(links.data(), u->links.p, (size_t)nu * 32, cudaMemcpyDeviceToHost, st));
  SHQ_CUDA(cudaStreamSynchronize(st));
  h->stats.d2h_bytes += total_seq + total_k * 4 + (size_t)nu * 36 + (size_t)n_seeds * 8 + (size_t)nv * 16;
  check_err_flag(h);
  auto rc_dbg = [](const std::string& s0) { std::string r(s0.rbegin(), s0.rend()); for (auto& c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; return r; };
  // The seed rule at t=1 (contig_assembly.cpp:1874-1880): seeds are taken in read order; a seed is skipped when its
  // KEY was traveled before.  All of a component's keys become traveled when its first seed is accepted, so a
  // seed is blocked iff another k-mer string sharing its key lives in a component explored earlier.
  {
    // same-key groups: the entry's first visitor pn plus every victim recorded against it
    std::unordered_map<uint32_t, std::vector<std::pair<uint32_t, uint32_t>>> groups;   // pn -> (node, root) members
    for (uint32_t i = 0; i < nv; i++) {
      uint32_t vn = vout[4 * i], vr = vout[4 * i + 1], pn = vout[4 * i + 2], pr = vout[4 * i + 3];
      if (pn == kNil) continue;
      auto& g = groups[pn];
      if (g.empty()) g.push_back({pn, pr});
      g.push_back({vn, vr});
    }
    std::unordered_map<uint32_t, std::vector<uint32_t>> partners;   // node slot -> roots of the other same-key nodes
    for (auto& kv : groups)
      for (auto& a : kv.second)
        for (auto& b : kv.second)
          if (a.first != b.first && b.second != kNil) partners[a.first].push_back(b.second);
    std::vector<uint32_t> t_explored(nu, 0xFFFFFFFFu);
    std::vector<uint64_t> has_partner(((size_t)u->cap + 63) / 64, 0);
    for (auto& kv : partners) has_partner[kv.first >> 6] |= 1ULL << (kv.first & 63);
    for (uint32_t i = 0; i < n_seeds; i++) {
      uint32_t r = seed_root[i];
      if (r == kNil || r >= nu) continue;
      if (t_explored[r] != 0xFFFFFFFFu) continue;                 // component already explored: nothing to decide
      bool blocked = false;
      uint32_t sn = seed_node[i];
      if ((has_partner[sn >> 6] >> (sn & 63)) & 1) {
        auto it = partners.find(sn);
        if (it != partners.end()) for (uint32_t pr : it->second) if (pr != r && pr < nu && t_explored[pr] < i) blocked = true;
      }
      if (!blocked && t_explored[r] == 0xFFFFFFFFu) t_explored[r] = i;
    }
    for (uint32_t i = 0; i < nu; i++) keep[i] = roots[i] < nu && t_explored[roots[i]] != 0xFFFFFFFFu;
    if (const char* dbg = getenv("SHQ_DEBUG_UNITIG")) {      // diagnostics for one unitig sequence (either strand)
      std::string want(dbg);
      uint64_t so = 0;
      for (uint32_t i = 0; i < nu; i++) {
        std::string sq(seq.data() + so, seqlen[i]);
        so += seqlen[i];
        if (sq != want && rc_dbg(sq) != want) continue;
        fprintf(stderr, "[dbg] unitig old=%u root=%u t_explored=%u keep=%d nv=%u n_seeds=%u\n", i, roots[i], t_explored[roots[i]], (int)keep[i], nv, n_seeds);
        for (uint32_t j = 0; j < n_seeds; j++) if (seed_root[j] == roots[i]) {
          fprintf(stderr, "[dbg]   seed %u node %u partners:", j, seed_node[j]);
          auto it = partners.find(seed_node[j]);
          if (it != partners.end()) for (uint32_t pr : it->second) fprintf(stderr, " root %u (t=%u)", pr, pr < nu ? t_explored[pr] : 0);
          fprintf(stderr, "\n");
        }
        for (int jj = 0; jj < 8; jj++) fprintf(stderr, "[dbg]   link[%d]=%d\n", jj, links[(size_t)i * 8 + jj]);
      }
    }
  }
  tm.lap("D2H + seed rule");
  // host: canonical orientation, deterministic order, renumbering, medians, FASTA (contig_assembly.cpp:603-629)
  auto rc_str = [](const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (auto& c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    return r;
  };
  struct Rec { std::string seq; int km; bool flipped; uint32_t old; };
  // every unitig is independent here: a few host threads share the work (index ranges, no locks)
  auto parallel_for = [](size_t n, const std::function<void(size_t, size_t)>& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t T = std::min<size_t>(hw ? hw : 1, 8);
    if (n < 4096 || T <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < T; t++) th.emplace_back(fn, n * t / T, n * (t + 1) / T);
    for (auto& x : th) x.join();
  };
  std::vector<uint64_t> soff(nu + 1, 0), koff(nu + 1, 0);
  for (uint32_t i = 0; i < nu; i++) { soff[i + 1] = soff[i] + seqlen[i]; koff[i + 1] = koff[i] + klen[i]; }
  std::vector<Rec> by_old(nu);
  parallel_for(nu, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      if (!keep[i]) continue;
      std::string s(seq.data() + soff[i], seqlen[i]);
      std::string r = rc_str(s);
      bool flip = r < s;
      uint32_t* c0 = kc.data() + koff[i]; size_t cn = klen[i];       // median in place (counts are not needed afterwards)
      size_t m = cn / 2;
      std::nth_element(c0, c0 + m, c0 + cn);
      double med = c0[m];
      if (cn % 2 == 0) { uint32_t lo2 = *std::max_element(c0, c0 + m); med = (lo2 + c0[m]) / 2.0; }
      by_old[i] = Rec{flip ? std::move(r) : std::move(s), (int)med, flip, (uint32_t)i};
    }
  });
  std::vector<Rec> recs;
  std::vector<int> newid(nu, -1);
  for (uint32_t i = 0; i < nu; i++) if (keep[i]) recs.push_back(std::move(by_old[i]));
  std::vector<uint32_t> order(recs.size());
  for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
  // the first 8 bases as a big-endian word decide almost every comparison
  std::vector<uint64_t> prefix(recs.size());
  for (size_t i = 0; i < recs.size(); i++) {
    uint64_t v = 0; const std::string& q = recs[i].seq;
    for (size_t j = 0; j < 8; j++) v = (v << 8) | (j < q.size() ? (unsigned char)q[j] : 0);
    prefix[i] = v;
  }
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    if (prefix[x] != prefix[y]) return prefix[x] < prefix[y];
    return recs[x].seq < recs[y].seq;
  });
  std::vector<bool> flipped_old(nu, false);
  for (uint32_t r = 0; r < order.size(); r++) { newid[recs[order[r]].old] = (int)r; flipped_old[recs[order[r]].old] = recs[order[r]].flipped; }
  const size_t nrec = order.size();
  u->seqs.assign(nrec, std::string()); u->km.assign(nrec, 0); u->after.assign(nrec, {}); u->before.assign(nrec, {}); u->fasta.clear();
  std::vector<std::string> lines(nrec);
  parallel_for(nrec, [&](size_t lo, size_t hi) {
    for (size_t r = lo; r < hi; r++) {
      Rec& R = recs[order[r]];
      std::vector<int> aft, bef;
      for (int side = 0; side < 2; side++) {
        for (int j = 0; j < 4; j++) {
          int ref = links[(size_t)R.old * 8 + side * 4 + j];
          if (!ref) continue;
          uint32_t o = (uint32_t)(abs(ref) - 1);
          if (newid[o] < 0) continue;
          bool plus = ref > 0;
          if (flipped_old[o]) plus = !plus;
          int v = plus ? newid[o] + 1 : -(newid[o] + 1);
          // a flipped unitig swaps its own sides
          if ((side == 0) != R.flipped) aft.push_back(v); else bef.push_back(v);
        }
      }
      std::string& f = lines[r];
      f.reserve(R.seq.size() + 96);
      f += ">" + std::to_string(r) + " LN:i:" + std::to_string(R.seq.size()) + " KC:i:" +
           std::to_string((long long)R.km * (long long)(R.seq.size() - K + 1)) + " km:f:" + std::to_string(R.km);
      for (int t : aft) f += t > 0 ? " L:+:" + std::to_string(t - 1) + ":+" : " L:+:" + std::to_string(-t - 1) + ":-";
      for (int t : bef) f += t > 0 ? " L:-:" + std::to_string(t - 1) + ":+" : " L:-:" + std::to_string(-t - 1) + ":-";
      f += "\n"; f += R.seq; f += "\n";
      u->km[r] = R.km; u->after[r] = std::move(aft); u->before[r] = std::move(bef); u->seqs[r] = std::move(R.seq);
    }
  });
  {
    size_t tot = 0;
    for (auto& l : lines) tot += l.size();
    u->fasta.reserve(tot);
    for (auto& l : lines) u->fasta += l;
  }
  tm.lap("host canonicalise+FASTA");
  u->built = true;
}

}  // namespace

extern "C" {

static thread_local uint64_t g_unitig_global_ndistinct = 0;
int shq_unitigs_begin(shq_t h, const shq_unitig_opts* opts, shq_unitigs_t* out);
// the unitig build over a SHARDED filter whose shards' images were opened with shq_mg_open_image_peers
int shq_unitigs_begin_sharded(shq_t h, const shq_unitig_opts* opts, uint64_t global_ndistinct, shq_unitigs_t* out) {
  g_unitig_global_ndistinct = global_ndistinct;
  return shq_unitigs_begin(h, opts, out);
}
int shq_unitigs_begin(shq_t h, const shq_unitig_opts* opts, shq_unitigs_t* out) {
  if (!h || !opts || !out) return SHQ_ERR_ARG;
  *out = nullptr;
  if (opts->K < 2 || opts->K > 64 || opts->K > kTilePos - 32 || opts->abundance_min < 1 || opts->solid_abundance_min < opts->abundance_min) {
    h->last_error = "unitigs need 2 <= K <= 64, -s >= 1 and -x >= -s";
    return SHQ_ERR_ARG;
  }
  // the device buffers of the previous build on this handle are kept (shq_unitigs_free parks the object in the
  // handle), so a repeated build allocates nothing
  shq_unitigs_s* u = h->unitig_cache ? (shq_unitigs_s*)h->unitig_cache : new shq_unitigs_s();
  h->unitig_cache = nullptr;
  u->h = h; u->opts = *opts; u->global_ndistinct = g_unitig_global_ndistinct; g_unitig_global_ndistinct = 0;
  u->built = false; u->master_rule = false; u->n_seed_slots = 0; u->st_fill = 0; u->st_ends.clear(); u->n_queries = 0;
  u->seqs.clear(); u->km.clear(); u->after.clear(); u->before.clear(); u->fasta.clear(); u->seeds.clear();
  struct { shq_unitigs_s* u; } args{u};
  int rc = guard(h, [](shq_handle_s* h, void* ap) {
    shq_unitigs_s* u = ((decltype(args)*)ap)->u;
    flush_stage(h);
    const bool sharded = h->nshards != 1;
    if (sharded && h->img_peer.empty()) SHQ_FAIL(SHQ_ERR_STATE, "unitigs on a sharded handle need the shards' images (shq_mg_open_image_peers)");
    if (!sharded && !h->image_only) materialise(h);
    uint64_t expect = sharded ? u->global_ndistinct : (h->image_only ? h->ndistinct : (uint64_t)h->n_state);
    if (expect < 1024) expect = 1024;
    uint64_t want = expect * 3 + 4096;
    uint32_t cap = 1024;
    while (cap < want && cap < (1u << 30)) cap <<= 1;
    u->cap = cap; u->max_nodes = (uint32_t)(cap * 0.7);
    cudaStream_t st = h->stream;
    u->lo.reserve((size_t)cap * 8); u->hi.reserve((size_t)cap * 8); u->fh.reserve((size_t)cap * 8); u->rh.reserve((size_t)cap * 8);
    u->state.reserve((size_t)cap * 4); u->count.reserve((size_t)cap * 4); u->list.reserve((size_t)u->max_nodes * 4 + 16); u->nn.reserve(64);
    SHQ_CUDA(cudaMemsetAsync(u->state.p, 0, (size_t)cap * 4, st));
    SHQ_CUDA(cudaMemsetAsync(u->nn.p, 0, 64, st));
    size_t nslots = (size_t)h->nblocks * 64;
    u->aux.reserve(nslots * 8); u->aux2.reserve(nslots * 4);
    SHQ_CUDA(cudaMemsetAsync(u->aux.p, 0, nslots * 8, st));
    SHQ_CUDA(cudaMemsetAsync(u->aux2.p, 0xFF, nslots * 4, st));
    u->victim_cap = (uint32_t)std::min<uint64_t>(u->max_nodes, (uint64_t)1 << 26);
    u->v_node.reserve((size_t)u->victim_cap * 4); u->v_slot.reserve((size_t)u->victim_cap * 8); u->v_n.reserve(64);
    SHQ_CUDA(cudaMemsetAsync(u->v_n.p, 0, 64, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
  }, &args);
  if (rc) { delete u; return rc; }
  *out = u;
  return SHQ_OK;
}

}  // extern "C"

namespace {
// Seeds + node enumeration for read text resident in device memory (16-byte aligned, whole records).
// host_text: the same bytes on the host, needed only when the text is irregular (a quality line that parses as a
// header derails the reference's cursor, contig_assembly.cpp:1868-1871).  Returns false, with nothing committed, if
// the text is irregular and host_text is null.
bool unitigs_process_text(shq_unitigs_s* u, const char* d_text, size_t n, const char* host_text) {
  shq_handle_s* h = u->h;
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  uint32_t n_lines = index_lines(h, d_text, n);
  if (n_lines & 3) SHQ_FAIL(SHQ_ERR_MALFORMED, "read text holds %u lines, not a multiple of 4", n_lines);
  const uint32_t n_reads = n_lines / 4;
  const int K = u->opts.K;
  {   // seed candidates: device kernel for regular text, the literal host cursor otherwise
    size_t need = (u->n_seed_slots + n_reads + 1) * (size_t)K;
    if (need > u->d_seed_slots.cap) {
      shq::DevBuf nb; nb.reserve(std::max(need * 2, (size_t)1 << 24));
      if (u->n_seed_slots) SHQ_CUDA(cudaMemcpyAsync(nb.p, u->d_seed_slots.p, u->n_seed_slots * (size_t)K, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaStreamSynchronize(st));
      u->d_seed_slots.release(); u->d_seed_slots = nb;
    }
    u->misc.reserve(256);
    int* d_irr = u->misc.as<int>() + 8;
    SHQ_CUDA(cudaMemsetAsync(d_irr, 0, 4, st));
    if (n_reads) {
      seed_extract_kernel<<<cdiv(n_reads, 128), 128, 0, st>>>(d_text, h->d_nl_pos.as<uint32_t>(), n_reads, K,
                                                            u->d_seed_slots.as<char>() + u->n_seed_slots * (size_t)K, d_irr, u->master_rule ? 1 : 0); L++;
    }
    if (read_scalar(h, d_irr)) {
      if (!host_text) return false;
      u->seeds.clear();
      collect_seeds(u, host_text, n);
      size_t m = u->seeds.size() / K;
      if (m > n_reads) SHQ_FAIL(SHQ_ERR_MALFORMED, "irregular chunk yields more seeds than records");
      if (m) SHQ_CUDA(cudaMemcpyAsync(u->d_seed_slots.as<char>() + u->n_seed_slots * (size_t)K, u->seeds.data(), m * (size_t)K, cudaMemcpyHostToDevice, st));
      SHQ_CUDA(cudaStreamSynchronize(st));
      u->n_seed_slots += m;
    } else {
      u->n_seed_slots += n_reads;
    }
  }
  EnumParams ep{d_text, h->d_nl_pos.as<uint32_t>(), n_reads, u->opts.K, h->hb, (uint32_t)u->opts.abundance_min};
  ImageView im = image_view(h);
  size_t smem = (size_t)kEnumWarps * ((kTilePos + 1) * 16 + 264 + 520);
  SHQ_CUDA(cudaFuncSetAttribute(enum_nodes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VictimList V{u->v_node.as<uint32_t>(), u->v_slot.as<uint64_t>(), u->v_n.as<uint32_t>(), u->victim_cap};
  enum_nodes_kernel<<<4 * kNumSMs, kEnumWarps * 32, smem, st>>>(ep, im, u->aux.as<uint64_t>(), u->aux2.as<uint32_t>(), V, node_table(u)); L++;
  SHQ_CUDA(cudaGetLastError());
  SHQ_CUDA(cudaStreamSynchronize(st));
  check_err_flag(h);
  return true;
}

// Everything staged so far (several reference chunks, concatenated) in one go; an irregular batch is replayed chunk
// by chunk with the host cursor, because the reference's cursor starts afresh in every chunk.
void unitigs_flush_stage(shq_unitigs_s* u) {
  if (u->st_fill == 0) return;
  shq_handle_s* h = u->h;
  cudaStream_t st = h->stream;
  DbgTimer tm(st);
  const size_t n = u->st_fill;
  std::vector<size_t> ends; ends.swap(u->st_ends);
  u->st_fill = 0;
  if (!unitigs_process_text(u, u->d_stage.as<char>(), n, nullptr)) {
    std::vector<char> host(n);
    SHQ_CUDA(cudaMemcpyAsync(host.data(), u->d_stage.p, n, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    size_t b = 0;
    for (size_t e : ends) {
      u->d_text.reserve(e - b + 64);
      SHQ_CUDA(cudaMemcpyAsync(u->d_text.p, u->d_stage.as<char>() + b, e - b, cudaMemcpyDeviceToDevice, st));
      unitigs_process_text(u, u->d_text.as<char>(), e - b, host.data() + b);
      b = e;
    }
  }
  tm.lap("add_reads: index+seeds+enum of a staged batch");
}
}  // namespace

extern "C" {

int shq_unitigs_add_reads(shq_unitigs_t u, const char* text, size_t n) {
  if (!u || (!text && n)) return SHQ_ERR_ARG;
  if (n == 0) return SHQ_OK;
  struct { shq_unitigs_s* u; const char* t; size_t n; } args{u, text, n};
  return guard(u->h, [](shq_handle_s* h, void* ap) {
    auto& a = *(decltype(args)*)ap;
    shq_unitigs_s* u = a.u;
    cudaStream_t st = h->stream;
    if (a.n >= ((size_t)1 << 31)) SHQ_FAIL(SHQ_ERR_ARG, "chunk too large");
    if (a.t[a.n - 1] != '\n') SHQ_FAIL(SHQ_ERR_MALFORMED, "chunk has no trailing newline");
    const size_t cap = std::max(u->st_cap_target, a.n);
    if (u->st_fill && u->st_fill + a.n > cap) unitigs_flush_stage(u);
    u->d_stage.reserve(cap + 64);
    // chunks are whole records, so their concatenation is valid FASTQ text; the copy is complete when the call
    // returns (the caller may reuse its buffer)
    SHQ_CUDA(cudaMemcpyAsync(u->d_stage.as<char>() + u->st_fill, a.t, a.n, cudaMemcpyHostToDevice, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    h->stats.h2d_bytes += a.n;
    u->st_fill += a.n; u->st_ends.push_back(u->st_fill);
  }, &args);
}

int shq_unitigs_set_seed_rule(shq_unitigs_t u, int master_rule) {
  if (!u) return SHQ_ERR_ARG;
  if (u->st_fill || u->n_seed_slots) return SHQ_ERR_STATE;      // before the first reads
  u->master_rule = master_rule != 0;
  return SHQ_OK;
}

int shq_unitigs_build(shq_unitigs_t u) {
  if (!u) return SHQ_ERR_ARG;
  struct { shq_unitigs_s* u; } args{u};
  return guard(u->h, [](shq_handle_s*, void* ap) {
    shq_unitigs_s* u = ((decltype(args)*)ap)->u;
    unitigs_flush_stage(u);
    unitigs_build_impl(u);
  }, &args);
}
int shq_unitigs_count(shq_unitigs_t u, size_t* n_unitigs, size_t* total_bases) {
  if (!u || !u->built) return SHQ_ERR_STATE;
  if (n_unitigs) *n_unitigs = u->seqs.size();
  if (total_bases) { size_t t = 0; for (auto& s : u->seqs) t += s.size(); *total_bases = t; }
  return SHQ_OK;
}
int shq_unitigs_fasta(shq_unitigs_t u, char* dst, size_t cap, size_t* written) {
  if (!u || !u->built || !written) return SHQ_ERR_STATE;
  *written = u->fasta.size();
  if (cap == 0) return SHQ_OK;
  if (cap < u->fasta.size()) return SHQ_ERR_ARG;
  memcpy(dst, u->fasta.data(), u->fasta.size());
  return SHQ_OK;
}
int shq_unitigs_write_fasta(shq_unitigs_t u, const char* path) {
  if (!u || !u->built || !path) return SHQ_ERR_STATE;
  FILE* f = fopen(path, "w");
  if (!f) return SHQ_ERR_IO;
  size_t w = fwrite(u->fasta.data(), 1, u->fasta.size(), f);
  fclose(f);
  return w == u->fasta.size() ? SHQ_OK : SHQ_ERR_IO;
}
void shq_unitigs_free(shq_unitigs_t u) {
  if (!u) return;
  if (u->h && !u->h->unitig_cache) {       // park the object (and its device buffers) for the next build on this handle
    u->seqs.clear(); u->seqs.shrink_to_fit(); u->fasta.clear(); u->fasta.shrink_to_fit(); u->km.clear(); u->after.clear(); u->before.clear();
    u->h->unitig_cache = u;
    return;
  }
  shq_unitigs_release(u);
}
void shq_unitigs_release(shq_unitigs_s* u) {
  if (!u) return;
  cudaSetDevice(u->h->device);
  shq::DevBuf* bufs[] = {&u->lo, &u->hi, &u->state, &u->count, &u->fh, &u->rh, &u->list, &u->nn, &u->aux, &u->outdeg, &u->succ, &u->smask, &u->next,
                         &u->prev, &u->jd, &u->minid, &u->jmp, &u->tail_of, &u->len_of, &u->uid_of_head, &u->unitig_of_oid, &u->u_head, &u->u_seqlen,
                         &u->u_klen, &u->seq, &u->kcounts, &u->links, &u->parent, &u->reached, &u->keep, &u->d_seeds, &u->d_text, &u->misc, &u->aux2, &u->v_node, &u->v_slot, &u->v_n, &u->v_out, &u->seed_node, &u->seed_root, &u->roots, &u->d_seed_slots, &u->d_stage};
  for (auto* b : bufs) b->release();
  delete u;
}

}  // extern "C"
