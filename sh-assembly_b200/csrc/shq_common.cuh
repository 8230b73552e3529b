// shq_common.cuh -- shared device/host helpers for libshq (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/shq.h"

namespace shq {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

struct CudaError {
  cudaError_t err; const char* what; const char* file; int line;
};

#define SHQ_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) throw ::shq::CudaError{_e, #call, __FILE__, __LINE__}; \
  } while (0)

struct StatusError { int status; std::string msg; };
#define SHQ_FAIL(st, ...)                                   \
  do {                                                      \
    char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__);     \
    throw ::shq::StatusError{(st), std::string(_b)};        \
  } while (0)

// Grow-only device buffer.
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) SHQ_CUDA(cudaFree(p));
    p = nullptr; cap = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    SHQ_CUDA(cudaMalloc(&p, want));
    cap = want;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

// ---------------------------------------------------------------- ntHash ----
// base/nthash.hpp:24-28 seeds; msTab[c][j] == rol(seed_c, j) (:30-118).
__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t v, unsigned s) {
  s &= 63; return s ? (v << s) | (v >> (64 - s)) : v;
}
__host__ __device__ __forceinline__ uint64_t rotr64(uint64_t v, unsigned s) {
  s &= 63; return s ? (v >> s) | (v << (64 - s)) : v;
}
#define SHQ_SEED_A 0x3c8bfbb395c60474ULL
#define SHQ_SEED_C 0x3193c18562a02b4cULL
#define SHQ_SEED_G 0x20323ed082572324ULL
#define SHQ_SEED_T 0x295549f54be24456ULL

// forward-strand seed of a text byte: msTab rows 'A','C','G','T' (either case) and the raw codes
// 4,7,3,1 that the table also populates (nthash.hpp:85-87); every other byte hashes as 0.
__host__ __device__ __forceinline__ uint64_t nt_seed_fwd(unsigned char c) {
  switch (c) {
    case 'A': case 'a': case 4: return SHQ_SEED_A;
    case 'C': case 'c': case 7: return SHQ_SEED_C;
    case 'G': case 'g': case 3: return SHQ_SEED_G;
    case 'T': case 't': case 1: return SHQ_SEED_T;
    default: return 0;
  }
}
// reverse-strand seed: msTab[c & 0x07] (cpOff, nthash.hpp:15): 'A'&7=1->T, 'C'&7=3->G, 'G'&7=7->C,
// 'T'&7=4->A.  NOTE the index is c&7 for ANY byte c, so e.g. 'Q'(0x51)&7 = 1 also maps to T.
__host__ __device__ __forceinline__ uint64_t nt_seed_rev(unsigned char c) {
  switch (c & 0x07) {
    case 1: return SHQ_SEED_T;
    case 3: return SHQ_SEED_G;
    case 4: return SHQ_SEED_A;
    case 7: return SHQ_SEED_C;
    default: return 0;
  }
}

// encode_counter length (cqf/gqf.c:1225-1255) for 8-bit slots: [rem] [0 if first digit > rem]
// [digits of count-1 in base 128, high ones with MSB set].
__host__ __device__ __forceinline__ int enc_len(uint32_t rem, uint64_t count) {
  if (count <= 1) return 1;
  uint64_t c = count - 1;
  int nd = 1; uint64_t t = c >> 7;
  while (t) { nd++; t >>= 7; }
  uint32_t first = nd > 1 ? (uint32_t)(((c >> (7 * (nd - 1))) & 0x7F) | 0x80) : (uint32_t)c;
  return 1 + (first > rem ? 1 : 0) + nd;
}
// writes the encoding into out[0..len)
__host__ __device__ __forceinline__ int enc_bytes(uint32_t rem, uint64_t count, uint8_t* out) {
  int n = 0;
  out[n++] = (uint8_t)rem;
  if (count <= 1) return n;
  uint64_t c = count - 1;
  int nd = 1; uint64_t t = c >> 7;
  while (t) { nd++; t >>= 7; }
  uint32_t first = nd > 1 ? (uint32_t)(((c >> (7 * (nd - 1))) & 0x7F) | 0x80) : (uint32_t)c;
  if (first > rem) out[n++] = 0;
  for (int d = nd - 1; d >= 1; d--) out[n++] = (uint8_t)(((c >> (7 * d)) & 0x7F) | 0x80);
  out[n++] = (uint8_t)(c & 0x7F);
  return n;
}

}  // namespace shq
