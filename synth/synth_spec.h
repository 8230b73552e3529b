/* synth_spec.h -- the synthetic read set of SURVEY.md 8(d), as a counter-based function so that the CUDA
 * generator (synth_gpu.cu, fills HBM directly at benchmark sizes) and the C generator (synth_cpu.c, feeds the
 * CPU oracle / reference arm and produces the committed digests) emit the SAME bytes for the same parameters.
 *
 * Bench / test tooling, not part of the drop-in ABI (the reference ships no read generator).
 *
 * Genome: iid uniform ACGT of length G.  Read r: L bases at a uniform position, reverse-complemented with
 * p = 1/2, every base substituted (uniformly by one of the other three) with p = err.  No N, no indels.
 * Record: "@r<id, zero padded>\n<seq>\n+\n<L x 'I'>\n".                                                       */
#ifndef SYNTH_SPEC_H
#define SYNTH_SPEC_H
#include <stdint.h>

#ifdef __CUDACC__
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD static inline
#endif

typedef struct {
  uint64_t genome_len;   /* G */
  uint64_t n_reads;      /* reads of the WHOLE set (fixes the id width) */
  uint64_t s0;           /* synth_mix(seed) */
  uint32_t read_len;     /* L */
  uint32_t err_thr;      /* substitution iff the high 32 bits of the draw < err_thr (= err * 2^32) */
  uint32_t id_width;     /* decimal digits of n_reads - 1 ... of n_reads (see synth_make) */
  uint32_t rec_bytes;    /* 2 + id_width + 1 + L + 1 + 2 + L + 1 */
} synth_params;

#define SYNTH_TAG_G 0x1000000000000000ULL
#define SYNTH_TAG_R 0x2000000000000000ULL
#define SYNTH_TAG_E 0x3000000000000000ULL

SYNTH_HD uint64_t synth_mix(uint64_t x) {   /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
/* 32 genome bases per word */
SYNTH_HD uint64_t synth_genome_word(const synth_params* p, uint64_t w) { return synth_mix(p->s0 ^ SYNTH_TAG_G ^ w); }
SYNTH_HD uint32_t synth_genome_code(const synth_params* p, uint64_t i) { return (uint32_t)(synth_genome_word(p, i >> 5) >> (2 * (i & 31))) & 3u; }
/* read r: start position and strand */
SYNTH_HD void synth_read_place(const synth_params* p, uint64_t r, uint64_t* pos, uint32_t* flip) {
  uint64_t h = synth_mix(p->s0 ^ SYNTH_TAG_R ^ r);
  *flip = (uint32_t)(h & 1);
  *pos = (h >> 1) % (p->genome_len - p->read_len + 1);
}
/* base j of read r given the genome code at its source position (already complemented when flipped) */
SYNTH_HD uint32_t synth_apply_error(const synth_params* p, uint64_t r, uint32_t j, uint32_t code) {
  uint64_t e = synth_mix(p->s0 ^ SYNTH_TAG_E ^ (r * p->read_len + j));
  if ((uint32_t)(e >> 32) < p->err_thr) code = (code + 1 + (uint32_t)(((e & 0xFFFF) * 3) >> 16)) & 3u;
  return code;
}
#endif
