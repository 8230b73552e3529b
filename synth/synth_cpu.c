/* synth_cpu.c -- C twin of synth_gpu.cu (same bytes, see synth_spec.h).  gcc -O2 -shared. */
#include <stdlib.h>
#include <string.h>
#include "synth_spec.h"

void synth_make(uint64_t genome_len, uint64_t n_reads, uint32_t read_len, double err, uint64_t seed, synth_params* p) {
  p->genome_len = genome_len; p->n_reads = n_reads; p->read_len = read_len; p->s0 = synth_mix(seed);
  double t = err * 4294967296.0;
  p->err_thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  uint32_t w = 1; uint64_t v = n_reads; while (v >= 10) { v /= 10; w++; }
  p->id_width = w;
  p->rec_bytes = 2 + w + 1 + read_len + 1 + 2 + read_len + 1;
}

/* records of reads [first, first + n) into out (n * rec_bytes bytes) */
void synth_cpu_fill(const synth_params* p, uint64_t first, uint64_t n, char* out) {
  static const char B[4] = {'A', 'C', 'G', 'T'};
  const uint32_t L = p->read_len, w = p->id_width;
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t r = first + i;
    char* o = out + i * (uint64_t)p->rec_bytes;
    *o++ = '@'; *o++ = 'r';
    uint64_t v = r;
    for (int d = (int)w - 1; d >= 0; d--) { o[d] = (char)('0' + v % 10); v /= 10; }
    o += w; *o++ = '\n';
    uint64_t pos; uint32_t flip;
    synth_read_place(p, r, &pos, &flip);
    for (uint32_t j = 0; j < L; j++) {
      uint64_t src = flip ? pos + L - 1 - j : pos + j;
      uint32_t code = synth_genome_code(p, src);
      if (flip) code = 3 - code;
      o[j] = B[synth_apply_error(p, r, j, code)];
    }
    o += L; *o++ = '\n'; *o++ = '+'; *o++ = '\n';
    memset(o, 'I', L); o += L; *o++ = '\n';
  }
}

/* the genome itself (tests: ground truth for k-mer sets) */
void synth_cpu_genome(const synth_params* p, uint64_t first, uint64_t n, char* out) {
  static const char B[4] = {'A', 'C', 'G', 'T'};
  for (uint64_t i = 0; i < n; i++) out[i] = B[synth_genome_code(p, first + i)];
}
