// synth_gpu.cu -- CUDA twin of synth_cpu.c: fills a device buffer with the synthetic FASTQ records of reads
// [first, first + n) (see synth_spec.h).  One thread per 16 output bytes, one 16-byte store per thread.
// Bench tooling (the BASELINE configs need up to 21 GB of text resident in HBM); not part of the drop-in ABI.
#include <cuda_runtime.h>
#include "synth_spec.h"

__global__ void __launch_bounds__(256) synth_fill_kernel(synth_params p, uint64_t first, uint64_t n_bytes, uint4* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t o0 = t * 16;
  if (o0 >= n_bytes) return;
  const uint32_t rec = p.rec_bytes, L = p.read_len, w = p.id_width;
  uint64_t i = o0 / rec;                  // local record
  uint32_t c = (uint32_t)(o0 - i * rec);  // offset inside it
  const uint32_t seq0 = 2 + w + 1, plus0 = seq0 + L + 1, q0 = plus0 + 2;
  uint64_t pos = 0; uint32_t flip = 0; bool placed = false;
  uint64_t gw = 0, gw_idx = ~0ULL;
  unsigned char b[16];
#pragma unroll 1
  for (int k = 0; k < 16; k++) {
    unsigned char ch;
    const uint64_t r = first + i;
    if (o0 + k >= n_bytes) ch = 0;
    else if (c == 0) ch = '@';
    else if (c == 1) ch = 'r';
    else if (c < 2 + w) {
      uint64_t v = r;
      for (uint32_t d = c - 2; d + 1 < w; d++) v /= 10;
      ch = (unsigned char)('0' + v % 10);
    } else if (c == 2 + w || c == plus0 - 1 || c == plus0 + 1 || c == rec - 1) ch = '\n';
    else if (c == plus0) ch = '+';
    else if (c >= q0) ch = 'I';
    else {
      const uint32_t j = c - seq0;
      if (!placed) { synth_read_place(&p, r, &pos, &flip); placed = true; }
      const uint64_t src = flip ? pos + L - 1 - j : pos + j;
      if ((src >> 5) != gw_idx) { gw_idx = src >> 5; gw = synth_genome_word(&p, gw_idx); }
      uint32_t code = (uint32_t)(gw >> (2 * (src & 31))) & 3u;
      if (flip) code = 3 - code;
      code = synth_apply_error(&p, r, j, code);
      ch = (unsigned char)("ACGT"[code]);
    }
    b[k] = ch;
    if (++c == rec) { c = 0; i++; placed = false; }
  }
  uint4 v;
  v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
  v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
  v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
  v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
  out[t] = v;
}

extern "C" {
// d_out: 16-byte aligned device buffer of at least round_up(n * rec_bytes, 16) bytes.  Asynchronous on `stream`.
int synth_gpu_fill(const synth_params* p, uint64_t first, uint64_t n, void* d_out, void* stream) {
  const uint64_t n_bytes = n * (uint64_t)p->rec_bytes;
  if (n_bytes == 0) return 0;
  if (reinterpret_cast<uintptr_t>(d_out) & 15) return 1;
  const uint64_t threads = (n_bytes + 15) / 16;
  const uint64_t blocks = (threads + 255) / 256;
  if (blocks > 0x7FFFFFFFULL) return 1;
  synth_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*p, first, n_bytes, (uint4*)d_out);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
}
