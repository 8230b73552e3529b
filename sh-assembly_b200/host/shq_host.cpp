// shq_host.cpp -- host-side C++ mirror of the reference's readers and parameter selection,
// behind the C ABI of include/shq.h.  Compiled into libshq.so next to the CUDA code.
//
// Mirrors (behaviour, not code): seqFile_batch / fastq_read_parts / getFileReader
// (cqf/CQF_mt.h:334-412, 735-816, 933-957, cqf/reader.h:32-57) and the parameter block of
// src/CQF-deNoise.cpp:83-161 with mean_CDF2deNoise (cqf/CQF_mt.h:94-133) and true2falseKmer_DP
// (cqf/true2falseKmer_DP.cpp:12-51).
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/shq.h"

namespace {

// ---- bzip2 through dlopen: the image has libbz2.so.1 but no bzlib.h --------------------------
struct Bz2Api {
  void* lib = nullptr;
  void* (*ReadOpen)(int*, FILE*, int, int, void*, int) = nullptr;
  int (*Read)(int*, void*, void*, int) = nullptr;
  void (*ReadClose)(int*, void*) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libbz2.so.1", RTLD_NOW);
    if (!lib) lib = dlopen("libbz2.so.1.0", RTLD_NOW);
    if (!lib) return false;
    ReadOpen = (decltype(ReadOpen))dlsym(lib, "BZ2_bzReadOpen");
    Read = (decltype(Read))dlsym(lib, "BZ2_bzRead");
    ReadClose = (decltype(ReadClose))dlsym(lib, "BZ2_bzReadClose");
    return ReadOpen && Read && ReadClose;
  }
};
Bz2Api g_bz2;
constexpr int kBzStreamEnd = 4;  // BZ_STREAM_END

// One input: plain FILE*, gzFile, BZFILE* or an in-memory buffer; read(n) + eof() like cqf/reader.h.
struct Source {
  char mode = 'f';
  FILE* in = nullptr; gzFile gz = nullptr; void* bz = nullptr; int bzerror = 0;
  const char* mem = nullptr; size_t mem_len = 0, mem_pos = 0; bool mem_eof = false;
  int index = 0;
  std::vector<char> tail;  // part_buffer / part_filled

  // prefetch: one producer thread per file input cuts (and, for gzip / bzip2, decompresses) chunks ahead of the
  // consumer; the chunk SEQUENCE of an input does not depend on the other inputs, so the round-robin order the
  // consumer sees is the serial one
  struct Item { std::vector<char>* buf; size_t size; int rc; };
  std::thread worker;
  std::mutex mu; std::condition_variable cv;
  std::deque<Item> ready; std::vector<std::vector<char>*> free_bufs;
  std::vector<std::unique_ptr<std::vector<char>>> owned;
  bool stop = false, started = false;

  size_t read(char* dst, size_t n) {
    if (mode == 'm') {
      size_t r = std::min(n, mem_len - mem_pos);
      memcpy(dst, mem + mem_pos, r); mem_pos += r;
      if (r < n) mem_eof = true;
      return r;
    }
    if (mode == 'f') return fread(dst, 1, n, in);
    if (mode == 'g') { int r = gzread(gz, dst, (unsigned)n); return r < 0 ? 0 : (size_t)r; }
    if (mode == 'b') { int r = g_bz2.Read(&bzerror, bz, dst, (int)n); return r < 0 ? 0 : (size_t)r; }
    return 0;
  }
  bool eof() const {   // is_eof, CQF_mt.h:561-570
    if (mode == 'm') return mem_eof;
    if (mode == 'f') return feof(in) != 0;
    if (mode == 'g') return gzeof(gz) != 0;
    if (mode == 'b') return bzerror == kBzStreamEnd;
    return true;
  }
  void close() {
    if (mode == 'f' && in) fclose(in);
    if (mode == 'g' && gz) gzclose(gz);
    if (mode == 'b' && in) { if (bz) g_bz2.ReadClose(&bzerror, bz); fclose(in); }
    in = nullptr; gz = nullptr; bz = nullptr;
  }
};

bool skip_next_eol(const char* part, int64_t& pos, int64_t max_pos) {
  int64_t i;
  for (i = pos; i < max_pos - 2; ++i)
    if ((part[i] == '\n' || part[i] == '\r') && !(part[i + 1] == '\n' || part[i + 1] == '\r')) break;
  if (i >= max_pos - 2) return false;
  pos = i + 1;
  return true;
}

}  // namespace

struct shq_reader_s {
  bool pin = false;                    // chunk buffers are page-locked (cudaHostRegister) so that H2D copies can be asynchronous
  std::vector<void*> registered;
  void pin_buffer(std::vector<char>& v) {
    if (!pin || v.empty()) return;
    void* p = v.data();
    for (void* q : registered) if (q == p) return;
    if (cudaHostRegister(p, v.size(), cudaHostRegisterPortable) == cudaSuccess) registered.push_back(p); else { (void)cudaGetLastError(); pin = false; }
  }
  std::deque<Source*> queue;           // FIFO of open inputs (round-robin)
  std::vector<std::unique_ptr<Source>> all;
  int num_files = 0;
  uint64_t part_size = 1ULL << 23;
  std::vector<char> part_serial;       // chunk storage of the in-memory inputs (no prefetch)
  std::vector<char>* held = nullptr; Source* held_src = nullptr;   // prefetched chunk handed out by the last next()
  std::string error; std::mutex err_mu;
  static constexpr int kDepth = 3;     // chunks cut ahead per file

  // 1 = chunk produced (size in *size), 0 = this input is exhausted, <0 = error
  int read_parts(Source* s, std::vector<char>& part, size_t* size) {
    const int64_t OVERHEAD_SIZE = 65535;
    if (part.size() != part_size + OVERHEAD_SIZE) { part.resize(part_size + OVERHEAD_SIZE); std::lock_guard<std::mutex> g(err_mu); pin_buffer(part); }
    size_t part_filled = s->tail.size();
    if (part_filled) memcpy(part.data(), s->tail.data(), part_filled);
    if (s->eof()) return 0;
    size_t readed = s->read(part.data() + part_filled, part_size);
    int64_t total_filled = (int64_t)(part_filled + readed);
    if ((int64_t)part_filled >= OVERHEAD_SIZE) { std::lock_guard<std::mutex> g(err_mu); error = "Error: Wrong input file!"; return -1; }
    if (s->eof()) { *size = (size_t)total_filled; s->tail.clear(); return 1; }
    int64_t cut = 0;
    int64_t line_start[9];
    int j;
    int64_t i = total_filled - OVERHEAD_SIZE / 2;
    if (i < 0) i = 0;
    for (j = 0; j < 9; ++j) {
      if (!skip_next_eol(part.data(), i, total_filled)) break;
      line_start[j] = i;
    }
    if (j == 9) {
      const char* P = part.data();
      int k;
      for (k = 0; k < 4; ++k) {
        if (P[line_start[k]] == '@' && P[line_start[k + 2]] == '+') {
          if (P[line_start[k + 2] + 1] == '\n' || P[line_start[k + 2] + 1] == '\r') break;
          if (line_start[k + 1] - line_start[k] == line_start[k + 3] - line_start[k + 2] &&
              memcmp(P + line_start[k] + 1, P + line_start[k + 2] + 1, (size_t)(line_start[k + 3] - line_start[k + 2] - 1)) == 0)
            break;
        }
      }
      cut = (k == 4) ? 0 : line_start[k];
    }
    s->tail.assign(part.data() + cut, part.data() + total_filled);
    *size = (size_t)cut;
    return 1;
  }

  void start_prefetch(Source* s) {
    for (int i = 0; i < kDepth + 2; i++) { s->owned.emplace_back(new std::vector<char>()); s->free_bufs.push_back(s->owned.back().get()); }
    s->started = true;
    s->worker = std::thread([this, s]() {
      for (;;) {
        std::vector<char>* buf = nullptr;
        {
          std::unique_lock<std::mutex> lk(s->mu);
          s->cv.wait(lk, [&] { return s->stop || (!s->free_bufs.empty() && (int)s->ready.size() < kDepth); });
          if (s->stop) return;
          buf = s->free_bufs.back(); s->free_bufs.pop_back();
        }
        size_t size = 0;
        int rc = read_parts(s, *buf, &size);
        {
          std::lock_guard<std::mutex> lk(s->mu);
          s->ready.push_back(Source::Item{buf, size, rc});
        }
        s->cv.notify_all();
        if (rc != 1) return;             // exhausted or malformed: nothing more to cut
      }
    });
  }
  void stop_prefetch(Source* s) {
    if (!s->started) return;
    { std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
    s->cv.notify_all();
    if (s->worker.joinable()) s->worker.join();
    s->started = false;
  }
  void give_back() {                      // the chunk handed out by the previous next() may be overwritten now
    if (!held) return;
    { std::lock_guard<std::mutex> lk(held_src->mu); held_src->free_bufs.push_back(held); }
    held_src->cv.notify_all();
    held = nullptr; held_src = nullptr;
  }
};

extern "C" {

int shq_reader_open(const char* const* paths, int nfiles, char fmode, shq_reader_t* out) {
  if (!paths || nfiles < 0 || !out) return SHQ_ERR_ARG;
  if (fmode != 'f' && fmode != 'g' && fmode != 'b') return SHQ_ERR_ARG;
  if (fmode == 'b' && !g_bz2.load()) { fprintf(stderr, "shq: libbz2.so.1 not found\n"); return SHQ_ERR_IO; }
  shq_reader_s* r = new shq_reader_s();
  for (int i = 0; i < nfiles; i++) {
    std::unique_ptr<Source> s(new Source());
    s->mode = fmode; s->index = i;
    bool ok = false;                                            // getFileReader, CQF_mt.h:933-957
    if (fmode == 'f') ok = (s->in = fopen(paths[i], "rb")) != nullptr;
    else if (fmode == 'g') { ok = (s->gz = gzopen(paths[i], "rb")) != nullptr; if (ok) gzbuffer(s->gz, 1u << 26); }
    else {
      s->in = fopen(paths[i], "rb");
      if (s->in) {
        setvbuf(s->in, nullptr, _IOFBF, (size_t)1 << 26);
        s->bz = g_bz2.ReadOpen(&s->bzerror, s->in, 0, 0, nullptr, 0);
        ok = s->bz != nullptr;
        if (!ok) { fclose(s->in); s->in = nullptr; }
      }
    }
    if (ok) { r->queue.push_back(s.get()); r->all.push_back(std::move(s)); r->num_files++; }
    // files that cannot be opened are silently skipped, exactly like seqFile_batch (CQF_mt.h:350-360)
  }
  for (auto& s : r->all) r->start_prefetch(s.get());
  *out = r;
  return SHQ_OK;
}

int shq_reader_open_memory(const char* const* bufs, const size_t* lens, int nfiles, uint64_t part_size, shq_reader_t* out) {
  if (!bufs || !lens || nfiles < 0 || !out) return SHQ_ERR_ARG;
  if (part_size && part_size < 65536) return SHQ_ERR_ARG;
  shq_reader_s* r = new shq_reader_s();
  if (part_size) r->part_size = part_size;
  for (int i = 0; i < nfiles; i++) {
    std::unique_ptr<Source> s(new Source());
    s->mode = 'm'; s->mem = bufs[i]; s->mem_len = lens[i]; s->index = i;
    r->queue.push_back(s.get()); r->all.push_back(std::move(s)); r->num_files++;
  }
  *out = r;
  return SHQ_OK;
}

int shq_reader_next(shq_reader_t r, const char** chunk, size_t* n, int* file_index) {
  if (!r || !chunk || !n) return SHQ_ERR_ARG;
  *chunk = nullptr; *n = 0;
  r->give_back();
  while (r->num_files) {               // seqFile_batch::getDataChunk, CQF_mt.h:364-390
    Source* s = r->queue.front(); r->queue.pop_front();
    size_t size = 0;
    int rc;
    const char* data = nullptr;
    if (s->started) {                  // next chunk of this input, cut ahead by its producer thread
      Source::Item it;
      {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return !s->ready.empty(); });
        it = s->ready.front(); s->ready.pop_front();
      }
      s->cv.notify_all();
      rc = it.rc; size = it.size; data = it.buf->data();
      if (rc == 1) { r->held = it.buf; r->held_src = s; }
    } else {
      rc = r->read_parts(s, r->part_serial, &size);
      data = r->part_serial.data();
    }
    if (rc < 0) return SHQ_ERR_MALFORMED;
    if (rc == 1) {
      r->queue.push_back(s);
      *chunk = data; *n = size;
      if (file_index) *file_index = s->index;
      return SHQ_OK;
    }
    r->stop_prefetch(s);
    s->close();
    r->num_files--;
  }
  return SHQ_OK;
}

void shq_reader_close(shq_reader_t r) {
  if (!r) return;
  for (auto& s : r->all) { r->stop_prefetch(s.get()); s->close(); }
  for (void* p : r->registered) cudaHostUnregister(p);
  delete r;
}

int shq_reader_pin(shq_reader_t r) {
  if (!r) return SHQ_ERR_ARG;
  std::lock_guard<std::mutex> g(r->err_mu);
  r->pin = true;
  for (auto& s : r->all) for (auto& b : s->owned) r->pin_buffer(*b);
  r->pin_buffer(r->part_serial);
  return r->pin ? SHQ_OK : SHQ_ERR_CUDA;
}

int shq_build_kmer_spectrum(shq_t h, shq_reader_t r, int K) {
  if (!h || !r) return SHQ_ERR_ARG;
  // the reader's chunk buffers are page-locked, so every chunk goes to the device with an asynchronous copy on the
  // handle's copy stream straight out of the buffer the producer thread filled (no staging memcpy), overlapping the
  // compute of the previous batch; the buffer goes back to its producer once that copy has landed
  const bool pinned = shq_reader_pin(r) == SHQ_OK;
  for (;;) {
    const char* chunk = nullptr; size_t n = 0;
    if (pinned) { int rc = shq_sync_copies(h); if (rc) return rc; }     // shq_reader_next recycles the previous chunk's buffer
    int rc = shq_reader_next(r, &chunk, &n, nullptr);
    if (rc) return rc;
    if (!chunk) break;
    // an empty chunk (no record boundary found, cut == 0) is still a chunk boundary for the deNoise trigger
    // (cqf/CQF_mt.h:829-839), so it is passed on
    rc = (pinned && n) ? shq_insert_fastq_pinned(h, chunk, n, K) : shq_insert_fastq_chunk(h, chunk, n, K);
    if (rc) return rc;
  }
  return shq_finish(h);
}

// ------------------------------------------------------------------ parameters ----
// boost::math::cdf(poisson(mean), x) = Q(floor(x)+1, mean) = e^-mean * sum_{j<=floor(x)} mean^j/j!
extern "C" double shq_cdfpoi(double x, double mean);
static double cdfpoi(double x, double mean) {
  if (x < 0) return 0;
  if (mean <= 0) return 1;
  long k = (long)floor(x);
  double sum = 0, lm = log(mean);
  for (long j = 0; j <= k; j++) sum += exp(-mean + j * lm - lgamma((double)j + 1.0));
  return sum > 1 ? 1 : sum;
}
double shq_cdfpoi(double x, double mean) { return cdfpoi(x, mean); }
double shq_cdfpoi_positive(double x, double mean) {
  double c0 = cdfpoi(0, mean);
  return (cdfpoi(x, mean) - c0) / (1 - c0);
}
static int mean_CDF2deNoise(double mean, double cdf_desired) {
  int start = 0, end = (int)(mean + 1), mid;
  while (shq_cdfpoi_positive(end, mean) < cdf_desired) end *= 2;
  while (start <= end) {
    if (start == end) return start;
    if (start + 1 == end) {
      double t1 = shq_cdfpoi_positive(start, mean), t2 = shq_cdfpoi_positive(end, mean);
      if (t2 <= cdf_desired) return end;
      if (t1 <= cdf_desired) return start;
      return std::max(start - 1, 0);
    }
    mid = (start + end) / 2;
    double cdf = shq_cdfpoi_positive(mid, mean);
    if (cdf < cdf_desired) start = mid + 1;
    else if (cdf > cdf_desired) end = mid - 1;
    else return start;
  }
  return start;
}
static double true2falseKmer_DP(const char* errorFile, size_t K) {
  std::vector<double> e;
  std::ifstream fin(errorFile);
  double t;
  while (fin >> t) e.push_back(t);
  size_t L = e.size();
  if (L < K) return 0;
  std::vector<double> DP(K + 1, 0), nDP(K + 1, 0);
  double tmp = 1, trueP;
  for (size_t x = 0; x < K; x++) tmp *= (1 - e[x]);
  DP[0] = tmp;
  for (size_t x = 1; x <= K; x++) {
    tmp = e[x - 1];
    for (size_t y = x; y < K; y++) tmp *= (1 - e[y]);
    DP[x] = tmp;
  }
  trueP = DP[0];
  for (size_t x = K; x < L; x++) {
    nDP[0] = DP[0] * (1 - e[x]);
    for (size_t y = 1; y <= K; y++) nDP[y - 1] += DP[y] * (1 - e[x]);
    nDP[K] = e[x];
    trueP += nDP[0];
    DP = nDP;
    nDP.assign(K + 1, 0);
  }
  return trueP / (L - K + 1 - trueP);
}

int shq_param_block(int K, uint64_t n_true_kmers, uint64_t total_kmers, double alpha, const char* error_profile,
                    int num_deNoise, double fr, shq_params* out) {
  if (!out || K < 1 || n_true_kmers == 0) return SHQ_ERR_ARG;
  uint64_t num_true_kmers;
  if (alpha == -1) {
    if (!error_profile || !*error_profile) return SHQ_ERR_ARG;
    double t = true2falseKmer_DP(error_profile, (size_t)K);
    num_true_kmers = (uint64_t)(total_kmers * t / (1 + t));
  } else {
    num_true_kmers = (uint64_t)(total_kmers * pow(1 - alpha, K));
  }
  uint64_t num_false_kmers = total_kmers - num_true_kmers;
  if (num_deNoise < 0) {
    if (!fr) fr = 1.0 / n_true_kmers;
    num_deNoise = mean_CDF2deNoise((double)(num_true_kmers / n_true_kmers), fr);
  }
  int ave = 0;
  for (uint64_t t = num_true_kmers / n_true_kmers + 1; t; t >>= 7) ave++;
  auto slots = [&](uint64_t d) { return (uint64_t)(n_true_kmers * (ave + (double)3 / 2) + num_false_kmers * 10 / ((d + 1) * 9)); };
  uint64_t num_slots = slots((uint64_t)num_deNoise);
  uint64_t qb = 1, base = 2;
  while (base < num_slots) { qb++; base <<= 1; }
  uint64_t upper = (uint64_t)num_deNoise, tmp = num_slots;
  while (num_deNoise && tmp < (1ULL << qb)) { num_deNoise--; tmp = slots((uint64_t)num_deNoise); }
  if (tmp >= (1ULL << qb)) num_deNoise++;
  out->trigger = n_true_kmers + num_false_kmers / (uint64_t)(num_deNoise + 1);
  out->lower_bound = num_deNoise;
  tmp = (uint64_t)(n_true_kmers * (ave + (double)3 / 2));
  if (tmp > (1ULL << (qb - 1))) upper = 0;
  else {
    tmp = num_slots;
    while (tmp >= (1ULL << (qb - 1))) { upper++; tmp = slots(upper); }
    if (tmp < (1ULL << (qb - 1))) upper--;
  }
  out->num_true_kmers = num_true_kmers; out->num_false_kmers = num_false_kmers;
  out->num_deNoise = num_deNoise; out->upper_bound = (int)upper;
  out->qb = qb; out->hb = qb + 8; out->fr = fr;
  return SHQ_OK;
}

}  // extern "C"
