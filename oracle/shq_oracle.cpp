// oracle/shq_oracle.cpp -- TEST INFRASTRUCTURE, not product code.
//
// CPU restatement of the counting half of SH-assembly's hot path, written from
// the reference's behaviour (file:line cited per function, paths relative to
// /root/reference).  It is the checker for the CUDA path: only tests/,
// __graft_entry__.smoke() and bench.py's CPU legs may load it.  It is pinned
// against the unmodified reference core (oracle/_ref/libshq_ref.so) and the
// golden vectors generated from it (tests/golden/).  See shq_oracle.h.
//
// Model: the CQF is held as a (key -> count) map; the reference's block image
// is a pure function of that multiset (verified against gqf.c), so the image,
// the iterator, the deNoise sweep and the traveled bits are all derived from
// the sorted multiset plus the canonical layout rule below.
#include "shq_oracle.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------- ntHash ----
// Seeds: base/nthash.hpp:24-28.  msTab[c][j] == rol(seed_c, j) (:30-118); the
// reverse strand indexes the table with (c & cpOff) where cpOff = 0x07 (:15),
// which maps 'A'&7=1->T, 'C'&7=3->G, 'G'&7=7->C, 'T'&7=4->A (:85-87).
const uint64_t kSeedA = 0x3c8bfbb395c60474ULL, kSeedC = 0x3193c18562a02b4cULL,
               kSeedG = 0x20323ed082572324ULL, kSeedT = 0x295549f54be24456ULL;

inline uint64_t rotl(uint64_t v, unsigned s) { s &= 63; return s ? (v << s) | (v >> (64 - s)) : v; }
inline uint64_t rotr(uint64_t v, unsigned s) { s &= 63; return s ? (v >> s) | (v << (64 - s)) : v; }

inline uint64_t seed_of(unsigned char c) {  // row of msTab (nthash.hpp:85-118)
  switch (c) {
    case 'A': case 'a': case 4: return kSeedA;
    case 'C': case 'c': case 7: return kSeedC;
    case 'G': case 'g': case 3: return kSeedG;
    case 'T': case 't': case 1: return kSeedT;
    default: return 0;
  }
}

void nt_base(const char* s, unsigned k, uint64_t& fh, uint64_t& rh) {  // nthash.hpp:295-302
  fh = rh = 0;
  for (unsigned i = 0; i < k; i++) {
    unsigned char c = (unsigned char)s[i];
    fh ^= rotl(seed_of(c), (k - 1 - i) % 64);
    rh ^= rotl(seed_of(c & 0x07), i % 64);
  }
}
void nt_roll(unsigned char out, unsigned char in, unsigned k, uint64_t& fh, uint64_t& rh) {  // :305-309
  fh = rotl(fh, 1) ^ rotl(seed_of(out), k % 64) ^ seed_of(in);
  rh = rotr(rh, 1) ^ rotl(seed_of(out & 0x07), 63) ^ rotl(seed_of(in & 0x07), (k - 1) % 64);
}

// ------------------------------------------------------- reads_to_kmers ----
// cqf/CQF_mt.h:610-731.  Line 2 of every 4-line record; no N test on the first
// K bases of a (sub)read; an upper-case 'N' at index >= K restarts after it.
template <class Emit>
bool for_each_key(const char* text, size_t n, int K, int hb, Emit emit) {
  const uint64_t mask = hb >= 64 ? ~0ULL : ((1ULL << hb) - 1);  // hash % range, gqf.c:2230-2231
  size_t fs = 0;
  while (fs != n) {
    const char* p = (const char*)memchr(text + fs, '\n', n - fs);
    if (!p) return false;
    fs = (size_t)(p - text) + 1;
    p = (const char*)memchr(text + fs, '\n', n - fs);
    if (!p) return false;
    size_t fe = (size_t)(p - text);
    const char* read = text + fs;
    size_t len = fe - fs, s0 = 0;
    while (s0 <= len && len - s0 >= (size_t)K) {
      uint64_t fh, rh;
      nt_base(read + s0, (unsigned)K, fh, rh);
      emit((fh < rh ? fh : rh) & mask);
      bool restart = false;
      for (size_t i = s0 + K; i < len; i++) {
        if (read[i] == 'N') { s0 = i + 1; restart = true; break; }
        nt_roll((unsigned char)read[i - K], (unsigned char)read[i], (unsigned)K, fh, rh);
        emit((fh < rh ? fh : rh) & mask);
      }
      if (!restart) break;
    }
    fs = fe + 1;
    for (int l = 0; l < 2; l++) {
      p = (const char*)memchr(text + fs, '\n', n - fs);
      if (!p) return false;
      fs = (size_t)(p - text) + 1;
    }
  }
  return true;
}

// -------------------------------------------------------------- chunker ----
struct MemFile {
  const char* buf; size_t len; size_t pos; bool eof; std::string tail; int idx;
};
struct Chunker {
  std::deque<MemFile*> q;  // boost::lockfree::queue used as a FIFO, CQF_mt.h:336
  std::vector<MemFile> files;
  uint64_t part_size;
  int num_files;
};

bool skip_next_eol(const char* part, int64_t& pos, int64_t max_pos) {  // CQF_mt.h:573-585
  int64_t i;
  for (i = pos; i < max_pos - 2; ++i)
    if ((part[i] == '\n' || part[i] == '\r') && !(part[i + 1] == '\n' || part[i + 1] == '\r')) break;
  if (i >= max_pos - 2) return false;
  pos = i + 1;
  return true;
}

// fastq_read_parts, CQF_mt.h:735-816.  0 = file finished, 1 = chunk, -2 = "Wrong input file".
int read_parts(MemFile* f, uint64_t part_size, std::string& chunk) {
  const int64_t OVERHEAD_SIZE = 65535;
  if (f->eof) return 0;                                   // :747
  std::string part = f->tail;                              // memcpy(part, part_buffer, part_filled)
  size_t part_filled = part.size();
  size_t readed = std::min<size_t>(part_size, f->len - f->pos);   // fread
  part.append(f->buf + f->pos, readed);
  f->pos += readed;
  if (readed < part_size) f->eof = true;                   // feof() after a short read
  int64_t total_filled = (int64_t)(part_filled + readed);
  if ((int64_t)part_filled >= OVERHEAD_SIZE) return -2;    // :765-769
  if (f->eof) { chunk.swap(part); f->tail.clear(); return 1; }  // :770-776
  int64_t size;
  {
    int64_t line_start[9];
    int j;
    int64_t i = total_filled - OVERHEAD_SIZE / 2;
    for (j = 0; j < 9; ++j) {
      if (!skip_next_eol(part.data(), i, total_filled)) break;
      line_start[j] = i;
    }
    if (j < 9) size = 0;
    else {
      int k;
      const char* P = part.data();
      for (k = 0; k < 4; ++k) {
        if (P[line_start[k]] == '@' && P[line_start[k + 2]] == '+') {
          if (P[line_start[k + 2] + 1] == '\n' || P[line_start[k + 2] + 1] == '\r') break;
          if (line_start[k + 1] - line_start[k] == line_start[k + 3] - line_start[k + 2] &&
              memcmp(P + line_start[k] + 1, P + line_start[k + 2] + 1,
                     (size_t)(line_start[k + 3] - line_start[k + 2] - 1)) == 0)
            break;
        }
      }
      size = (k == 4) ? 0 : line_start[k];
    }
  }
  f->tail.assign(part.data() + size, (size_t)(total_filled - size));
  part.resize((size_t)size);
  chunk.swap(part);
  return 1;
}

// ------------------------------------------------------------ CQF model ----
struct Entry { uint64_t key, count; };

struct Cqf {
  int qb, hb; uint32_t seed;
  uint64_t nslots, xnslots, nblocks;
  std::unordered_map<uint64_t, uint64_t> m;
  std::unordered_set<uint64_t> traveled;
  uint64_t ndistinct = 0, nelts = 0;   // runtime counters, CQF_mt.h:729-730, 1037-1038

  Cqf(int qb_, int hb_, uint32_t seed_) : qb(qb_), hb(hb_), seed(seed_) {
    nslots = 1ULL << qb;
    xnslots = nslots + (uint64_t)(10 * sqrt((double)nslots));   // gqf.c:2198
    nblocks = (xnslots + 63) / 64;
  }
  std::vector<Entry> sorted() const {
    std::vector<Entry> v; v.reserve(m.size());
    for (auto& kv : m) v.push_back({kv.first, kv.second});
    std::sort(v.begin(), v.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
    return v;
  }
};

// encode_counter, gqf.c:1225-1255 (8-bit slots: base 128, MSB = "more digits")
void encode_entry(uint8_t rem, uint64_t count, std::vector<uint8_t>& out) {
  out.clear();
  out.push_back(rem);
  if (count <= 1) return;
  uint64_t c = count - 1;
  uint8_t digits[12]; int nd = 0;
  digits[nd++] = (uint8_t)(c % 128); c /= 128;
  while (c) { digits[nd++] = (uint8_t)((c % 128) | 0x80); c /= 128; }
  uint8_t first = digits[nd - 1];
  if (first > rem) out.push_back(0);
  for (int i = nd - 1; i >= 0; i--) out.push_back(digits[i]);
}

struct Layout {
  std::vector<uint64_t> run_q, run_start, run_end;    // per occupied quotient
  std::vector<uint64_t> entry_slot;                   // first slot of each sorted entry
  std::vector<uint32_t> entry_run;
  std::vector<uint64_t> clus_start, clus_end;         // maximal stretches of used slots
  bool overflow = false;
};

// Canonical layout: runs in quotient order, start(q) = max(q, end(prev)+1)
// (what repeated insert1_advance / shift_remainders converge to, gqf.c:1614-1915).
Layout make_layout(const Cqf& q, const std::vector<Entry>& v, std::vector<uint8_t>* slots) {
  Layout L;
  if (slots) slots->assign(q.nblocks * 64, 0);
  std::vector<uint8_t> enc;
  uint64_t next_free = 0;
  L.entry_slot.resize(v.size()); L.entry_run.resize(v.size());
  size_t i = 0;
  while (i < v.size()) {
    uint64_t quo = v[i].key >> 8;
    uint64_t start = std::max(quo, next_free), pos = start;
    size_t j = i;
    for (; j < v.size() && (v[j].key >> 8) == quo; j++) {
      encode_entry((uint8_t)(v[j].key & 0xFF), v[j].count, enc);
      L.entry_slot[j] = pos; L.entry_run[j] = (uint32_t)L.run_q.size();
      for (uint8_t b : enc) {
        if (pos >= q.xnslots) L.overflow = true;
        else if (slots) (*slots)[pos] = b;
        pos++;
      }
    }
    L.run_q.push_back(quo); L.run_start.push_back(start); L.run_end.push_back(pos - 1);
    if (L.clus_end.empty() || start > L.clus_end.back() + 1) { L.clus_start.push_back(start); L.clus_end.push_back(pos - 1); }
    else L.clus_end.back() = pos - 1;
    next_free = pos;
    i = j;
  }
  return L;
}

// find_first_nonempty_slot (gqf.c:751-774): next occupied quotient >= from, else xnslots.
uint64_t fne(const Cqf& q, const Layout& L, uint64_t from) {
  auto it = std::lower_bound(L.run_q.begin(), L.run_q.end(), from);
  return it == L.run_q.end() ? q.xnslots : *it;
}
// find_first_empty_slot (gqf.c:738-748): first slot >= from not covered by a run.
uint64_t ffe(const Layout& L, uint64_t from) {
  size_t i = std::upper_bound(L.clus_start.begin(), L.clus_start.end(), from) - L.clus_start.begin();
  if (i == 0) return from;
  return L.clus_end[i - 1] >= from ? L.clus_end[i - 1] + 1 : from;
}

// One DeNoise round at t=1: the range hand-out of CQF_mt.h:884-901 (ranges of
// NUM_SLOTS_TO_LOCK<<4 = 2^20 slots, :964, stretched to a cluster end) and the
// per-range cleaning of qf_clean_singleton_with_lock (:999-1039) ->
// qf_clean_singleton{,_discrete,_with_lock_atStart} (gqf.c:2792-3040).  All of
// those loop `while (start < end_bucket_id)`, so a cluster that is a single
// slot sitting exactly on the range end is never visited: its (count-1) key
// survives the sweep.  Everything else with count 1 is deleted.
uint64_t denoise(Cqf& q, bool ideal) {
  std::vector<Entry> v = q.sorted();
  std::unordered_set<uint64_t> keep;
  if (!ideal) {
    Layout L = make_layout(q, v, nullptr);
    const uint64_t min_len = (1ULL << 16) << 4;
    uint64_t cur = fne(q, L, 0);
    while (cur < q.nslots) {
      uint64_t start = cur;
      uint64_t end = (start + min_len) > q.nslots ? q.nslots : (start + min_len);
      end = ffe(L, end) - 1;
      cur = fne(q, L, end + 1);
      // last cluster of the range == [end,end] ?  (a one-slot cluster holds one count-1 key
      // whose quotient is the slot itself)
      size_t ci = std::upper_bound(L.clus_start.begin(), L.clus_start.end(), end) - L.clus_start.begin();
      if (ci > 0 && L.clus_start[ci - 1] == end && L.clus_end[ci - 1] == end && end >= start) {
        size_t ei = std::lower_bound(v.begin(), v.end(), end << 8,
                                     [](const Entry& e, uint64_t k) { return e.key < k; }) - v.begin();
        if (ei < v.size() && (v[ei].key >> 8) == end && v[ei].count == 1) keep.insert(v[ei].key);
      }
    }
  }
  uint64_t removed = 0;
  for (auto& e : v)
    if (e.count == 1 && !keep.count(e.key)) { q.m.erase(e.key); q.traveled.erase(e.key); removed++; }
  q.nelts -= removed; q.ndistinct -= removed;   // CQF_mt.h:1037-1038
  return removed;
}

inline void put64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
inline uint64_t get64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

}  // namespace

extern "C" {

void orc_ntpc64_base(const char* kmer, unsigned k, uint64_t* fh, uint64_t* rh) { nt_base(kmer, k, *fh, *rh); }
void orc_ntpc64_roll(unsigned char out, unsigned char in, unsigned k, uint64_t* fh, uint64_t* rh) { nt_roll(out, in, k, *fh, *rh); }

uint64_t orc_reads_to_keys(const char* text, size_t n, int K, int hb, uint64_t* keys, uint64_t cap) {
  uint64_t cnt = 0;
  bool ok = for_each_key(text, n, K, hb, [&](uint64_t k) { if (cnt < cap) keys[cnt] = k; cnt++; });
  return ok ? cnt : (uint64_t)-1;
}

void* orc_chunker_new(int nfiles, const char* const* bufs, const size_t* lens, uint64_t part_size) {
  Chunker* c = new Chunker();
  c->part_size = part_size ? part_size : (1ULL << 23);
  c->files.resize(nfiles);
  for (int i = 0; i < nfiles; i++) c->files[i] = MemFile{bufs[i], lens[i], 0, false, std::string(), i};
  for (int i = 0; i < nfiles; i++) c->q.push_back(&c->files[i]);
  c->num_files = nfiles;
  return c;
}
int64_t orc_chunker_next(void* cv, char* out, size_t cap, int* file_idx) {
  Chunker* c = (Chunker*)cv;
  while (c->num_files) {                       // seqFile_batch::getDataChunk, CQF_mt.h:364-390
    MemFile* f = c->q.front(); c->q.pop_front();
    std::string chunk;
    int r = read_parts(f, c->part_size, chunk);
    if (r == -2) return -2;
    if (r == 1) {
      c->q.push_back(f);
      if (file_idx) *file_idx = f->idx;
      if (chunk.size() > cap) return -3;
      memcpy(out, chunk.data(), chunk.size());
      return (int64_t)chunk.size();
    }
    c->num_files--;
  }
  return -1;
}
void orc_chunker_free(void* c) { delete (Chunker*)c; }

void* orc_cqf_new(int qb, int hb, uint32_t seed) { return new Cqf(qb, hb, seed); }
void orc_cqf_free(void* q) { delete (Cqf*)q; }

uint64_t orc_cqf_insert_keys(void* qv, const uint64_t* keys, uint64_t n) {
  Cqf* q = (Cqf*)qv;
  uint64_t fresh = 0;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t& c = q->m[keys[i]];
    if (c == 0) fresh++;
    c++;
  }
  q->ndistinct += fresh; q->nelts += n;
  return fresh;
}
uint64_t orc_cqf_insert_counted(void* qv, const uint64_t* keys, const uint64_t* counts, uint64_t n) {
  Cqf* q = (Cqf*)qv;
  uint64_t fresh = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (!counts[i]) continue;
    uint64_t& c = q->m[keys[i]];
    if (c == 0) fresh++;
    c += counts[i];
    q->nelts += counts[i];
  }
  q->ndistinct += fresh;
  return fresh;
}
void orc_cqf_counts(void* qv, uint64_t* nd, uint64_t* ne) { Cqf* q = (Cqf*)qv; *nd = q->ndistinct; *ne = q->nelts; }
void orc_cqf_set_counts(void* qv, uint64_t nd, uint64_t ne) { Cqf* q = (Cqf*)qv; q->ndistinct = nd; q->nelts = ne; }
uint64_t orc_cqf_denoise(void* q) { return denoise(*(Cqf*)q, false); }
uint64_t orc_cqf_denoise_ideal(void* q) { return denoise(*(Cqf*)q, true); }

uint64_t orc_cqf_iterate(void* qv, uint64_t* keys, uint64_t* counts, uint64_t cap) {  // qfi_*, gqf.c:2474-2601
  Cqf* q = (Cqf*)qv;
  std::vector<Entry> v = q->sorted();
  for (size_t i = 0; i < v.size() && i < cap; i++) { keys[i] = v[i].key; counts[i] = v[i].count; }
  return v.size();
}
uint64_t orc_cqf_image_size(void* qv) { return 128 + ((Cqf*)qv)->nblocks * 89; }

// qf_serialize (gqf.c:2379-2394): 128-byte qfmetadata (gqf.h:62-77) + packed 89-byte blocks (gqf.c:63-86).
int orc_cqf_export(void* qv, uint8_t* dst, uint64_t cap) {
  Cqf* q = (Cqf*)qv;
  uint64_t total = 128 + q->nblocks * 89;
  if (cap < total) return -1;
  std::vector<Entry> v = q->sorted();
  std::vector<uint8_t> slots;
  Layout L = make_layout(*q, v, &slots);
  if (L.overflow) return -2;
  memset(dst, 0, total);
  put64(dst + 0, q->nblocks * 89);
  uint32_t seed = q->seed; memcpy(dst + 8, &seed, 4);
  put64(dst + 16, q->nslots); put64(dst + 24, q->xnslots);
  put64(dst + 32, (uint64_t)q->hb); put64(dst + 40, 0);
  put64(dst + 48, (uint64_t)(q->hb - q->qb)); put64(dst + 56, (uint64_t)(q->hb - q->qb));
  unsigned __int128 range = (unsigned __int128)q->nslots << (q->hb - q->qb);
  memcpy(dst + 64, &range, 16);
  put64(dst + 80, q->nblocks); put64(dst + 88, q->nelts); put64(dst + 96, q->ndistinct);
  put64(dst + 104, 0); put64(dst + 112, q->xnslots / 65536 + 2);
  uint8_t* B = dst + 128;
  for (uint64_t b = 0; b < q->nblocks; b++) memcpy(B + b * 89 + 25, slots.data() + b * 64, 64);
  auto setbit = [&](uint64_t slot, int field) {
    uint8_t* w = B + (slot / 64) * 89 + 1 + 8 * field;
    put64(w, get64(w) | (1ULL << (slot % 64)));
  };
  for (size_t r = 0; r < L.run_q.size(); r++) { setbit(L.run_q[r], 0); setbit(L.run_end[r], 1); }
  for (size_t i = 0; i < v.size(); i++) if (q->traveled.count(v[i].key)) setbit(L.entry_slot[i], 2);
  // offsets: block_offset_strict (gqf.c:598-601) saturated at 255 (:1395-1399, :2868-2875)
  size_t r = 0;
  uint64_t last_end = 0; bool have = false;
  for (uint64_t b = 1; b < q->nblocks; b++) {
    while (r < L.run_q.size() && L.run_q[r] < 64 * b) { last_end = L.run_end[r]; have = true; r++; }
    uint64_t off = 0;
    if (have && last_end + 1 > 64 * b) off = last_end - 64 * b + 1;
    B[b * 89] = (uint8_t)std::min<uint64_t>(off, 255);
  }
  return 0;
}

void* orc_cqf_import(const uint8_t* src, uint64_t n) {
  if (n < 128) return nullptr;
  uint64_t nslots = get64(src + 16), hb = get64(src + 32), nblocks = get64(src + 80);
  uint32_t seed; memcpy(&seed, src + 8, 4);
  int qb = 0; while ((1ULL << qb) < nslots) qb++;
  if (n < 128 + nblocks * 89 || get64(src + 56) != 8) return nullptr;
  Cqf* q = new Cqf(qb, (int)hb, seed);
  const uint8_t* B = src + 128;
  auto slot = [&](uint64_t s) { return B[(s / 64) * 89 + 25 + (s % 64)]; };
  auto bit = [&](uint64_t s, int f) { return (get64(B + (s / 64) * 89 + 1 + 8 * f) >> (s % 64)) & 1; };
  uint64_t next_free = 0, total = nblocks * 64;
  for (uint64_t quo = 0; quo < total; quo++) {
    if (!bit(quo, 0)) continue;
    uint64_t pos = std::max(quo, next_free), end = pos;
    while (!bit(end, 1)) end++;
    while (pos <= end) {                         // decode_counter, gqf.c:1259-1299
      uint64_t first = pos;
      uint64_t rem = slot(pos), cnt = 1;
      if (pos != end && slot(pos + 1) <= rem) {
        uint64_t e = pos + 1, d = slot(e), c = 0;
        if (d == 0) { e++; d = slot(e); }
        while (d & 0x80) { c = c * 128 + (d & 0x7F); e++; d = slot(e); }
        c = c * 128 + d; cnt = c + 1; pos = e;
      }
      uint64_t key = (quo << 8) | rem;
      q->m[key] = cnt;
      if (bit(first, 2)) q->traveled.insert(key);
      pos++;
    }
    next_free = end + 1;
  }
  q->nelts = get64(src + 88); q->ndistinct = get64(src + 96);
  return q;
}

// qf_count_key_value_set_traveled, gqf.c:3092-3128
int orc_cqf_query_set_traveled(void* qv, uint64_t key, uint64_t* count) {
  Cqf* q = (Cqf*)qv;
  auto it = q->m.find(key);
  if (it == q->m.end()) { *count = 0; return 0; }
  *count = it->second;
  return q->traveled.insert(key).second ? 0 : 1;
}
void orc_cqf_query_batch(void* q, const uint64_t* keys, uint64_t n, uint64_t* counts, uint8_t* was) {
  for (uint64_t i = 0; i < n; i++) was[i] = (uint8_t)orc_cqf_query_set_traveled(q, keys[i], &counts[i]);
}

// fastq_to_uint64kmers_prod at one worker (CQF_mt.h:821-931): ExtractKmer until the
// trigger (tested after every whole chunk, :837) or until the files run out, then a
// DeNoise round, and so on; --endDeNoise adds a round once num_files hits 0 (:860).
void* orc_count_run(int nfiles, const char* const* bufs, const size_t* lens, uint64_t part_size,
                    int K, int qb, int hb, uint64_t trigger, uint32_t num_denoise, int end_denoise,
                    int* rounds_done) {
  Cqf* q = new Cqf(qb, hb, 2038074761u);
  Chunker* c = (Chunker*)orc_chunker_new(nfiles, bufs, lens, part_size);
  int rounds = 0;
  enum { Extract, DeNoise, Idle } mode = Extract;
  while (mode != Idle) {
    if (mode == Extract) {
      while (c->num_files) {
        MemFile* f = c->q.front(); c->q.pop_front();
        std::string chunk;
        int r = read_parts(f, c->part_size, chunk);
        if (r == 1) {
          c->q.push_back(f);
          uint64_t fresh = 0, all = 0;
          for_each_key(chunk.data(), chunk.size(), K, hb, [&](uint64_t k) {
            uint64_t& cnt = q->m[k];
            if (cnt == 0) fresh++;
            cnt++; all++;
          });
          q->ndistinct += fresh; q->nelts += all;
          if (num_denoise && q->ndistinct >= trigger) break;
        } else {
          c->num_files--;
        }
      }
      bool need = num_denoise && q->ndistinct >= trigger;
      if (need || (end_denoise && !c->num_files)) {
        if (num_denoise) num_denoise--;
        mode = DeNoise;
      } else if (!c->num_files) {
        mode = Idle;
      }
    } else {
      denoise(*q, false);
      rounds++;
      mode = c->num_files ? Extract : Idle;
    }
  }
  orc_chunker_free(c);
  if (rounds_done) *rounds_done = rounds;
  return q;
}

// ------------------------------------------------------ parameter block ----
// boost::math::cdf(poisson(mean), x) == Q(floor(x)+1, mean); for an integer first
// argument that is the finite sum  e^-mean * sum_{j<=floor(x)} mean^j / j!.
double orc_cdfpoi(double x, double mean) {                       // CQF_mt.h:79-82
  if (x < 0) return 0;
  long k = (long)floor(x);
  if (mean <= 0) return 1.0;
  double sum = 0, lm = log(mean);
  for (long j = 0; j <= k; j++) sum += exp(-mean + j * lm - lgamma((double)j + 1.0));
  return sum > 1.0 ? 1.0 : sum;
}
double orc_cdfpoi_positive(double x, double mean) {              // CQF_mt.h:84-92
  double c0 = orc_cdfpoi(0, mean);
  return (orc_cdfpoi(x, mean) - c0) / (1 - c0);
}
int orc_mean_CDF2deNoise(double mean, double cdf_desired) {      // CQF_mt.h:94-133
  int start = 0, end = (int)(mean + 1), mid;
  double cdf0 = orc_cdfpoi(0, mean);
  auto cdf_positive = [&](double x) { return (orc_cdfpoi(x, mean) - cdf0) / (1 - cdf0); };
  while (cdf_positive(end) < cdf_desired) end *= 2;
  while (start <= end) {
    if (start == end) return start;
    else if (start + 1 == end) {
      double t1 = cdf_positive(start), t2 = cdf_positive(end);
      if (t2 <= cdf_desired) return end;
      else if (t1 <= cdf_desired) return start;
      else return std::max(start - 1, 0);
    }
    mid = (start + end) / 2;
    double cdf = cdf_positive(mid);
    if (cdf < cdf_desired) start = mid + 1;
    else if (cdf > cdf_desired) end = mid - 1;
    else return start;
  }
  return start;
}

double orc_true2falseKmer_DP(const double* baseErrors, size_t seq_len, size_t K) {  // true2falseKmer_DP.cpp:12-51
  std::vector<double> DP(K + 1, 0), new_DP(K + 1, 0);
  double tmp = 1, trueP;
  for (size_t x = 0; x < K; x++) tmp *= (1 - baseErrors[x]);
  DP[0] = tmp;
  for (size_t x = 1; x <= K; x++) {
    tmp = baseErrors[x - 1];
    for (size_t y = x; y < K; y++) tmp *= (1 - baseErrors[y]);
    DP[x] = tmp;
  }
  trueP = DP[0];
  for (size_t x = K; x < seq_len; x++) {
    new_DP[0] = DP[0] * (1 - baseErrors[x]);
    for (size_t y = 1; y <= K; y++) new_DP[y - 1] += DP[y] * (1 - baseErrors[x]);
    new_DP[K] = baseErrors[x];
    trueP += new_DP[0];
    DP = new_DP;
    new_DP.assign(K + 1, 0);
  }
  return trueP / (seq_len - K + 1 - trueP);
}

void orc_param_block(int K, uint64_t n_true_kmers, uint64_t total_kmers, double alpha, double true2false,
                     int num_deNoise, double fr, orc_params* out) {  // CQF-deNoise.cpp:83-161
  uint64_t num_true_kmers, num_false_kmers;
  if (alpha == -1) {
    double tmp = true2false;
    num_true_kmers = (uint64_t)(total_kmers * tmp / (1 + tmp));
  } else {
    num_true_kmers = (uint64_t)(total_kmers * pow(1 - alpha, K));
  }
  num_false_kmers = total_kmers - num_true_kmers;
  if (num_deNoise < 0) {
    if (!fr) fr = 1.0 / n_true_kmers;
    num_deNoise = orc_mean_CDF2deNoise((double)(num_true_kmers / n_true_kmers), fr);
  }
  int ave = 0;
  uint64_t tmp = num_true_kmers / n_true_kmers + 1;
  while (tmp) { tmp >>= 7; ave++; }
  uint64_t num_slots = (uint64_t)(n_true_kmers * (ave + (double)3 / 2) + num_false_kmers * 10 / ((num_deNoise + 1) * 9));
  uint64_t qb = 1, base = 2;
  while (base < num_slots) { qb++; base <<= 1; }
  uint64_t upper = (uint64_t)num_deNoise, lower, num_slots_tmp = num_slots;
  while (num_deNoise && num_slots_tmp < (1ULL << qb)) {
    num_deNoise--;
    num_slots_tmp = (uint64_t)(n_true_kmers * (ave + (double)3 / 2) + num_false_kmers * 10 / ((num_deNoise + 1) * 9));
  }
  if (num_slots_tmp >= (1ULL << qb)) num_deNoise++;
  uint64_t trigger = n_true_kmers + num_false_kmers / (num_deNoise + 1);
  lower = (uint64_t)num_deNoise;
  num_slots_tmp = (uint64_t)(n_true_kmers * (ave + (double)3 / 2));
  if (num_slots_tmp > (1ULL << (qb - 1))) upper = 0;
  else {
    num_slots_tmp = num_slots;
    while (num_slots_tmp >= (1ULL << (qb - 1))) {
      upper++;
      num_slots_tmp = (uint64_t)(n_true_kmers * (ave + (double)3 / 2) + num_false_kmers * 10 / ((upper + 1) * 9));
    }
    if (num_slots_tmp < (1ULL << (qb - 1))) upper--;
  }
  out->num_true_kmers = num_true_kmers; out->num_false_kmers = num_false_kmers;
  out->num_deNoise = num_deNoise; out->lower_bound = (int)lower; out->upper_bound = (int)upper;
  out->qb = qb; out->hb = qb + 8; out->trigger = trigger; out->fr = fr;
}

}  // extern "C"
