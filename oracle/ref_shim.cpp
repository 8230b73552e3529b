// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" shim around the UNMODIFIED reference sources, compiled from
// where they lie under /root/reference (never copied into this repo):
//   /root/reference/cqf/gqf.c      (the counting quotient filter core)
//   /root/reference/base/nthash.hpp (ntHash v1)
// Built by oracle/Makefile into oracle/_ref/libshq_ref.so.  It exists to
//   (1) pin oracle/shq_oracle.cpp (our restatement) against the real code,
//   (2) generate the golden vectors under tests/golden/,
//   (3) serve as the "reference" CPU baseline in bench.py (cpu_baseline.kind).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include <string>

// gqf.h forward-declares this type; the real definition lives in CQF_mt.h
// (boost-dependent, not buildable here).  An empty stub is enough because the
// only user, qf_clean_singleton_with_lock, is declared but defined in CQF_mt.h.
struct CQF_runtime_mt {};

#include "cqf/gqf.h"
#include "base/nthash.hpp"

extern "C" {

// ---- ntHash (base/nthash.hpp:295-309) -------------------------------------
void ref_ntpc64_base(const char* kmer, unsigned k, uint64_t* fh, uint64_t* rh) {
  NTPC64(kmer, k, *fh, *rh);
}
void ref_ntpc64_roll(unsigned char out, unsigned char in, unsigned k, uint64_t* fh, uint64_t* rh) {
  NTPC64(out, in, k, *fh, *rh);
}

// ---- CQF lifetime (cqf/gqf.c:2187-2317) -----------------------------------
void* ref_qf_new(uint64_t qb, uint64_t hb, uint32_t seed) {
  QF* qf = (QF*)calloc(1, sizeof(QF));
  qf_init(qf, 1ULL << qb, hb, 0, true, "", seed);
  return qf;
}
void ref_qf_free(void* h) {
  QF* qf = (QF*)h;
  qf_destroy(qf, true);
  free(qf);
}
void* ref_qf_load(const char* path) {
  QF* qf = (QF*)calloc(1, sizeof(QF));
  qf_deserialize(qf, path);
  return qf;
}
void ref_qf_save(void* h, const char* path) { qf_serialize((QF*)h, path); }

uint64_t ref_qf_meta(void* h, int field) {
  QF* qf = (QF*)h;
  switch (field) {
    case 0: return qf->metadata->size;
    case 1: return qf->metadata->seed;
    case 2: return qf->metadata->nslots;
    case 3: return qf->metadata->xnslots;
    case 4: return qf->metadata->key_bits;
    case 5: return qf->metadata->value_bits;
    case 6: return qf->metadata->key_remainder_bits;
    case 7: return qf->metadata->bits_per_slot;
    case 8: return qf->metadata->nblocks;
    case 9: return qf->metadata->nelts;
    case 10: return qf->metadata->ndistinct_elts;
    case 11: return qf->metadata->noccupied_slots;
    case 12: return qf->metadata->num_locks;
    case 13: return sizeof(qfmetadata);
  }
  return 0;
}
void ref_qf_set_counts(void* h, uint64_t nelts, uint64_t ndistinct) {
  QF* qf = (QF*)h;
  qf->metadata->nelts = nelts;
  qf->metadata->ndistinct_elts = ndistinct;
}
const void* ref_qf_metadata_ptr(void* h) { return ((QF*)h)->metadata; }
const void* ref_qf_blocks_ptr(void* h) { return ((QF*)h)->blocks; }

// ---- insert / query (cqf/gqf.c:2432-2469, 3092-3128) ----------------------
// Returns number of keys that were new (isNew), inserts each key with count 1
// exactly as reads_to_kmers does (cqf/CQF_mt.h:637, lock=true, spin=false;
// single-threaded so the try-lock always succeeds).
uint64_t ref_qf_insert_keys(void* h, const uint64_t* keys, uint64_t n) {
  QF* qf = (QF*)h;
  uint64_t fresh = 0;
  for (uint64_t i = 0; i < n; i++) {
    bool isNew = false;
    qf_insert_advance(qf, keys[i], 0, 1, true, false, isNew);
    fresh += isNew ? 1 : 0;
  }
  return fresh;
}
uint64_t ref_qf_insert_counted(void* h, const uint64_t* keys, const uint64_t* counts, uint64_t n) {
  QF* qf = (QF*)h;
  uint64_t fresh = 0;
  for (uint64_t i = 0; i < n; i++) {
    bool isNew = false;
    qf_insert_advance(qf, keys[i], 0, counts[i], true, true, isNew);
    fresh += isNew ? 1 : 0;
  }
  return fresh;
}
uint64_t ref_qf_count(void* h, uint64_t key) { return qf_count_key_value((QF*)h, key, 0); }
int ref_qf_count_set_traveled(void* h, uint64_t key, uint64_t* count) {
  return qf_count_key_value_set_traveled((QF*)h, key, 0, count) ? 1 : 0;
}
void ref_qf_query_batch(void* h, const uint64_t* keys, uint64_t n, uint64_t* counts, uint8_t* was) {
  for (uint64_t i = 0; i < n; i++)
    was[i] = (uint8_t)ref_qf_count_set_traveled(h, keys[i], &counts[i]);
}

// ---- iteration (cqf/gqf.c:2474-2601) --------------------------------------
uint64_t ref_qf_iterate(void* h, uint64_t* keys, uint64_t* counts, uint64_t cap) {
  QF* qf = (QF*)h;
  QFi it;
  uint64_t n = 0;
  if (!qf_iterator(qf, &it, 0)) return 0;
  do {
    uint64_t k = 0, v = 0, c = 0;
    if (qfi_get(&it, &k, &v, &c)) break;
    if (n < cap) { keys[n] = k; counts[n] = c; }
    n++;
  } while (!qfi_next(&it));
  return n;
}

// ---- deNoise sweep at t=1 -------------------------------------------------
// Control flow of qf_clean_singleton_with_lock (cqf/CQF_mt.h:999-1039), which
// cannot be included here (boost).  At one thread the region locks are no-ops,
// so only the call sequence into gqf.c matters; it is kept call-for-call,
// including the strict `start < end_bucket_id` tests that leave a one-slot
// cluster sitting exactly on a range end untouched.
static void clean_range_like_reference(QF* qf, uint64_t start_bucket_id, uint64_t end_bucket_id,
                                       uint64_t* removed) {
  uint64_t start, end;
  uint64_t start_block_idx = start_bucket_id / SLOTS_PER_BLOCK;
  uint64_t end_block_idx = end_bucket_id / SLOTS_PER_BLOCK;
  if (start_block_idx == end_block_idx) {
    qf_clean_singleton_discrete(qf, start_bucket_id, end_bucket_id, removed);
    return;
  }
  start = start_bucket_id;
  end = find_first_empty_slot(qf, start + 1) - 1;
  while ((end / SLOTS_PER_BLOCK) == start_block_idx) {
    qf_clean_singleton(qf, start, end, removed);
    start = find_first_nonempty_slot(qf, end + 1);
    end = find_first_empty_slot(qf, start + 1) - 1;
  }
  if (start / SLOTS_PER_BLOCK == start_block_idx) {
    qf_clean_singleton_with_lock_atStart(qf, start, end, removed);
    start = find_first_nonempty_slot(qf, end + 1);
    end = find_first_empty_slot(qf, start + 1) - 1;
  }
  while (start < end_bucket_id) {
    if (end / SLOTS_PER_BLOCK == end_block_idx) {
      qf_clean_singleton_discrete(qf, start, end_bucket_id, removed);
      break;
    } else {
      qf_clean_singleton(qf, start, end, removed);
    }
    start = find_first_nonempty_slot(qf, end + 1);
    end = find_first_empty_slot(qf, start + 1) - 1;
  }
}

// DeNoise mode of fastq_to_uint64kmers_prod for one worker
// (cqf/CQF_mt.h:866 sets current_index, :884-901 hands out ranges of
// min_DeNoise_len = NUM_SLOTS_TO_LOCK<<4 slots, :964).
uint64_t ref_qf_denoise(void* h) {
  QF* qf = (QF*)h;
  uint64_t removed = 0;
  uint64_t nslots = qf->metadata->nslots;
  uint64_t min_len = NUM_SLOTS_TO_LOCK << 4;
  uint64_t cur = find_first_nonempty_slot(qf, 0);
  while (cur < nslots) {
    uint64_t start = cur;
    uint64_t end = (start + min_len) > nslots ? nslots : (start + min_len);
    end = find_first_empty_slot(qf, end) - 1;
    cur = find_first_nonempty_slot(qf, end + 1);
    clean_range_like_reference(qf, start, end, &removed);
  }
  return removed;
}

// Idealised sweep: every cluster cleaned (no range-end quirk).  Used only to
// show where the real driver and the idealised semantics differ.
uint64_t ref_qf_denoise_all_clusters(void* h) {
  QF* qf = (QF*)h;
  uint64_t removed = 0;
  uint64_t xn = qf->metadata->xnslots;
  uint64_t start = find_first_nonempty_slot(qf, 0);
  while (start < qf->metadata->nslots && start < xn) {
    uint64_t end = find_first_empty_slot(qf, start + 1) - 1;
    qf_clean_singleton(qf, start, end, &removed);
    start = find_first_nonempty_slot(qf, end + 1);
  }
  return removed;
}

// check_offset (cqf/gqf.c:3056-3069) plus popcount(occupieds)==popcount(runends).
// (The reference's popcnt_occupieds reads one block past the end, so the
// popcounts are taken here from the raw block bytes instead.)
int ref_qf_check(void* h) {
  QF* qf = (QF*)h;
  const uint8_t* b = (const uint8_t*)qf->blocks;
  uint64_t bsz = qf->metadata->size / qf->metadata->nblocks;
  uint64_t po = 0, pr = 0;
  for (uint64_t i = 0; i < qf->metadata->nblocks; i++) {
    uint64_t o, r;
    memcpy(&o, b + i * bsz + 1, 8);
    memcpy(&r, b + i * bsz + 9, 8);
    po += __builtin_popcountll(o);
    pr += __builtin_popcountll(r);
  }
  if (po != pr) return 1;
  if (!check_offset(qf)) return 2;
  return 0;
}


// ---- multi-threaded CPU baseline ---------------------------------------------------
// Worker loop of fastq_to_uint64kmers_prod / reads_to_kmers (cqf/CQF_mt.h:610-731, 821-931)
// with std::thread instead of boost::thread: every key goes through the REAL
// qf_insert_advance(lock=true, spin=false); on a failed try-lock it is parked in a thread-local
// 2^16-slot QF (QBITS_LOCAL_QF, CQF_mt.h:57) that is dumped into the main filter once it holds
// more than 2^15 keys (:642-645, :588-607) and after every chunk (:833-836).
// DeNoise rounds are run by one thread between extraction phases (the reference spreads the
// sweep over its workers; the sweep is a few percent of the run).
namespace {
struct MtShared {
  QF* main_qf; const char* text; const uint64_t* chunk_ends; uint64_t nchunks; int K;
  std::atomic<uint64_t> next_chunk{0}, ndistinct{0}, nelts{0}, processed{0};
  uint64_t trigger; std::atomic<uint32_t> rounds_left{0};
  double max_seconds; std::chrono::steady_clock::time_point t0;
  std::atomic<bool> out_of_time{false};
};

void dump_local(QF* local, QF* main_qf, MtShared* sh) {
  uint64_t new_elts = 0, total = 0;
  QFi it;
  if (qf_iterator(local, &it, 0)) {
    do {
      uint64_t key = 0, value = 0, count = 0;
      if (qfi_get(&it, &key, &value, &count)) break;
      bool isNew = false;
      qf_insert_advance(main_qf, key, 0, count, true, true, isNew);
      if (isNew) new_elts++;
      total += count;
    } while (!qfi_next(&it));
    qf_reset(local);
  }
  sh->ndistinct += new_elts; sh->nelts += total;
}

void mt_worker(MtShared* sh) {
  QF local;
  qf_init(&local, 1ULL << 16, sh->main_qf->metadata->key_bits, 0, true, "", sh->main_qf->metadata->seed);
  uint32_t parked = 0;
  const __uint128_t range = sh->main_qf->metadata->range;
  const unsigned K = (unsigned)sh->K;
  for (;;) {
    if (sh->rounds_left.load() && sh->ndistinct.load() >= sh->trigger) break;      // CQF_mt.h:837
    if (sh->max_seconds > 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - sh->t0).count() > sh->max_seconds) { sh->out_of_time = true; break; }
    uint64_t c = sh->next_chunk.fetch_add(1);
    if (c >= sh->nchunks) break;
    const char* fs = sh->text + (c ? sh->chunk_ends[c - 1] : 0);
    const char* end = sh->text + sh->chunk_ends[c];
    uint64_t new_elts = 0, total_elts = 0, nput = 0;
    auto put = [&](uint64_t hash) {
      bool isNew = false; nput++;
      if (!qf_insert_advance(sh->main_qf, hash % range, 0, 1, true, false, isNew)) {
        qf_insert(&local, hash % range, 0, 1, false, false);
        if (++parked > (1ULL << 15)) { dump_local(&local, sh->main_qf, sh); parked = 0; }
      } else if (isNew) new_elts++; else total_elts++;
    };
    while (fs && fs != end) {
      fs = (const char*)memchr(fs, '\n', end - fs); fs++;
      const char* fe = (const char*)memchr(fs, '\n', end - fs);
      std::string read(fs, fe - fs);
      uint64_t hash, hash_RC;
    start_read:
      if (read.length() >= K) {
        NTPC64(read.c_str(), K, hash, hash_RC);
        put(hash < hash_RC ? hash : hash_RC);
        for (uint32_t i = K; i < read.length(); i++) {
          if (read[i] == 'N') { read = read.substr(i + 1, read.length()); goto start_read; }
          NTPC64(read[i - K], read[i], K, hash, hash_RC);
          put(hash < hash_RC ? hash : hash_RC);
        }
      }
      fs = ++fe;
      fs = (const char*)memchr(fs, '\n', end - fs); fs++;
      fs = (const char*)memchr(fs, '\n', end - fs); fs++;
    }
    sh->ndistinct += new_elts; sh->nelts += new_elts + total_elts; sh->processed += nput;
    if (parked) { dump_local(&local, sh->main_qf, sh); parked = 0; }
  }
  qf_destroy(&local, true);
}
}  // namespace

// Returns elapsed seconds.  Stops taking new chunks after max_seconds (0 = no limit).
double ref_count_mt(void* h, const char* text, const uint64_t* chunk_ends, uint64_t nchunks, int K, int nthreads,
                    uint64_t trigger, uint32_t rounds, int end_denoise, double max_seconds, uint64_t* kmers_done,
                    uint64_t* chunks_done, uint32_t* rounds_done) {
  MtShared sh;
  sh.main_qf = (QF*)h; sh.text = text; sh.chunk_ends = chunk_ends; sh.nchunks = nchunks; sh.K = K;
  sh.trigger = trigger; sh.rounds_left = rounds; sh.max_seconds = max_seconds;
  sh.t0 = std::chrono::steady_clock::now();
  uint32_t done_rounds = 0;
  for (;;) {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(mt_worker, &sh);
    for (auto& t : th) t.join();
    bool files_left = sh.next_chunk.load() < nchunks && !sh.out_of_time;
    bool need = sh.rounds_left.load() && sh.ndistinct.load() >= sh.trigger;
    if (need || (end_denoise && !files_left)) {                                    // CQF_mt.h:860
      if (sh.rounds_left.load()) sh.rounds_left--;
      uint64_t removed = ref_qf_denoise(h);
      sh.ndistinct -= removed; sh.nelts -= removed;
      done_rounds++;
    }
    if (!files_left) break;
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - sh.t0).count();
  QF* qf = (QF*)h;
  qf->metadata->nelts = sh.nelts; qf->metadata->ndistinct_elts = sh.ndistinct;
  if (kmers_done) *kmers_done = sh.processed.load();
  if (chunks_done) *chunks_done = sh.next_chunk.load() < nchunks ? sh.next_chunk.load() : nchunks;
  if (rounds_done) *rounds_done = done_rounds;
  return secs;
}

}  // extern "C"
