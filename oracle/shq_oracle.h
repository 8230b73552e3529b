/* oracle/shq_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * C interface of the CPU restatement of SH-assembly's counting / deNoise /
 * unitig hot path (see shq_oracle.cpp and contiger_oracle.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liboracle.so; the product (libshq.so) never links or calls it.
 *
 * Parity pinning: the reference ships NO tests, golden vectors or fixtures
 * for this path, so the oracle is pinned against the
 * reference's own code run here: oracle/_ref/libshq_ref.so (unmodified
 * gqf.c + nthash.hpp) and the vectors under tests/golden/ generated from it
 * by tests/golden/make_golden.py.  The Poisson CDF used by the automatic
 * deNoise-round selection comes from Boost.Math (absent, version unpinned):
 * that one scalar path is "parity unpinned" and is checked against scipy only.
 */
#ifndef SHQ_ORACLE_H
#define SHQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ntHash v1 canonical (base/nthash.hpp:295-309) */
void orc_ntpc64_base(const char* kmer, unsigned k, uint64_t* fh, uint64_t* rh);
void orc_ntpc64_roll(unsigned char out, unsigned char in, unsigned k, uint64_t* fh, uint64_t* rh);

/* reads_to_kmers key stream (cqf/CQF_mt.h:610-731).  Returns number of keys;
 * writes at most cap of them.  Returns (uint64_t)-1 on malformed input (the
 * reference would dereference NULL+1 there). */
uint64_t orc_reads_to_keys(const char* text, size_t n, int K, int hb, uint64_t* keys, uint64_t cap);

/* fastq_read_parts chunker over in-memory "files" (cqf/CQF_mt.h:735-816 and
 * the round-robin queue :364-390, :824-856).  part_size is 1<<23 in the
 * reference; tests may shrink it.  Each call yields the next chunk as a copy
 * into `out` (capacity cap) and returns its size, or -1 at the end, or -2 on
 * the reference's "Wrong input file" exit. */
void* orc_chunker_new(int nfiles, const char* const* bufs, const size_t* lens, uint64_t part_size);
int64_t orc_chunker_next(void* c, char* out, size_t cap, int* file_idx);
void orc_chunker_free(void* c);

/* CQF as (key,count) multiset + canonical layout (cqf/gqf.c) */
void* orc_cqf_new(int qb, int hb, uint32_t seed);
void orc_cqf_free(void* q);
uint64_t orc_cqf_insert_keys(void* q, const uint64_t* keys, uint64_t n);  /* returns #new */
uint64_t orc_cqf_insert_counted(void* q, const uint64_t* keys, const uint64_t* counts, uint64_t n);
void orc_cqf_counts(void* q, uint64_t* ndistinct, uint64_t* nelts);
void orc_cqf_set_counts(void* q, uint64_t ndistinct, uint64_t nelts);
uint64_t orc_cqf_denoise(void* q);            /* reference driver semantics (range-end quirk) */
uint64_t orc_cqf_denoise_ideal(void* q);      /* every count-1 key removed */
uint64_t orc_cqf_iterate(void* q, uint64_t* keys, uint64_t* counts, uint64_t cap);
uint64_t orc_cqf_image_size(void* q);         /* 128 + nblocks*89 */
int orc_cqf_export(void* q, uint8_t* dst, uint64_t cap); /* .cqf bytes; <0 on capacity overflow */
void* orc_cqf_import(const uint8_t* src, uint64_t n);    /* parse .cqf bytes (incl. traveled) */
int orc_cqf_query_set_traveled(void* q, uint64_t key, uint64_t* count);
void orc_cqf_query_batch(void* q, const uint64_t* keys, uint64_t n, uint64_t* counts, uint8_t* was);

/* Whole CQF-deNoise run at t=1 over in-memory files (cqf/CQF_mt.h:821-931,
 * 959-995).  Returns a cqf handle; rounds_done receives #sweeps executed. */
void* orc_count_run(int nfiles, const char* const* bufs, const size_t* lens, uint64_t part_size,
                    int K, int qb, int hb, uint64_t trigger, uint32_t num_denoise, int end_denoise,
                    int* rounds_done);

/* Parameter block (src/CQF-deNoise.cpp:83-161, cqf/CQF_mt.h:84-133) */
typedef struct {
  uint64_t num_true_kmers, num_false_kmers;
  int num_deNoise;          /* effective */
  int lower_bound, upper_bound;
  uint64_t qb, hb;
  uint64_t trigger;         /* n_distinct_elts_for_DeNoise */
  double fr;                /* as printed in the banner */
} orc_params;
void orc_param_block(int K, uint64_t n_true, uint64_t N_total, double alpha, double true2false,
                     int deNoise, double fr, orc_params* out);
double orc_cdfpoi(double x, double mean);
double orc_cdfpoi_positive(double x, double mean);
int orc_mean_CDF2deNoise(double mean, double cdf_desired);
double orc_true2falseKmer_DP(const double* base_errors, size_t seq_len, size_t K);

/* Contiger (src/contig_assembly.cpp live parts), sequential restatement; see contiger_oracle.cpp.
 * flags: bit0 = master seed rule (len/2) instead of worker rule (len/2 - K/2); bit1 = LIFO job order. */
void* orc_contiger_run(void* cqf, int hb, int nfiles, const char* const* bufs, const size_t* lens, uint64_t part_size,
                       int K, int s_min, int x_min, int x_max, int flags);
uint64_t orc_contiger_fasta_size(void* r);
void orc_contiger_fasta(void* r, char* dst);
uint64_t orc_contiger_queries(void* r);
void orc_contiger_free(void* r);

#ifdef __cplusplus
}
#endif
#endif
