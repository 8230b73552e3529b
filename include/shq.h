/* include/shq.h -- C ABI of libshq.so: the B200-native replacement for SH-assembly's
 * data-parallel hot path (k-mer counting into a counting quotient filter with
 * deNoise; batched CQF queries; unitig contraction).
 *
 * The reference has no FFI for this path: in-process it is the extern "C" API of
 * cqf/gqf.h driven per key by cqf/CQF_mt.h (class CQF_mt) and
 * src/contig_assembly.cpp.  This header is the batch-granular equivalent a
 * maintainer binds instead; every entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes; every function returns an shq_status
 * (0 = ok) and never calls exit(); the library owns all device memory behind the
 * opaque handle, the caller owns every host buffer.  A handle is single-caller
 * (calls on one handle must not overlap); different handles are independent.
 * There is no CPU fallback: without a CUDA device shq_create fails with
 * SHQ_ERR_CUDA.
 */
#ifndef SHQ_H
#define SHQ_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct shq_handle_s* shq_t;
#define SHQ_NO_KEY (~(uint64_t)0)

typedef enum {
  SHQ_OK = 0,
  SHQ_ERR_ARG = 1,          /* bad argument / unsupported geometry */
  SHQ_ERR_CUDA = 2,         /* CUDA runtime error (no device, OOM, launch failure) */
  SHQ_ERR_CAPACITY = 3,     /* table over-full: a run would pass xnslots (the reference hangs or
                               writes out of bounds here) */
  SHQ_ERR_MALFORMED = 4,    /* FASTQ chunk is not a whole number of 4-line records ending in '\n'
                               (the reference dereferences NULL+1, cqf/CQF_mt.h:723-726) */
  SHQ_ERR_IO = 5,           /* file open/read/write */
  SHQ_ERR_FORMAT = 6,       /* not a .cqf image this build understands (bits_per_slot != 8, ...) */
  SHQ_ERR_COUNT_OVERFLOW = 7, /* a single key was seen more than 2^32-1 times */
  SHQ_ERR_STATE = 8         /* call not valid in the handle's current state */
} shq_status;

const char* shq_strerror(int status);
const char* shq_last_error(shq_t h);      /* detail of the last failure on this handle */

/* ---- lifetime -------------------------------------------------------------
 * replaces qf_init / qf_destroy (cqf/gqf.c:2187-2317) as used by
 * CQF_mt::CQF_mt (cqf/CQF_mt.h:427-445).  nslots = 2^qb, key bits = hb,
 * bits_per_slot = hb-qb must be 8 (always true for the CLIs, CQF-deNoise.cpp:161).
 * shard/nshards: this handle owns quotients [shard, shard+1) * 2^qb / nshards
 * (SURVEY.md 8(e)); pass 0/1 for a whole filter.                                */
int shq_create(int qb, int hb, uint32_t seed, int device, shq_t* out);
int shq_create_shard(int qb, int hb, uint32_t seed, int device, int shard, int nshards, shq_t* out);
void shq_destroy(shq_t h);
/* replaces qf_reset (cqf/gqf.c:2292-2304): empty filter, schedule cleared, device buffers kept. */
int shq_reset(shq_t h);

/* ---- counting -------------------------------------------------------------
 * shq_set_denoise_schedule: the arguments of CQF_mt::build_KmerSpectrum
 * (cqf/CQF_mt.h:959) that drive the scheduler fastq_to_uint64kmers_prod
 * (:821-931): trigger = n_distinct_elts_for_DeNoise, rounds = n_deNoise,
 * end_denoise = --endDeNoise.  Rounds fire at reference-chunk boundaries with
 * t=1 semantics: after the first chunk that leaves ndistinct >= trigger.        */
int shq_set_denoise_schedule(shq_t h, uint64_t trigger_ndistinct, uint32_t rounds, int end_denoise);
/* the scheduler's stderr progress lines "Ready for DeNoise: ..." / "Finished DeNoise: ..."
 * (cqf/CQF_mt.h:867-868, 912-913); ndistinct_true_elts = the -n value they echo.  Off by default. */
int shq_set_progress_log(shq_t h, int enable, uint64_t ndistinct_true_elts);

/* replaces reads_to_kmers (cqf/CQF_mt.h:610-731) for ONE chunk as cut by
 * fastq_read_parts (:735-816).  text is a host buffer (pageable or pinned) and
 * may be reused as soon as the call returns.  Work is batched on the device;
 * counters become exact at shq_sync / shq_counts / shq_finish.                  */
int shq_insert_fastq_chunk(shq_t h, const char* text, size_t n, int K);

/* Zero-copy variant for PINNED (page-locked) host memory: the chunk is copied straight from the caller's buffer
 * with an asynchronous DMA on a separate copy stream, double-buffered against the compute of the previous
 * batch.  The buffer must stay valid and unchanged until shq_sync / shq_finish / shq_counts returns.          */
int shq_insert_fastq_pinned(shq_t h, const char* pinned_text, size_t n, int K);
/* the same for nchunks consecutive chunks of one pinned buffer (chunk c = [chunk_ends[c-1], chunk_ends[c])) */
int shq_insert_fastq_pinned_chunks(shq_t h, const char* pinned_text, size_t n, const uint64_t* chunk_ends, size_t nchunks,
                                   int K);

/* Same for text already resident in device memory (benchmark / pipeline use):
 * d_text holds nchunks consecutive chunks, chunk c = [chunk_ends[c-1], chunk_ends[c])
 * (host array of byte offsets).  Processed before the call returns control
 * to the stream; counters exact after shq_sync.                                 */
int shq_insert_fastq_device(shq_t h, const char* d_text, size_t n, const uint64_t* chunk_ends,
                            size_t nchunks, int K);

/* replaces a loop of qf_insert_advance(qf, key, 0, 1, ...) (cqf/gqf.c:2432-2440):
 * keys are already reduced (hash % range).  Host or device pointer variants.    */
int shq_insert_keys(shq_t h, const uint64_t* keys, size_t n);
int shq_insert_keys_device(shq_t h, const uint64_t* d_keys, size_t n);

/* replaces the tail of build_KmerSpectrum (cqf/CQF_mt.h:979-987): drains pending
 * chunks, runs the --endDeNoise round if scheduled, fixes nelts/ndistinct.       */
int shq_finish(shq_t h);
int shq_sync(shq_t h);
/* waits until every host buffer handed to shq_insert_fastq_pinned so far has been read (its H2D copy is complete);
 * does not wait for the compute                                                                                    */
int shq_sync_copies(shq_t h);

/* replaces runtime->ndistinct_elts / runtime->nelts (cqf/CQF_mt.h:279-280).     */
int shq_counts(shq_t h, uint64_t* ndistinct, uint64_t* nelts);
/* number of deNoise sweeps executed so far by the schedule / by shq_denoise     */
int shq_rounds_done(shq_t h, uint32_t* rounds);

/* replaces one DeNoise mode round (cqf/CQF_mt.h:884-922 -> qf_clean_singleton*,
 * cqf/gqf.c:2792-3040), including the reference's range hand-out.               */
int shq_denoise(shq_t h, uint64_t* removed);

/* ---- sharded / multi-GPU building blocks ---------------------------------
 * Every rank hashes its own reads, partitions the keys by owner shard, exchanges them (NCCL
 * all-to-all, done by the caller) and inserts what it receives with shq_insert_keys_device.
 * shq_hash_fastq_device: all reduced keys of the chunk(s) into a device array; *nkeys = slots written, slots
 *   that the N restart rule leaves unused hold SHQ_NO_KEY (~0) and are skipped by the partition step.
 * shq_partition_keys_device: stable-by-owner grouping; owner = key >> (hb - log2(nparts)).
 * shq_state_view / shq_adopt_state: the sorted distinct (key,count) state, for gathering shards
 * into one filter (export, exact deNoise sweep) and for handing a key range back to a shard.  */
int shq_hash_fastq_device(shq_t h, const char* d_text, size_t n, int K, uint64_t* d_keys, size_t cap, size_t* nkeys);
int shq_partition_keys_device(shq_t h, const uint64_t* d_keys_in, size_t n, int nparts, uint64_t* d_keys_out,
                              uint64_t* counts /* host, nparts */);
int shq_state_view(shq_t h, const uint64_t** d_keys, const uint32_t** d_counts, size_t* n);
int shq_adopt_state(shq_t h, const uint64_t* d_keys, const uint32_t* d_counts, size_t n, uint64_t ndistinct,
                    uint64_t nelts);
int shq_set_counts(shq_t h, uint64_t ndistinct, uint64_t nelts);
/* Two-level exchange (4 bytes per k-mer on the wire instead of 8):
 * shq_mg_bin_fastq_device: hash the chunk(s) and bin the keys over the WHOLE key space into 2^coarse_bits bins
 *   (owner = high log2(world) bits of the bin index, so one owner's bins are one slice).  *d_words = binned
 *   32-bit words (low hb - coarse_bits key bits), *d_bin_off = 2^coarse_bits + 1 offsets; both device pointers
 *   into the handle, valid until its next call.
 * shq_mg_insert_binned: d_recv holds, for source s = 0..world-1 starting at word seg_start[s] (host array), this
 *   shard's coarse bins in order; d_counts[s * cps + c] (device, cps = 2^coarse_bits / world) = words of coarse bin
 *   c from source s.  Splits every coarse bin into 2^fine_bits bins and inserts (aggregate + merge).           */
/* Sharded image (no gather): the slot layout is a max-plus scan, so a shard needs from its predecessors only their
 * composed map (two int64).  shq_mg_layout_aggregate: this shard's composed map in global slot coordinates.
 * shq_mg_layout_tail: lay the shard out behind `carry` (composition of all previous shards' maps) and copy the
 *   entries whose slots reach past the shard's last home slot -- they belong to the NEXT shard's blocks -- into
 *   the caller's device buffers.  shq_mg_materialize_shard: the blocks of this shard's quotient range (the last
 *   shard also owns the xnslots extension), given the previous shard's spill; *d_image points into the handle.
 * shq_header: the 128-byte qfmetadata header for given global counters.                                         */
int shq_mg_layout_aggregate(shq_t h, int64_t agg[2]);
int shq_mg_layout_tail(shq_t h, const int64_t carry[2], size_t cap, uint64_t* d_keys, uint32_t* d_counts, uint64_t* d_pos,
                       size_t* n_tail);
int shq_mg_materialize_shard(shq_t h, const uint64_t* d_spill_keys, const uint32_t* d_spill_counts,
                             const uint64_t* d_spill_pos, size_t n_spill, const uint8_t** d_image, uint64_t* first_block,
                             uint64_t* n_blocks);
int shq_header(shq_t h, uint64_t ndistinct, uint64_t nelts, uint8_t out[128]);
/* Pull exchange over peer memory (one process per GPU on one NVLink/NVSwitch node): no all-to-all of k-mers at all.
 * shq_mg_setup: two fixed buffers of max_words 32-bit words for this rank's binned k-mers; their CUDA IPC handles
 *   (2 x 64 bytes) go to ipc_handles_out.  shq_mg_open_peers: all ranks' handles (world x 2 x 64 bytes, rank-major)
 *   are mapped into this process.
 * shq_mg_bin_local: hash the text ONCE and bin the k-mers over the WHOLE key space (2048 fixed-capacity coarse bins)
 *   into buffer `parity`; d_fine_hist (caller's device buffer, world * 4096 words) receives the LAYOUT of that buffer:
 *   word b = first word of coarse bin b, word 2048 + b = one past its last (a bin that overflows makes this call bin
 *   again with a counting pass and exact offsets; the owners only ever see (begin, end) pairs).
 * shq_mg_insert_pull: d_all_fine_hist = the all-gathered layouts [world][world * 4096].  The owner reads its
 *   coarse bins straight out of every source's buffer `parity` (P2P loads inside the split kernel), splits them into
 *   its fine bins (fixed capacity; exact offsets after a counting pass over the same peer memory if one overflows)
 *   and inserts.  A source may refill a buffer two batches later (the layout all-gather of the batch in between
 *   orders the accesses).                                                                                        */
int shq_mg_setup(shq_t h, size_t max_words, void* ipc_handles_out);
int shq_mg_open_peers(shq_t h, const void* all_handles, int world);
int shq_mg_bin_local(shq_t h, const char* d_text, size_t n, int K, int parity, uint32_t* d_fine_hist, size_t* n_words);
int shq_mg_insert_pull(shq_t h, const uint32_t* d_all_fine_hist, int parity);
/* "The same job at every N": while deNoise rounds are pending a batch may only go in if the GLOBAL number of distinct keys
 * stays below the trigger (cqf/CQF_mt.h:837).  shq_mg_stage_pull = shq_mg_insert_pull up to that decision: this shard's
 * share of the batch is pulled, split, aggregated and its new keys counted (merge into a shadow copy of the state, or
 * binary searches for a batch much smaller than the state); the caller sums n_new over the ranks and then commits or
 * discards on every rank.  A deNoise boundary inside a batch is located by staging ever smaller chunk ranges.
 * shq_mg_denoise_walk / _finalize / _apply: one sharded sweep (cqf/CQF_mt.h:884-922, cqf/gqf.c:2792-3040): the chain of
 * 2^20-slot ranges runs over global slot positions, so rank r walks its entries (laid out behind `carry`, see
 * shq_mg_layout_aggregate) and hands the 16-word chain state to rank r + 1; after the last rank _finalize closes the
 * open range; _apply drops the count-1 entries except the protected ones (protect_last = bit `rank` of state[11]).
 * shq_mg_denoise_positions (optional, before the walk): lays the shard out behind `carry` -- on all ranks at once, so
 * that the chain, which is serial from rank to rank, only runs the walk itself.                                     */
int shq_mg_stage_pull(shq_t h, const uint32_t* d_all_fine_hist, int parity, uint64_t* n_new, uint64_t* n_occ);
int shq_mg_commit(shq_t h);
int shq_mg_discard(shq_t h);
int shq_mg_denoise_positions(shq_t h, const int64_t carry[2]);
int shq_mg_denoise_walk(shq_t h, const int64_t carry[2], uint64_t state[16]);
int shq_mg_denoise_finalize(uint64_t state[16]);
int shq_mg_denoise_apply(shq_t h, int protect_last, uint64_t* removed);
/* Sharded image across processes: shq_mg_image_ipc = CUDA IPC handle (64 bytes) of this rank's block array (after
 * shq_mg_materialize_shard); a rank that opens all of them (shq_mg_open_image_peers, world x 64 bytes rank-major) sees
 * the WHOLE filter: the query and unitig kernels load from (and set traveled bits in) the owners' HBM directly.      */
int shq_mg_image_ipc(shq_t h, void* ipc_handle_out);
int shq_mg_open_image_peers(shq_t h, const void* all_handles, int world);
int shq_mg_close_image_peers(shq_t h);
int shq_mg_bin_fastq_device(shq_t h, const char* d_text, size_t n, int K, int coarse_bits, const uint32_t** d_words,
                            const uint32_t** d_bin_off, size_t* n_words);
int shq_mg_insert_binned(shq_t h, const uint32_t* d_recv, const uint64_t* seg_start, const uint32_t* d_counts, int world,
                         int coarse_bits, int fine_bits, size_t n_words);
/* device-to-device copy on the handle's stream, synchronous (plumbing for callers without a CUDA binding) */
int shq_copy_device(shq_t h, void* d_dst, const void* d_src, size_t bytes);

/* ---- (de)serialisation ------------------------------------------------------
 * replaces qf_serialize / qf_deserialize (cqf/gqf.c:2379-2420) behind
 * CQF_mt::save / load (cqf/CQF_mt.h:514-523).  Byte-compatible both ways.       */
int shq_image_size(shq_t h, size_t* bytes);
int shq_export_cqf(shq_t h, void* dst, size_t cap, size_t* written);
int shq_save(shq_t h, const char* path);
int shq_load(const char* path, int device, shq_t* out);
int shq_load_image(const void* image, size_t n, int device, shq_t* out);
int shq_geometry(shq_t h, int* qb, int* hb, uint64_t* nslots, uint64_t* xnslots, uint64_t* nblocks);

/* replaces qf_iterator/qfi_get/qfi_next (cqf/gqf.c:2474-2601): sorted (key,count)
 * dump.  Pass cap=0 to obtain the entry count in *n.                             */
int shq_iterate(shq_t h, uint64_t* keys, uint64_t* counts, size_t cap, size_t* n);

/* ---- queries ----------------------------------------------------------------
 * replaces qf_count_key_value_set_traveled (cqf/gqf.c:3092-3128) behind
 * CQF_mt::count_key_value_set_traveled (cqf/CQF_mt.h:507-509), batched.
 * was_traveled[i] = traveled bit before the call; duplicates inside one batch
 * behave as if issued in index order.                                           */
int shq_query_set_traveled(shq_t h, const uint64_t* keys, size_t n, uint64_t* counts,
                           uint8_t* was_traveled);
/* replaces qf_count_key_value (cqf/gqf.c:2442-2469): no side effect.             */
int shq_query(shq_t h, const uint64_t* keys, size_t n, uint64_t* counts);

/* ---- hashing (test hook) ------------------------------------------------------
 * The key stream reads_to_kmers would insert for this chunk (NTPC64,
 * base/nthash.hpp:295-309, then % range).  Order is unspecified: compare sorted. */
int shq_hash_fastq_chunk(shq_t h, const char* text, size_t n, int K, uint64_t* keys, size_t cap,
                         size_t* nkeys);

/* ---- unitigs ------------------------------------------------------------------
 * replaces find_unitigs_mt_master ... FASTA writer
 * (src/contig_assembly.cpp:2034-2172, 3028-3218, 935-1084, 603-629): the whole
 * Contiger device path over read text (seed k-mers) and this filter.
 * Options mirror Params (base/Params.h:14-61): K, abundance_min (-s),
 * solid_abundance_min (-x), solid_abundance_max (-X).                           */
typedef struct {
  int K;
  int abundance_min;
  int solid_abundance_min;
  int solid_abundance_max;
} shq_unitig_opts;
typedef struct shq_unitigs_s* shq_unitigs_t;
int shq_unitigs_begin(shq_t h, const shq_unitig_opts* opts, shq_unitigs_t* out);
/* the same over a sharded filter (multi-GPU): h is one rank's shard whose peers' images were opened with
 * shq_mg_open_image_peers; global_ndistinct (all shards) sizes the node table                                      */
int shq_unitigs_begin_sharded(shq_t h, const shq_unitig_opts* opts, uint64_t global_ndistinct, shq_unitigs_t* out);
/* Which k-mer of a read seeds the traversal.  The reference has two consumers of read chunks with different rules: its
 * worker threads take the k-mer at len/2 - K/2 (contig_assembly.cpp:1859-1863; the default here: with -t 16 workers
 * see 16 of every 17 chunks), its master thread the one at len/2 (:2067-2071).  The rules differ only in which small
 * islands get a seed.  Call before the first shq_unitigs_add_reads.                                              */
int shq_unitigs_set_seed_rule(shq_unitigs_t u, int master_rule);
int shq_unitigs_add_reads(shq_unitigs_t u, const char* text, size_t n);   /* one chunk of FASTQ */
int shq_unitigs_build(shq_unitigs_t u);
int shq_unitigs_count(shq_unitigs_t u, size_t* n_unitigs, size_t* total_bases);
int shq_unitigs_write_fasta(shq_unitigs_t u, const char* path);
/* FASTA text into a caller buffer (cap=0: size query) */
int shq_unitigs_fasta(shq_unitigs_t u, char* dst, size_t cap, size_t* written);
void shq_unitigs_free(shq_unitigs_t u);

/* ---- host-side readers and drivers (C++ behind a C ABI) ---------------------------
 * shq_reader_*: seqFile_batch + fastq_read_parts + getFileReader
 * (cqf/CQF_mt.h:334-412, 735-816, 933-957; cqf/reader.h): files are read in 2^23-byte
 * parts cut at a FASTQ record boundary and handed out round-robin over the file list,
 * i.e. the chunk sequence one reference worker thread would see.  fmode: 'f' plain,
 * 'g' gzip, 'b' bzip2 (the reference's -f option).                                  */
typedef struct shq_reader_s* shq_reader_t;
int shq_reader_open(const char* const* paths, int nfiles, char fmode, shq_reader_t* out);
int shq_reader_open_memory(const char* const* bufs, const size_t* lens, int nfiles, uint64_t part_size,
                           shq_reader_t* out);
/* next chunk; *chunk == NULL at the end.  The buffer stays valid until the next call. */
int shq_reader_next(shq_reader_t r, const char** chunk, size_t* n, int* file_index);
/* page-lock the reader's chunk buffers (cudaHostRegister) so that chunks can go to the device with asynchronous copies
 * (shq_insert_fastq_pinned); SHQ_ERR_CUDA if that is not possible -- the staged path still works then.        */
int shq_reader_pin(shq_reader_t r);
void shq_reader_close(shq_reader_t r);
/* CQF_mt::build_KmerSpectrum (cqf/CQF_mt.h:959-995): every chunk of the reader goes into the
 * filter under the schedule set by shq_set_denoise_schedule, then shq_finish.           */
int shq_build_kmer_spectrum(shq_t h, shq_reader_t r, int K);
/* force the device image to be built now (normally lazy: export/save/query)             */
int shq_materialize(shq_t h);

/* parameter block of src/CQF-deNoise.cpp:83-161 (+ mean_CDF2deNoise, cqf/CQF_mt.h:94-133,
 * true2falseKmer_DP, cqf/true2falseKmer_DP.cpp:12-51)                                   */
typedef struct {
  uint64_t num_true_kmers, num_false_kmers;
  int num_deNoise;              /* effective number of mid-stream rounds */
  int lower_bound, upper_bound; /* "#deNoise rounds leading to the same size of CQF" (0 = +oo) */
  uint64_t qb, hb;
  uint64_t trigger;             /* n_distinct_elts_for_DeNoise */
  double fr;
} shq_params;
int shq_param_block(int K, uint64_t n_true_kmers, uint64_t total_kmers, double alpha, const char* error_profile,
                    int deNoise, double fr, shq_params* out);
double shq_cdfpoi_positive(double x, double mean);
double shq_cdfpoi(double x, double mean);

/* ---- instrumentation ----------------------------------------------------------- */
typedef struct {
  uint64_t kernels_launched;   /* kernels launched by this handle so far */
  uint64_t batches;            /* device batches processed */
  uint64_t keys_hashed;        /* k-mer occurrences hashed */
  double ms_hash_count, ms_scatter, ms_aggregate, ms_merge, ms_layout, ms_denoise; /* CUDA-event ms */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t bin_fallbacks;      /* batches whose one-pass binning overflowed a bin and were redone with exact offsets */
  /* work counters per phase, for per-launch algorithmic bytes (bench.py roofline): launches of the phase's dominant
     kernel and the units they processed, summed over the launches */
  uint64_t scatter_launches, scatter_text_bytes, scatter_words;   /* roll_scatter_kernel: text read, words written */
  uint64_t agg_launches, agg_words, agg_distinct;                 /* bin_aggregate_kernel: words read, distinct keys written */
  uint64_t merge_launches, merge_in, merge_out;                   /* merge kernels: entries read (state + run), entries written */
  uint64_t denoise_launches, denoise_entries;                     /* deNoise sweeps: state entries scanned */
  uint64_t layout_launches, layout_entries;                       /* materialise: state entries laid out */
  uint64_t host_syncs;                                            /* device->host scalar round trips */
  uint64_t spec_rejected;      /* speculative merges of armed batches dropped because the deNoise trigger was crossed inside */
  double ms_split;             /* second binning level (split_bins_kernel), CUDA-event ms */
  uint64_t split_launches, split_words;
  uint64_t wide_batches;       /* batches past the armed chunk cap, merged speculatively without chunk ids (and kept) */
  uint64_t wide_fallbacks;     /* ... dropped because the trigger was reached inside: redone in halves */
} shq_stats;
int shq_get_stats(shq_t h, shq_stats* out);
int shq_reset_stats(shq_t h);
void* shq_stream(shq_t h);           /* the cudaStream_t all work of this handle runs on */

#ifdef __cplusplus
}
#endif
#endif /* SHQ_H */
