// shq_core.cu -- handle, batching, deNoise schedule and the C ABI of libshq.so (include/shq.h).
//
// Host-side orchestration of the device pipeline that replaces CQF_mt::build_KmerSpectrum /
// fastq_to_uint64kmers_prod / reads_to_kmers / qf_insert_advance / qf_clean_singleton*
// (reference cqf/CQF_mt.h:610-731, 821-931, 959-1039; cqf/gqf.c).  No CPU fallback: every
// data-path step below is a CUDA kernel from shq_hash.cuh / shq_aggregate.cuh / shq_layout.cuh.
#include <math.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>
#include <chrono>
#include <deque>

#include "shq_aggregate.cuh"
#include "shq_common.cuh"
#include "shq_exchange.cuh"
#include "shq_hash.cuh"
#include "shq_layout.cuh"
#include "shq_roll.cuh"
#include "shq_scan.cuh"

using namespace shq;

struct PhaseTimer;
struct shq_unitigs_s;
extern "C" void shq_unitigs_release(shq_unitigs_s* u);

struct shq_handle_s {
  int device = 0;
  int qb = 0, hb = 0;
  uint32_t seed = 0;
  int shard = 0, nshards = 1;
  uint64_t nslots = 0, xnslots = 0, nblocks = 0;   // whole-filter geometry (gqf.c:2187-2290)
  uint64_t key_lo = 0, key_hi = 0;                 // keys owned by this shard
  cudaStream_t stream = nullptr;
  std::string last_error;

  // resident state: sorted distinct keys + counts
  DevBuf st_keys, st_counts, st_keys2, st_counts2;
  size_t n_state = 0;
  uint64_t ndistinct = 0, nelts = 0;               // runtime counters (CQF_mt.h:279-280)

  // image (materialised lazily); for loaded filters the image is authoritative
  DevBuf image;
  bool image_valid = false;
  bool image_only = false;       // loaded from .cqf and not yet decoded into state
  uint8_t header[128];           // metadata of a loaded image
  bool has_header = false;

  // distinct keys per occurrence seen by the last full aggregation of a batch: sizes the bins (the aggregator's
  // shared-memory table holds kAggMaxDistinct keys); 1/8 until the first batch has been seen
  double distinct_ratio = 0.125;
  bool ratio_seen = false;       // distinct_ratio comes from a batch of this job (not the prior)
  double new_rate = -1.0;        // new keys per occurrence of the last committed run (< 0: none seen yet)

  // deNoise schedule
  uint64_t trigger = 0; uint32_t rounds_left = 0; bool end_denoise = false; uint32_t rounds_done = 0;
  bool finished = false;

  // host staging of chunks
  char* h_stage = nullptr; size_t h_stage_cap = 0, h_stage_fill = 0;
  std::vector<uint64_t> stage_chunk_ends;
  int stage_K = 0;
  size_t max_batch_bytes = (size_t)1 << 28;   // first batches; later ones grow with the state (batch_target_bytes)
  size_t batch_cap_bytes = (size_t)3 << 30;   // ... up to this much text (SHQ_BATCH_CAP_MB)
  int wide_batches = 1;          // SHQ_WIDE: 0 = never batch past the armed chunk cap while rounds are pending, 1 = when the trigger is
                                 // predicted to be far away, 2 = always (tests: exercises the fallback)
  bool walk_pos_ready = false; size_t walk_pos_n = 0; int64_t walk_pos_carry[2] = {0, 0};   // shq_mg_denoise_positions ran for this state
  bool armed_near = true;        // SHQ_ARMED_NEAR=0: armed batches always at full size
  int armed_chunk_cap = 0;       // SHQ_ARMED_CHUNKS: override of the chunks an armed batch may hold (tests)
  double bytes_per_kmer = 3.0;   // text bytes per k-mer, from the last batch
  bool batch_fixed = false;                   // SHQ_MAX_BATCH_MB given: every batch at most that size
  // tunables (environment, read once per handle in init_handle)
  bool spec_bins = true;         // SHQ_SPEC_BINS=0 forces the two-pass (count, then scatter) binning
  bool merge_lookback = true;    // SHQ_MERGE_LOOKBACK=0: count pass + scan + write pass instead of the single-pass merge
  bool merge_tma = false;        // SHQ_MERGE_TMA=1: persistent single-pass merge with inputs staged by bulk async copies
  int bin_bits = 0;              // SHQ_BIN_BITS: 0 = automatic
  // progress lines of the reference scheduler on stderr (cqf/CQF_mt.h:867-868, 912)
  bool log_progress = false; uint64_t ndistinct_true = 0;

  // Pinned pipeline (shq_insert_fastq_pinned): a device RING that the chunks are copied into as they are handed over
  // (copy stream), decoupled from the batches: a batch is whatever run of chunks has ARRIVED when the GPU is free for
  // the next one.  pin_chunks = chunks copied (or being copied) and not yet processed, in order.
  struct PinChunk { size_t off, n; int K; cudaEvent_t ev; };
  cudaStream_t cstream = nullptr;
  DevBuf pin_ring; size_t pin_ring_size = 0;
  size_t pin_head = 0;           // next write offset
  size_t pin_unproc = 0;         // bytes of pin_chunks
  std::deque<PinChunk> pin_chunks;
  std::vector<cudaEvent_t> pin_ev_pool;

  // scratch
  DevBuf d_text, d_nl_tiles, d_nl_pos, d_chunk_ends, d_chunk_line_end, d_chunk_keys, d_chunk_hist, d_slow;
  DevBuf d_hist, d_tot, d_binoff, d_packed, d_tmp_keys, d_tmp_counts, d_tmp_chunk, d_nout, d_occ, d_dense_off;
  DevBuf d_tmp_keys2, d_tmp_counts2, d_run2_keys, d_run2_counts;
  DevBuf d_run_keys, d_run_counts, d_run_chunk, d_split, d_heads, d_heads64, d_scan_scratch, d_misc;
  DevBuf d_pairs, d_pos, d_flags, d_keep, d_q_in, d_q_cnt, d_q_was;
  DevBuf d_sample, d_aggctl;
  DevBuf d_fine2;   // offsets of the second-level bins
  DevBuf d_seg, d_fine, d_fine_cur, d_packed2, d_bases;   // second binning level (single GPU: fine bins; multi-GPU: owner-side split)
  int fine_bits_forced = -1;     // SHQ_FINE_BITS: force the second binning level (0 = off)
  double fine_target = 3500.0;   // SHQ_FINE_TARGET: distinct keys per fine bin the second level aims below
  // pull exchange: two fixed binned-word buffers that peers map through CUDA IPC, and the peers' buffers mapped here
  uint32_t* mg_send[2] = {nullptr, nullptr};
  size_t mg_cap_words = 0;
  std::vector<const uint32_t*> mg_peer;            // [rank * 2 + parity]
  std::vector<void*> mg_opened;                    // pointers obtained from cudaIpcOpenMemHandle
  // sharded image: every rank's block array (own + peers mapped through CUDA IPC), for queries / the unitig build
  std::vector<uint8_t*> img_peer; std::vector<void*> img_opened; DevBuf d_img_peer;
  // staged (not yet committed) exchange batch: its run sits in d_run_*, optionally merged into the shadow state
  struct { bool valid = false, has_shadow = false; size_t run_n = 0; uint64_t run_occ = 0, n_out = 0, n_new = 0; } mg_pending;
  DevBuf d_walk;                 // WalkState of the sharded deNoise range walk
  int* d_err = nullptr;          // device error flag
  void* h_pinned = nullptr;      // small pinned readback area (4 KiB + chunk hist)
  size_t h_pinned_cap = 0;

  void* unitig_cache = nullptr;  // parked shq_unitigs_s of the last unitig build (its device buffers are reused)
  shq_stats stats;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;                 // free timing events
  struct PendingPhase { cudaEvent_t e0, e1; double* acc; };
  std::vector<PendingPhase> ev_pending;             // recorded, not yet read
};

namespace {

// Phase timing without host round trips: a pair of pooled events is recorded around the phase and the
// elapsed time is added to the statistic when the statistics are read (or the pool runs dry).
struct PhaseRec { cudaEvent_t e0, e1; double* acc; };
void resolve_phases(shq_handle_s* h);
struct Phase {
  shq_handle_s* h; cudaEvent_t e0, e1; double* acc;
  Phase(shq_handle_s* h_, double* acc_);
  ~Phase();
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

void resolve_phases(shq_handle_s* h) {
  for (auto& p : h->ev_pending) {
    if (cudaEventSynchronize(p.e1) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) *p.acc += ms;
    }
    h->ev_pool.push_back(p.e0); h->ev_pool.push_back(p.e1);
  }
  h->ev_pending.clear();
}
static cudaEvent_t phase_event(shq_handle_s* h) {
  if (h->ev_pool.empty()) {
    if (h->ev_pending.size() >= 256) resolve_phases(h);
    if (h->ev_pool.empty()) { cudaEvent_t e; SHQ_CUDA(cudaEventCreate(&e)); return e; }
  }
  cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back();
  return e;
}
Phase::Phase(shq_handle_s* h_, double* acc_) : h(h_), acc(acc_) {
  e0 = phase_event(h); e1 = phase_event(h);
  cudaEventRecord(e0, h->stream);
}
Phase::~Phase() {
  cudaEventRecord(e1, h->stream);
  h->ev_pending.push_back({e0, e1, acc});
}

void check_err_flag(shq_handle_s* h) {
  int e = 0;
  SHQ_CUDA(cudaMemcpyAsync(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  SHQ_CUDA(cudaStreamSynchronize(h->stream));
  h->stats.d2h_bytes += 4; h->stats.host_syncs++;
  if (e) {
    int z = 0;
    SHQ_CUDA(cudaMemcpyAsync(h->d_err, &z, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    SHQ_FAIL(e, "device reported: %s", shq_strerror(e));
  }
}

void* pinned(shq_handle_s* h, size_t bytes) {
  if (bytes > h->h_pinned_cap) {
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned = nullptr;
    SHQ_CUDA(cudaMallocHost(&h->h_pinned, bytes + 4096));
    h->h_pinned_cap = bytes + 4096;
  }
  return h->h_pinned;
}

template <class T>
T read_scalar(shq_handle_s* h, const T* d) {
  T* p = (T*)pinned(h, sizeof(T));
  SHQ_CUDA(cudaMemcpyAsync(p, d, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  SHQ_CUDA(cudaStreamSynchronize(h->stream));
  h->stats.d2h_bytes += sizeof(T); h->stats.host_syncs++;
  return *p;
}

// two device words (possibly in different buffers) with one host round trip
template <class T, class U>
void read_pair(shq_handle_s* h, const T* da, T* a, const U* db, U* b) {
  char* p = (char*)pinned(h, 32);
  SHQ_CUDA(cudaMemcpyAsync(p, da, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  SHQ_CUDA(cudaMemcpyAsync(p + 16, db, sizeof(U), cudaMemcpyDeviceToHost, h->stream));
  SHQ_CUDA(cudaStreamSynchronize(h->stream));
  h->stats.d2h_bytes += sizeof(T) + sizeof(U); h->stats.host_syncs++;
  memcpy(a, p, sizeof(T)); memcpy(b, p + 16, sizeof(U));
}

// ---------------------------------------------------------------------------
// A "run" = sorted distinct (key, count[, first chunk]) produced from one batch of binned keys.
struct BatchBins {
  int n_bins = 0, lb = 0;
  uint32_t n_keys = 0;          // occurrences scattered (an upper bound when spec)
  bool armed = false;
  int nchunks = 0;
  bool spec = false;            // fixed-capacity bins, no counting pass: bin p = [p * capbin, cursor[p])
  mutable bool spec_checked = false;
  uint32_t span = 0;            // words spanned by the binned array when spec
  // the binned words, the bin offsets (n_bins + 1) and, for fixed-capacity bins, the end of every bin
  const uint32_t* words = nullptr; const uint32_t* off = nullptr; const uint32_t* end = nullptr;
  int fine_bits = 0;            // second binning level applied (n_bins and lb already include it)
  bool level1_spec = false;     // the first level used fixed-capacity bins too
  // from the sample (coarse bin 0 aggregated on its own): distinct keys / new keys per occurrence of THIS batch; < 0: none
  double ratio = -1.0, new_frac = -1.0; uint32_t sample_new = 0;
};
struct BinOverflow {};          // a fixed-capacity bin overflowed: redo the batch with exact offsets

struct Run { DevBuf* keys; DevBuf* counts; DevBuf* chunk; size_t n = 0; uint64_t occ = 0; };

int pick_bin_bits(shq_handle_s* h, uint64_t est_keys, int chunk_bits_needed) {
  // The hashing pass is the faster the fewer bins it scatters into (per round of 12 288 keys it pays one reservation and
  // three shared-memory words per bin: 83 -> 66 ms per step from 4096 to 2048 bins on configs[2]); the aggregator wants at
  // most ~3000 distinct keys per bin, which the second binning level provides (up to 128 fine bins per coarse bin).  So:
  // the fewest coarse bins (2^10 .. 2^12) that need no second level, and if one is needed anyway, the fewest that let it
  // reach the target.  Distinct keys are estimated with the ratio of the previous batch (+30 %).
  // The chunk id of an armed batch shares the 32-bit word with the key bits below the bin: that may ask for more bins.
  int shard_bits = 0; while ((1 << shard_bits) < h->nshards) shard_bits++;
  int local_bits = h->hb - shard_bits;
  const double est_d = (double)est_keys * (h->ratio_seen ? std::min(1.0, h->distinct_ratio * 1.3) : 1.0);
  int pb = -1;
  if (!h->ratio_seen) pb = 12;          // first batch of a job: nothing known, as many bins as the hashing pass takes
  for (int b = 10; b <= 12 && pb < 0; b++) if (est_d / (double)(1 << b) <= h->fine_target) pb = b;
  if (pb < 0) { pb = 10; while (pb < 12 && est_d / (double)((uint64_t)1 << (pb + 7)) > 3000.0) pb++; }
  while (pb > 8 && (est_keys >> pb) < 2048) pb--;          // small batches: at least ~2000 words per bin
  if (h->bin_bits) pb = h->bin_bits;
  if (pb > local_bits) pb = local_bits;
  // the low key bits share a 32-bit word with the chunk id, and 0xFFFFFFFF is the aggregator's empty marker
  while ((local_bits - pb + chunk_bits_needed > 32 || local_bits - pb > 31) && pb < 13) pb++;
  if (local_bits - pb + chunk_bits_needed > 32 || local_bits - pb > 31) SHQ_FAIL(SHQ_ERR_ARG, "hb=%d too wide for 32-bit packed keys", h->hb);
  return pb;
}

// How much text the next device batch should hold.  Every batch pays one merge over the whole state, so batches grow
// with the state (fewer, larger merges); a batch that contains a deNoise boundary is aggregated a second time, so they
// start small; and the distinct keys of a batch must fit the aggregator's per-bin tables (~6500 keys x 4096 bins).
size_t batch_target_bytes(shq_handle_s* h) {
  if (h->batch_fixed) return h->max_batch_bytes;
  // (a small state merges in no time: E. coli-size jobs keep 256 MiB batches, whose bins need no second level)
  double t = (double)h->max_batch_bytes * (h->n_state < 25000000 ? 1.0 : (double)h->n_state / 8.0e6);
  t = std::min(t, (double)h->batch_cap_bytes);
  return (size_t)t;
}
// Chunks per batch while rounds are pending: the chunk id shares the 32-bit packed word with the low key bits
size_t batch_chunk_cap(shq_handle_s* h) {
  if (h->rounds_left == 0) return (size_t)1 << 20;
  if (h->armed_chunk_cap > 0) return (size_t)h->armed_chunk_cap;
  int shard_bits = 0; while ((1 << shard_bits) < h->nshards) shard_bits++;
  const int local_bits = h->hb - shard_bits;
  int bits = 32 - (local_bits - 12);                // with 4096 bins
  if (bits < 3) bits = 32 - (local_bits - 13);      // 8192 bins
  if (bits < 0) bits = 0;
  if (bits > 10) bits = 10;                         // histogram arrays of the merge / aggregation kernels
  return (size_t)1 << bits;
}
// Text bytes and chunks the next batch may hold.  While rounds are pending a batch normally keeps the chunk id of every
// word (<= batch_chunk_cap chunks, <= 1 GiB), so that a deNoise boundary inside it can be located.  When the new-key rate
// of the previous batch says that the trigger is still far away, the batch is made WIDE instead: as much text as is
// predicted to stay below the trigger (up to batch_cap_bytes), binned without chunk ids and merged speculatively; should
// the trigger be reached after all, process_batch() drops the shadow and redoes the batch in halves.  Fewer, larger
// batches mean fewer passes over the resident state.
constexpr size_t kArmedBatchBytes = (size_t)1 << 30;
constexpr size_t kNearBatch = (size_t)256 << 20;
void batch_limits(shq_handle_s* h, size_t* bytes, size_t* chunks) {
  size_t t = batch_target_bytes(h), c = batch_chunk_cap(h);
  if (h->rounds_left > 0) {
    const size_t armed_bytes = std::min(t, kArmedBatchBytes);
    bool wide = false;
    double room_bytes = -1.0;                               // text until the predicted crossing (< 0: no prediction)
    if (h->new_rate >= 0 && h->n_state > 0 && h->ndistinct < h->trigger) {
      const double room = (double)(h->trigger - h->ndistinct) / std::max(h->new_rate * 1.3, 1e-9);   // k-mers, 30 % margin
      room_bytes = room * h->bytes_per_kmer;
    }
    if (h->wide_batches == 2) { c = (size_t)1 << 20; wide = true; }
    else if (h->wide_batches && room_bytes >= 0) {
      if (h->armed_chunk_cap > 0 || (t > armed_bytes && room_bytes > 1.5 * (double)armed_bytes)) {
        t = (size_t)std::min((double)t, room_bytes); c = (size_t)1 << 20; wide = true;
      }
    }
    if (!wide) {
      t = armed_bytes;
      // The batch that holds the boundary is aggregated twice more (first chunks, then both sides of c*) and merged three
      // times: keep it small.  A batch that is predicted to stay below the trigger ends short of the crossing (it commits
      // as a whole); once the crossing is nearer than kNearBatch the batches are that small.
      if (h->armed_near && room_bytes >= 0 && !h->batch_fixed && h->armed_chunk_cap == 0)
        t = std::min(armed_bytes, (size_t)std::max(room_bytes, (double)kNearBatch));
    }
  }
  *bytes = t; *chunks = c;
}

// Capacity of a fixed-size bin that is to receive `mean` words on average.  ntHash keys are uniform, but the c
// occurrences of one k-mer inside the batch all land in the same bin, so the spread is sqrt(c * mean), not sqrt(mean);
// c is taken from the distinct ratio of the previous batch (8 while nothing is known: the first batch is small).
uint64_t bin_capacity(shq_handle_s* h, uint64_t mean) {
  const double c = h->ratio_seen ? 1.0 / std::max(h->distinct_ratio, 1e-3) : 8.0;
  const double slack = std::max((double)mean / 16.0, 6.0 * sqrt((double)mean * c)) + 1024.0;
  return ((uint64_t)((double)mean + slack) + 7) & ~7ULL;
}

// K1: newline index of a batch; returns the number of lines.  d_text may start at any byte: the vector loads start at
// the 16-byte boundary below it and the bytes in front of the batch are skipped (no staging copy of unaligned batches).
uint32_t index_lines(shq_handle_s* h, const char* d_text, size_t n_text) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  const uint32_t skip = (uint32_t)(reinterpret_cast<uintptr_t>(d_text) & 15);
  const char* base = d_text - skip;
  const size_t n_al = n_text + skip;
  unsigned ntiles = cdiv(n_al, kNlTile);
  h->d_nl_tiles.reserve((size_t)(ntiles + 1) * 4);
  nl_count_kernel<<<ntiles, kNlThreads, 0, st>>>(base, n_al, skip, h->d_nl_tiles.as<uint32_t>()); L++;
  h->d_misc.reserve(256);
  exclusive_scan_inplace<SumU32>(h->d_nl_tiles.as<uint32_t>(), ntiles, h->d_scan_scratch, st, h->d_misc.as<uint32_t>(), &L);
  uint32_t n_lines = read_scalar(h, h->d_misc.as<uint32_t>());
  h->d_nl_pos.reserve((size_t)(n_lines + 8) * 4);
  nl_write_kernel<<<ntiles, kNlThreads, 0, st>>>(base, n_al, skip, h->d_nl_tiles.as<uint32_t>(), h->d_nl_pos.as<uint32_t>()); L++;
  return n_lines;
}

// A 1/n_bins sample of the batch, for free of any assumption about the read set: coarse bin 0 (ntHash keys are uniform,
// so it holds a uniform sample of the batch's k-mers, from all chunks) is aggregated on its own by one CTA, and its
// distinct keys are looked up in the state.  Gives, before the batch is processed, its distinct-key ratio (-> second
// binning level, insert strategy, bin capacities) and its new-key rate (-> is the deNoise trigger crossed inside?).
__global__ void count_new_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ n_keys, const uint64_t* __restrict__ state, size_t n_state,
                                 uint32_t* __restrict__ n_new) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool is_new = false;
  if (i < *n_keys) {
    const uint64_t k = keys[i];
    size_t lo = 0, hi = n_state;
    while (lo < hi) { const size_t mid = lo + ((hi - lo) >> 1); if (state[mid] < k) lo = mid + 1; else hi = mid; }
    is_new = !(lo < n_state && state[lo] == k);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, is_new);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_new, (uint32_t)__popc(m));
}
void sample_batch(shq_handle_s* h, BatchBins* bb, uint32_t capbin) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  h->d_tmp_keys.reserve(((size_t)capbin + 8) * 8); h->d_tmp_counts.reserve(((size_t)capbin + 8) * 4);
  h->d_sample.reserve(64);
  uint32_t* d_s = h->d_sample.as<uint32_t>();          // [0] distinct, [1] occurrences, [2] new, [3] split overflow (unused)
  SHQ_CUDA(cudaMemsetAsync(d_s, 0, 16, st));
  AggOut o{};
  o.keys = h->d_tmp_keys.as<uint64_t>(); o.counts = h->d_tmp_counts.as<uint32_t>(); o.chunk = nullptr;
  o.n_out = d_s; o.occ_out = d_s + 1; o.n_out2 = d_s + 3; o.occ_out2 = d_s + 3; o.split_overflow = reinterpret_cast<int*>(d_s + 3);
  o.fixed_bits = 0; o.sub_stride = 0; o.bin_end = bb->end; o.direct = 1; o.first_hist = nullptr;
  // sample by key range: about 6000 words of the bin, so that even an all-distinct sample fits the table
  { const uint32_t mean = bb->n_keys / (uint32_t)bb->n_bins; int sb = 0; while (sb < bb->lb - 1 && (mean >> sb) > 6000u) sb++; o.sample_bits = sb; }
  const size_t smem = agg_smem_bytes(false, false);
  SHQ_CUDA(cudaFuncSetAttribute(bin_aggregate_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bin_aggregate_kernel<false, false><<<1, kAggThreads, smem, st>>>(bb->words, bb->off, bb->lb, h->key_lo, 0u, 0xFFFFFFFFu, 0u, o); L++;
  count_new_kernel<<<cdiv(capbin, 256), 256, 0, st>>>(o.keys, d_s, h->st_keys.as<uint64_t>(), h->n_state, d_s + 2); L++;
  SHQ_CUDA(cudaGetLastError());
  uint32_t* p = (uint32_t*)pinned(h, 16);
  SHQ_CUDA(cudaMemcpyAsync(p, d_s, 12, cudaMemcpyDeviceToHost, st));
  SHQ_CUDA(cudaStreamSynchronize(st));
  h->stats.d2h_bytes += 12; h->stats.host_syncs++;
  if (p[1] >= 256) {
    bb->ratio = (double)p[0] / (double)p[1]; bb->new_frac = (double)p[2] / (double)p[1]; bb->sample_new = p[2];
    h->distinct_ratio = bb->ratio; h->ratio_seen = true;
  }
}

// After the first binning level: sample the batch (coarse bin 0) and, when a coarse bin would hold more distinct keys
// than the aggregator's table takes -- large batches of a large genome, where almost every k-mer of a batch is
// distinct -- split every coarse bin into 2^fb fine bins (split_bins_kernel).  bb describes the coarse bins on entry
// and the bins to aggregate on return.  max_bin_words bounds the words of one coarse bin (sizes the sample's output).
void refine_bins(shq_handle_s* h, BatchBins* bb, uint32_t max_bin_words) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  sample_batch(h, bb, max_bin_words);
  const int lb = bb->lb, n_bins = bb->n_bins;
  const uint64_t n_ub = bb->n_keys;
  const double ratio = bb->ratio >= 0 ? std::min(1.0, bb->ratio * 1.1) : 1.0;
  const double est_d = (double)n_ub / (double)n_bins * ratio;
  int fb = 0;
  while (fb < 7 && lb - fb > 1 && est_d / (double)(1 << fb) > h->fine_target) fb++;      // load factor of the aggregator's table <= ~0.45
  if (h->fine_bits_forced >= 0) fb = std::min(h->fine_bits_forced, std::min(7, lb - 1));
  const uint64_t n_fine = (uint64_t)n_bins << fb;
  const uint64_t capfine = bin_capacity(h, n_ub / n_fine);
  const uint64_t span2 = capfine * n_fine;
  if (fb == 0 || span2 + kSplitTile + 64 >= ((uint64_t)1 << 32)) return;
  Phase ph2(h, &h->stats.ms_split);
  h->d_packed2.reserve((size_t)(span2 + kSplitTile + 64) * 4);
  h->d_fine2.reserve((size_t)(n_fine + 2) * 4); h->d_fine_cur.reserve((size_t)(n_fine + 2) * 4);
  h->d_misc.reserve(256);
  int* d_ovf2 = h->d_misc.as<int>() + 24;
  SHQ_CUDA(cudaMemsetAsync(d_ovf2, 0, 4, st));
  fine_bin_offsets_kernel<<<cdiv(n_fine + 1, 256), 256, 0, st>>>(h->d_fine2.as<uint32_t>(), h->d_fine_cur.as<uint32_t>(), (uint32_t)n_fine, (uint32_t)capfine); L++;
  int parts = 1; while (n_bins * parts < 8 * kNumSMs) parts *= 2;
  // exact-offset bins (no cursor array): bin c ends where bin c + 1 starts
  const uint32_t* bin_end = bb->end ? bb->end : bb->off + 1;
  split_bins_kernel<<<n_bins * parts, kSplitThreads, 0, st>>>(bb->words, bb->off, bin_end, parts, lb, fb, h->d_fine_cur.as<uint32_t>(), (uint32_t)capfine,
                                                            (uint32_t)span2, d_ovf2, h->d_packed2.as<uint32_t>()); L++;
  SHQ_CUDA(cudaGetLastError());
  h->stats.split_launches++; h->stats.split_words += n_ub;
  bb->n_bins = (int)n_fine; bb->lb = lb - fb; bb->span = (uint32_t)span2; bb->fine_bits = fb; bb->spec = true;
  bb->words = h->d_packed2.as<uint32_t>(); bb->off = h->d_fine2.as<uint32_t>(); bb->end = h->d_fine_cur.as<uint32_t>();
}

// Both hashing passes over the batch text (or a raw key array): count -> offsets -> scatter.
// On return d_packed holds the binned words and d_binoff the (n_bins+1) bin offsets.
// global_bits > 0: bin over the WHOLE key space [0, 2^hb) into 2^global_bits bins, whatever this handle's shard
// is (first level of the multi-GPU exchange: bins of one owner are contiguous).
void bin_batch(shq_handle_s* h, const char* d_text, size_t n_text, const uint64_t* h_chunk_ends, int nchunks, int K,
               const uint64_t* d_keys, size_t n_keys_in, bool armed, BatchBins* bb, int global_bits = 0, bool allow_spec = false) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  HashParams hp{};
  const bool from_text = d_keys == nullptr;
  if (from_text) {
    if (K < 1 || K > 1024) SHQ_FAIL(SHQ_ERR_ARG, "K=%d outside [1,1024]", K);
    if (n_text >= ((size_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "batch text of %zu bytes exceeds 4 GiB", n_text);
    uint32_t n_lines = index_lines(h, d_text, n_text);
    h->d_chunk_ends.reserve((size_t)nchunks * 8 + 8);
    h->d_chunk_line_end.reserve((size_t)nchunks * 4 + 4);
    SHQ_CUDA(cudaMemcpyAsync(h->d_chunk_ends.p, h_chunk_ends, (size_t)nchunks * 8, cudaMemcpyHostToDevice, st));
    h->stats.h2d_bytes += (size_t)nchunks * 8;
    chunk_validate_kernel<<<cdiv(nchunks, 128), 128, 0, st>>>(d_text, h->d_nl_pos.as<uint32_t>(), n_lines, h->d_chunk_ends.as<uint64_t>(),
                                                            nchunks, h->d_chunk_line_end.as<uint32_t>(), h->d_err); L++;
    // the device error flag is read together with the key total below (the kernels in between only index
    // nl_pos[0 .. n_lines), whatever the text holds)
    if (n_lines & 3) SHQ_FAIL(SHQ_ERR_MALFORMED, "batch holds %u lines, not a multiple of 4", n_lines);
    hp.text = d_text; hp.nl_pos = h->d_nl_pos.as<uint32_t>(); hp.n_reads = n_lines / 4; hp.K = K;
  }
  int chunk_bits = 0;
  if (armed) { while ((1 << chunk_bits) < nchunks) chunk_bits++; if (chunk_bits > 12) SHQ_FAIL(SHQ_ERR_ARG, "too many chunks in one batch"); }
  uint64_t est = from_text ? (uint64_t)(n_text / 3) : (uint64_t)n_keys_in;
  int pb = global_bits ? global_bits : pick_bin_bits(h, est / (uint64_t)h->nshards + 1, chunk_bits);
  int shard_bits = 0; while ((1 << shard_bits) < h->nshards) shard_bits++;
  if (global_bits) shard_bits = 0;
  int lb = h->hb - shard_bits - pb;
  if (lb < 1 || lb > 31) SHQ_FAIL(SHQ_ERR_ARG, "bin geometry: %d low key bits do not fit the packed word", lb);
  int n_bins = 1 << pb;
  hp.hb = h->hb; hp.lb = lb; hp.key_lo = global_bits ? 0 : h->key_lo; hp.key_hi = global_bits ? (1ULL << h->hb) : h->key_hi; hp.n_bins = n_bins;
  hp.nchunks = nchunks;
  hp.chunk_line_end = (armed && from_text) ? h->d_chunk_line_end.as<uint32_t>() : nullptr;
  h->d_hist.reserve((size_t)(n_bins + 1) * 4);
  h->d_binoff.reserve((size_t)(n_bins + 2) * 4);
  h->d_tot.reserve((size_t)(n_bins + 2) * 4);      // scatter cursors
  h->d_chunk_keys.reserve((size_t)(nchunks + 1) * 4);
  hp.chunk_keys = nullptr;   // per-chunk key counts are no longer needed: nelts comes from the aggregator's occurrence totals
  SHQ_CUDA(cudaMemsetAsync(h->d_hist.p, 0, (size_t)(n_bins + 1) * 4, st));
  SHQ_CUDA(cudaMemsetAsync(h->d_chunk_keys.p, 0, (size_t)(nchunks + 1) * 4, st));
  const int G = kNumSMs * 2;
  const size_t smem_cle = hp.chunk_line_end ? (size_t)nchunks * 4 + 16 : 0;
  size_t smem_count = sizeof(RollTables) + (size_t)n_bins * 4 + smem_cle;
  const bool whole = h->nshards == 1 || global_bits != 0;   // every key of the hb-bit space is binned
  const bool wide = n_bins <= 4096;          // 16 keys per thread and round fit beside the bin arrays
  size_t smem_scatter = sizeof(RollTables) + ((size_t)3 * n_bins + 1) * 4 + (size_t)kScatThreads * (wide ? 16 : 8) * 14 + 16 + smem_cle;
  size_t smem_keys = (size_t)n_bins * 8;
  Sink sk{};
  if (allow_spec && h->spec_bins && from_text && whole && wide && !global_bits) {
    // One hashing pass.  ntHash keys are uniform, so every bin gets the mean + 6 % + 2048 words; a bin that would not
    // fit is caught by the scatter kernel and the caller falls back to the exact two-pass binning for this batch.
    h->d_misc.reserve(256);
    unsigned long long* d_ub = h->d_misc.as<unsigned long long>() + 12;
    SHQ_CUDA(cudaMemsetAsync(d_ub, 0, 8, st));
    window_total_kernel<<<kNumSMs * 4, 256, 0, st>>>(hp.nl_pos, hp.n_reads, K, d_ub); L++;
    unsigned long long n_ub = 0; int dev_err = 0;
    read_pair(h, d_ub, &n_ub, h->d_err, &dev_err);
    if (dev_err) { SHQ_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), st)); SHQ_FAIL(dev_err, "device reported: %s", shq_strerror(dev_err)); }
    const uint64_t capbin = bin_capacity(h, n_ub / (uint64_t)n_bins);
    const uint64_t span = capbin * (uint64_t)n_bins;
    const uint64_t round = (uint64_t)kScatThreads * 16;
    if (n_ub > 0 && span + round + 64 < ((uint64_t)1 << 32)) {
      h->d_packed.reserve((size_t)(span + round + 64) * 4);
      int* d_ovf = h->d_misc.as<int>() + 20;
      SHQ_CUDA(cudaMemsetAsync(d_ovf, 0, 4, st));
      fixed_bin_offsets_kernel<<<cdiv(n_bins + 1, 256), 256, 0, st>>>(h->d_binoff.as<uint32_t>(), h->d_tot.as<uint32_t>(), n_bins, (uint32_t)capbin); L++;
      {
        Phase ph(h, &h->stats.ms_scatter);
        SHQ_CUDA(cudaFuncSetAttribute(roll_scatter_kernel<kScatThreads, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
        roll_scatter_kernel<kScatThreads, 16, true><<<kNumSMs, kScatThreads, smem_scatter, st>>>(hp, h->d_tot.as<uint32_t>(), h->d_packed.as<uint32_t>(),
                                                                                             (uint32_t)capbin, (uint32_t)span, d_ovf);
        L++;
      }
      SHQ_CUDA(cudaGetLastError());
      bb->n_bins = n_bins; bb->lb = lb; bb->n_keys = (uint32_t)n_ub; bb->armed = armed; bb->nchunks = nchunks;
      bb->spec = true; bb->span = (uint32_t)span; bb->level1_spec = true;
      bb->words = h->d_packed.as<uint32_t>(); bb->off = h->d_binoff.as<uint32_t>(); bb->end = h->d_tot.as<uint32_t>(); bb->fine_bits = 0;
      h->stats.keys_hashed += n_ub;
      h->stats.scatter_launches++; h->stats.scatter_text_bytes += n_text; h->stats.scatter_words += n_ub;
      refine_bins(h, bb, (uint32_t)capbin);
      return;
    }
  }
  {
    Phase ph(h, &h->stats.ms_hash_count);
    if (from_text) {
      if (whole) {
        SHQ_CUDA(cudaFuncSetAttribute(roll_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
        roll_count_kernel<true><<<G, kRollThreads, smem_count, st>>>(hp, h->d_hist.as<uint32_t>());
      } else {
        SHQ_CUDA(cudaFuncSetAttribute(roll_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
        roll_count_kernel<false><<<G, kRollThreads, smem_count, st>>>(hp, h->d_hist.as<uint32_t>());
      }
      L++;
    } else {
      SHQ_CUDA(cudaFuncSetAttribute(hash_keys_kernel<kSinkCount>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_keys));
      hash_keys_kernel<kSinkCount><<<G, 256, smem_keys, st>>>(d_keys, n_keys_in, hp, sk, h->d_hist.as<uint32_t>(), nullptr); L++;
    }
    SHQ_CUDA(cudaMemcpyAsync(h->d_binoff.p, h->d_hist.p, (size_t)n_bins * 4, cudaMemcpyDeviceToDevice, st));
    exclusive_scan_inplace<SumU32>(h->d_binoff.as<uint32_t>(), n_bins, h->d_scan_scratch, st, h->d_binoff.as<uint32_t>() + n_bins, &L);
    SHQ_CUDA(cudaMemcpyAsync(h->d_tot.p, h->d_binoff.p, (size_t)n_bins * 4, cudaMemcpyDeviceToDevice, st));
  }
  uint32_t n_keys = 0; int dev_err = 0;
  read_pair(h, h->d_binoff.as<uint32_t>() + n_bins, &n_keys, h->d_err, &dev_err);
  if (dev_err) {
    SHQ_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), st));
    SHQ_FAIL(dev_err, "device reported: %s", shq_strerror(dev_err));
  }
  h->d_packed.reserve((size_t)(n_keys + 4) * 4);
  {
    Phase ph(h, &h->stats.ms_scatter);
    if (from_text) {
#define SHQ_SCATTER(WW, WH) do { \
        SHQ_CUDA(cudaFuncSetAttribute(roll_scatter_kernel<kScatThreads, WW, WH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter)); \
        roll_scatter_kernel<kScatThreads, WW, WH><<<kNumSMs, kScatThreads, smem_scatter, st>>>(hp, h->d_tot.as<uint32_t>(), h->d_packed.as<uint32_t>()); } while (0)
      if (wide) { if (whole) SHQ_SCATTER(16, true); else SHQ_SCATTER(16, false); }
      else { if (whole) SHQ_SCATTER(8, true); else SHQ_SCATTER(8, false); }
#undef SHQ_SCATTER
      L++;
    } else {
      sk.packed = h->d_packed.as<uint32_t>();
      SHQ_CUDA(cudaFuncSetAttribute(hash_keys_kernel<kSinkScatter>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_keys));
      hash_keys_kernel<kSinkScatter><<<G, 256, smem_keys, st>>>(d_keys, n_keys_in, hp, sk, h->d_tot.as<uint32_t>(), nullptr); L++;
    }
  }
  SHQ_CUDA(cudaGetLastError());
  bb->n_bins = n_bins; bb->lb = lb; bb->n_keys = n_keys; bb->armed = armed; bb->nchunks = nchunks;
  bb->spec = false; bb->span = 0;
  bb->words = h->d_packed.as<uint32_t>(); bb->off = h->d_binoff.as<uint32_t>(); bb->end = nullptr; bb->fine_bits = 0;
  h->stats.keys_hashed += n_keys;
  h->stats.scatter_launches++; h->stats.scatter_text_bytes += from_text ? n_text : n_keys_in * 8; h->stats.scatter_words += n_keys;
}

// Aggregate the binned batch (chunks [clo, chi]) into dense sorted run(s).
//  split: occurrences with chunk <= cstar go to run A, the rest to run B (one pass).
//  first_hist (host vector, may be null): histogram of the first chunk of every distinct key.
// Returns false if the split counters overflowed (caller falls back to two plain passes).
bool aggregate_bins(shq_handle_s* h, const BatchBins& bb, uint32_t clo, uint32_t chi, bool split, uint32_t cstar, Run* A, Run* B,
                    std::vector<uint32_t>* first_hist, bool track = false) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  Phase ph(h, &h->stats.ms_aggregate);
  size_t cap = (size_t)(bb.spec ? bb.span : bb.n_keys) + 4;
  int fixed_bits = 0;               // CTAs per bin: keep each CTA's share of the bin's distinct keys inside its table
  { // a CTA's table holds kAggMaxDistinct keys: split a bin over 2^fixed_bits CTAs (key sub-ranges) when the expected
    // number of distinct keys -- occurrences x the ratio seen in the previous batch -- would not fit
    const double per_bin = (double)bb.n_keys / (double)bb.n_bins * h->distinct_ratio * 1.2;
    while (fixed_bits < 4 && fixed_bits < bb.lb && per_bin / (double)(1 << fixed_bits) > 0.9 * kAggMaxDistinct) fixed_bits++; }
  const size_t S = (size_t)1 << fixed_bits;
  h->d_tmp_keys.reserve(cap * 8 * S); h->d_tmp_counts.reserve(cap * 4 * S);
  const bool chunks = track || first_hist != nullptr;     // first chunk of every key wanted
  if (chunks) h->d_tmp_chunk.reserve(cap * 2 * S);
  if (split) { h->d_tmp_keys2.reserve(cap * 8 * S); h->d_tmp_counts2.reserve(cap * 4 * S); }
  const size_t nb = (size_t)bb.n_bins * S;          // (bin, sub-range) groups
  h->d_nout.reserve((nb + 1) * 4 * 4);           // n_out, n_out2, occ, occ2
  uint32_t* n_out = h->d_nout.as<uint32_t>(); uint32_t* n_out2 = n_out + nb + 1; uint32_t* occ = n_out2 + nb + 1; uint32_t* occ2 = occ + nb + 1;
  h->d_dense_off.reserve((nb + 2) * 4 * 2);
  uint32_t* doff = h->d_dense_off.as<uint32_t>(); uint32_t* doff2 = doff + nb + 2;
  h->d_misc.reserve(256);
  int* d_ovf = h->d_misc.as<int>() + 16;
  SHQ_CUDA(cudaMemsetAsync(d_ovf, 0, 4, st));
  uint32_t* d_fh = nullptr;
  if (first_hist) {
    if (bb.nchunks > kAggBuckets) SHQ_FAIL(SHQ_ERR_ARG, "too many chunks in one batch (max %d)", kAggBuckets);
    h->d_chunk_hist.reserve((size_t)(kMergeHistMax + kAggBuckets) * 4);
    d_fh = h->d_chunk_hist.as<uint32_t>();
    SHQ_CUDA(cudaMemsetAsync(d_fh, 0, (size_t)kAggBuckets * 4, st));
  }
  AggOut o{};
  o.keys = h->d_tmp_keys.as<uint64_t>(); o.counts = h->d_tmp_counts.as<uint32_t>(); o.chunk = chunks ? h->d_tmp_chunk.as<uint16_t>() : nullptr;
  o.n_out = n_out; o.keys2 = split ? h->d_tmp_keys2.as<uint64_t>() : nullptr; o.counts2 = split ? h->d_tmp_counts2.as<uint32_t>() : nullptr;
  o.n_out2 = n_out2; o.occ_out = occ; o.occ_out2 = occ2; o.first_hist = d_fh; o.split_overflow = d_ovf;
  o.fixed_bits = fixed_bits; o.sub_stride = cap;
  o.bin_end = bb.end;
  // dense output: the runs are written in order straight away (look-back over the groups' distinct counts): no per-bin
  // staging of the whole batch, no compaction pass
  const uint64_t bound = (uint64_t)bb.n_keys + 8;           // distinct keys <= occurrences
  A->keys->reserve(bound * 8); A->counts->reserve(bound * 4);
  if (chunks) A->chunk->reserve(bound * 2);
  if (split) { B->keys->reserve(bound * 8); B->counts->reserve(bound * 4); }
  h->d_aggctl.reserve((nb + 16) * 8 + 64);
  unsigned long long* d_ctl = h->d_aggctl.as<unsigned long long>();       // [0..3] totals, [4] ticket, [8 ...] status
  SHQ_CUDA(cudaMemsetAsync(d_ctl, 0, (nb + 16) * 8, st));
  o.dense = 1; o.n_groups = (uint32_t)nb; o.totals = d_ctl; o.ticket = reinterpret_cast<unsigned int*>(d_ctl + 4); o.status = d_ctl + 8;
  o.scratch_keys = o.keys; o.scratch_counts = o.counts; o.scratch_chunk = o.chunk; o.scratch_keys2 = o.keys2; o.scratch_counts2 = o.counts2;
  o.keys = A->keys->as<uint64_t>(); o.counts = A->counts->as<uint32_t>(); o.chunk = chunks ? A->chunk->as<uint16_t>() : nullptr;
  o.keys2 = split ? B->keys->as<uint64_t>() : nullptr; o.counts2 = split ? B->counts->as<uint32_t>() : nullptr;
  // most words of the batch are first occurrences (a batch of a large genome covers it less than once): insert directly
  o.direct = (bb.ratio >= 0 ? bb.ratio : (h->ratio_seen ? h->distinct_ratio : 1.0)) > 0.3 ? 1 : 0;
#define SHQ_AGG_LAUNCH(SP, CH) do { \
    const size_t smem = agg_smem_bytes(CH, SP); \
    SHQ_CUDA(cudaFuncSetAttribute(bin_aggregate_kernel<SP, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    bin_aggregate_kernel<SP, CH><<<(unsigned)nb, kAggThreads, smem, st>>>(bb.words, bb.off, bb.lb, h->key_lo, clo, chi, cstar, o); } while (0)
  if (split) SHQ_AGG_LAUNCH(true, false);
  else if (chunks) SHQ_AGG_LAUNCH(false, true);
  else SHQ_AGG_LAUNCH(false, false);
#undef SHQ_AGG_LAUNCH
  L++;
  if (bb.spec && !bb.spec_checked) {   // did every bin (of both levels) fit its fixed range?  (read once per binned batch)
    bb.spec_checked = true;
    int o1 = 0, o2 = 0;
    read_pair(h, h->d_misc.as<int>() + 20, &o1, h->d_misc.as<int>() + 24, &o2);
    if ((bb.level1_spec && o1) || (bb.fine_bits && o2)) throw BinOverflow{};
  }
  {
    // totals and the split-overflow flag with one round trip
    unsigned long long* pz = (unsigned long long*)pinned(h, 64);
    SHQ_CUDA(cudaMemcpyAsync(pz, d_ctl, 32, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaMemcpyAsync(pz + 4, d_ovf, 4, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    h->stats.d2h_bytes += 36; h->stats.host_syncs++;
    if (split && (uint32_t)pz[4]) return false;
    A->n = (size_t)pz[0]; A->occ = pz[2];
    if (split) { B->n = (size_t)pz[1]; B->occ = pz[3]; }
  }
  h->stats.agg_launches++; h->stats.agg_words += (uint64_t)bb.n_keys << fixed_bits; h->stats.agg_distinct += A->n + (split ? B->n : 0);
  if (!split && clo == 0 && (chi == 0xFFFFFFFFu || chi + 1 >= (uint32_t)bb.nchunks) && A->occ > (uint64_t)bb.n_bins * 64)
    { h->distinct_ratio = (double)A->n / (double)A->occ; h->ratio_seen = true; }   // whole batch seen: the next one is sized by it
  if (first_hist) {
    uint32_t* p = (uint32_t*)pinned(h, (size_t)bb.nchunks * 4);
    SHQ_CUDA(cudaMemcpyAsync(p, d_fh, (size_t)bb.nchunks * 4, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    first_hist->assign(p, p + bb.nchunks);
  }
  SHQ_CUDA(cudaGetLastError());
  return true;
}

// Merge run R with the state into the SHADOW copy (st_keys2 / st_counts2); the state itself stays untouched until
// commit_shadow().  Returns the number of entries of the merged state.  chunk_hist_host (optional; R must carry first
// chunks): per-chunk histogram of the keys of R that are NEW, a by-product of the same pass.  An armed batch is merged
// speculatively this way: if the deNoise trigger turns out to be crossed inside the batch the shadow is simply dropped.
uint64_t merge_shadow(shq_handle_s* h, const Run& R, int nchunks, std::vector<uint32_t>* chunk_hist_host) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  Phase ph(h, &h->stats.ms_merge);
  const size_t na = h->n_state, nb = R.n;
  const size_t total = na + nb;
  if (chunk_hist_host) chunk_hist_host->assign((size_t)nchunks, 0);
  h->st_keys2.reserve((total + 4) * 8); h->st_counts2.reserve((total + 4) * 4);
  if (total == 0) return 0;
  const uint16_t* Bch = (chunk_hist_host && R.chunk) ? R.chunk->as<uint16_t>() : nullptr;
  if (chunk_hist_host && !Bch) SHQ_FAIL(SHQ_ERR_STATE, "merge: first chunks requested from a run without them");
  if (chunk_hist_host && nchunks > kMergeHistMax) SHQ_FAIL(SHQ_ERR_ARG, "too many chunks in one batch (max %d)", kMergeHistMax);
  const bool lookback = h->merge_lookback && total < ((uint64_t)1 << 32);
  const int tile = (lookback && !h->merge_tma) ? kLbTile : kMergeTile;
  const size_t ntiles = (total + tile - 1) / tile;
  h->d_split.reserve((ntiles + 2) * 8);
  h->d_heads.reserve((ntiles + 2) * 4);
  h->d_heads64.reserve((ntiles + 4) * 8);
  const uint64_t* A = h->st_keys.as<uint64_t>(); const uint32_t* Ac = h->st_counts.as<uint32_t>();
  const uint64_t* Bk = R.keys->as<uint64_t>(); const uint32_t* Bc = R.counts->as<uint32_t>();
  uint32_t* d_hist = nullptr;
  if (chunk_hist_host) {
    h->d_chunk_hist.reserve((size_t)(kMergeHistMax + kAggBuckets) * 4);
    d_hist = h->d_chunk_hist.as<uint32_t>();
    SHQ_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)kMergeHistMax * 4, st));
  }
  merge_partition_kernel<<<cdiv(ntiles + 1, 128), 128, 0, st>>>(A, na, Bk, nb, h->d_split.as<uint64_t>(), ntiles, tile); L++;
  uint64_t n_out = 0;
  if (lookback) {
    // single pass: the output can hold na + nb entries at most; the real size comes back with the last tile
    unsigned long long* status = h->d_heads64.as<unsigned long long>();
    unsigned long long* d_nout = status + ntiles; unsigned int* ticket = (unsigned int*)(status + ntiles + 1);
    SHQ_CUDA(cudaMemsetAsync(status, 0, (ntiles + 2) * 8, st));
    if (h->merge_tma) {
      // persistent CTAs, inputs staged by the TMA engine (cp.async.bulk + mbarrier), two stages per CTA
      SHQ_CUDA(cudaFuncSetAttribute(merge_tiles_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMergeTmaSmem));
      SHQ_CUDA(cudaFuncSetAttribute(merge_tiles_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMergeTmaSmem));
      const unsigned grid = (unsigned)std::min<size_t>(ntiles, (size_t)kNumSMs * 2);   // 2 CTAs of 78 KB fit an SM
      if (d_hist)
        merge_tiles_tma_kernel<true><<<grid, kMergeThreads, kMergeTmaSmem, st>>>(A, Ac, na, Bk, Bc, nb, h->d_split.as<uint64_t>(), ntiles, status, ticket,
                                                                              h->st_keys2.as<uint64_t>(), h->st_counts2.as<uint32_t>(), d_nout, h->d_err, Bch, d_hist);
      else
        merge_tiles_tma_kernel<false><<<grid, kMergeThreads, kMergeTmaSmem, st>>>(A, Ac, na, Bk, Bc, nb, h->d_split.as<uint64_t>(), ntiles, status, ticket,
                                                                               h->st_keys2.as<uint64_t>(), h->st_counts2.as<uint32_t>(), d_nout, h->d_err);
    } else {
    SHQ_CUDA(cudaFuncSetAttribute(merge_tiles_lookback_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMergeLbSmem));
    SHQ_CUDA(cudaFuncSetAttribute(merge_tiles_lookback_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMergeLbSmem));
    if (d_hist)
      merge_tiles_lookback_kernel<true><<<(unsigned)ntiles, kMergeThreads, kMergeLbSmem, st>>>(A, Ac, na, Bk, Bc, nb, h->d_split.as<uint64_t>(), ntiles, status, ticket,
                                                                                h->st_keys2.as<uint64_t>(), h->st_counts2.as<uint32_t>(), d_nout, h->d_err, Bch, d_hist);
    else
      merge_tiles_lookback_kernel<false><<<(unsigned)ntiles, kMergeThreads, kMergeLbSmem, st>>>(A, Ac, na, Bk, Bc, nb, h->d_split.as<uint64_t>(), ntiles, status, ticket,
                                                                                 h->st_keys2.as<uint64_t>(), h->st_counts2.as<uint32_t>(), d_nout, h->d_err);
    }
    L++;
    SHQ_CUDA(cudaGetLastError());
    if (d_hist) {      // total and histogram with one round trip
      char* pz = (char*)pinned(h, 16 + (size_t)nchunks * 4);
      SHQ_CUDA(cudaMemcpyAsync(pz, d_nout, 8, cudaMemcpyDeviceToHost, st));
      SHQ_CUDA(cudaMemcpyAsync(pz + 16, d_hist, (size_t)nchunks * 4, cudaMemcpyDeviceToHost, st));
      SHQ_CUDA(cudaStreamSynchronize(st));
      h->stats.d2h_bytes += 8 + (size_t)nchunks * 4; h->stats.host_syncs++;
      memcpy(&n_out, pz, 8);
      chunk_hist_host->assign((uint32_t*)(pz + 16), (uint32_t*)(pz + 16) + nchunks);
    } else {
      n_out = read_scalar(h, (const uint64_t*)d_nout);
    }
  } else {
    // count pass + scan + write pass (SHQ_MERGE_LOOKBACK=0, or a state of 2^32 entries and more)
    merge_tiles_kernel<false><<<(unsigned)ntiles, kMergeThreads, 0, st>>>(A, Ac, na, Bk, Bc, Bch, nb, h->d_split.as<uint64_t>(),
                                                                         h->d_heads.as<uint32_t>(), nullptr, nullptr, nullptr, d_hist, h->d_err); L++;
    widen_u32_kernel<<<cdiv(ntiles, 256), 256, 0, st>>>(h->d_heads.as<uint32_t>(), h->d_heads64.as<uint64_t>(), ntiles); L++;
    exclusive_scan_inplace<SumU64>(h->d_heads64.as<uint64_t>(), ntiles, h->d_scan_scratch, st, h->d_heads64.as<uint64_t>() + ntiles, &L);
    n_out = read_scalar(h, h->d_heads64.as<uint64_t>() + ntiles);
    if (chunk_hist_host) {
      uint32_t* pz = (uint32_t*)pinned(h, (size_t)nchunks * 4);
      SHQ_CUDA(cudaMemcpyAsync(pz, d_hist, (size_t)nchunks * 4, cudaMemcpyDeviceToHost, st));
      SHQ_CUDA(cudaStreamSynchronize(st));
      h->stats.d2h_bytes += (size_t)nchunks * 4; h->stats.host_syncs++;
      chunk_hist_host->assign(pz, pz + nchunks);
    }
    if (n_out >= ((uint64_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "more than 2^32 distinct keys in one shard");
    merge_tiles_kernel<true><<<(unsigned)ntiles, kMergeThreads, 0, st>>>(A, Ac, na, Bk, Bc, nullptr, nb, h->d_split.as<uint64_t>(), nullptr,
                                                                        h->d_heads64.as<uint64_t>(), h->st_keys2.as<uint64_t>(),
                                                                        h->st_counts2.as<uint32_t>(), nullptr, h->d_err); L++;
    SHQ_CUDA(cudaGetLastError());
  }
  h->stats.merge_launches++; h->stats.merge_in += total; h->stats.merge_out += n_out;
  return n_out;
}
// the shadow becomes the state; returns the number of new keys
uint64_t commit_shadow(shq_handle_s* h, uint64_t n_out) {
  const uint64_t n_new = n_out - h->n_state;
  std::swap(h->st_keys, h->st_keys2); std::swap(h->st_counts, h->st_counts2);
  h->n_state = n_out;
  h->image_valid = false;
  return n_new;
}
// merge + commit; the very first run simply becomes the state
uint64_t merge_run(shq_handle_s* h, const Run& R) {
  if (h->n_state == 0) {
    cudaStream_t st = h->stream;
    Phase ph(h, &h->stats.ms_merge);
    const size_t nb = R.n;
    h->st_keys.reserve((nb + 4) * 8); h->st_counts.reserve((nb + 4) * 4);
    if (nb) {
      SHQ_CUDA(cudaMemcpyAsync(h->st_keys.p, R.keys->p, nb * 8, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(h->st_counts.p, R.counts->p, nb * 4, cudaMemcpyDeviceToDevice, st));
    }
    h->n_state = nb; h->image_valid = false;
    return nb;
  }
  return commit_shadow(h, merge_shadow(h, R, 1, nullptr));
}

// positions of every state entry in the canonical layout -> d_pos; returns one past the last used slot.
// q0 = quotient of slot 0 of the coordinate system, init = composed map of everything before entry 0 (both trivial
// for an unsharded filter); agg_out (host, optional) receives the composed map of THIS state alone.
uint64_t compute_positions(shq_handle_s* h, uint64_t q0, MaxPlus init, MaxPlus* agg_out = nullptr, bool positions = true) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  size_t n = h->n_state;
  if (n == 0) { if (agg_out) *agg_out = MaxPlusOp::id(); return (uint64_t)(init.a > init.b ? init.a : init.b); }
  const size_t ntiles = (n + kScanTile - 1) / kScanTile;
  h->d_pairs.reserve((ntiles + 4) * sizeof(MaxPlus)); h->d_pos.reserve((n + 4) * 8);
  h->d_misc.reserve(256);
  MaxPlus* agg = h->d_pairs.as<MaxPlus>();
  layout_scan_kernel<false><<<(unsigned)ntiles, kScanThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), n, q0, agg, nullptr, nullptr, init); L++;
  exclusive_scan_inplace<MaxPlusOp>(agg, ntiles, h->d_scan_scratch, st, h->d_misc.as<MaxPlus>() + 1, &L);
  if (agg_out) *agg_out = read_scalar(h, h->d_misc.as<MaxPlus>() + 1);
  if (!positions) return 0;
  layout_scan_kernel<true><<<(unsigned)ntiles, kScanThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), n, q0, agg,
                                                                   h->d_pos.as<uint64_t>(), h->d_misc.as<MaxPlus>(), init); L++;
  MaxPlus tot = read_scalar(h, h->d_misc.as<MaxPlus>());
  long long e = tot.a > tot.b ? tot.a : tot.b;   // f_total(0)
  return (uint64_t)e;
}
uint64_t compute_positions(shq_handle_s* h) { return compute_positions(h, h->key_lo >> 8, MaxPlusOp::id()); }

void ensure_state(shq_handle_s* h);

void materialise(shq_handle_s* h) {
  if (h->image_valid) return;
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  Phase ph(h, &h->stats.ms_layout);
  h->stats.layout_launches++; h->stats.layout_entries += h->n_state;
  uint64_t used = compute_positions(h);
  if (used > h->xnslots) SHQ_FAIL(SHQ_ERR_CAPACITY, "filter over-full: layout needs slot %llu but xnslots = %llu",
                                  (unsigned long long)used, (unsigned long long)h->xnslots);
  size_t tiles = (h->nblocks + kTileBlocks - 1) / kTileBlocks;
  h->image.reserve(tiles * kTileBytes + 16);
  image_tiles_kernel<<<(unsigned)tiles, kImgThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), h->d_pos.as<uint64_t>(),
                                                            h->n_state, h->key_lo >> 8, h->nblocks, h->image.as<uint8_t>()); L++;
  SHQ_CUDA(cudaGetLastError());
  h->image_valid = true;
}

// image -> state (after shq_load)
void ensure_state(shq_handle_s* h) {
  if (!h->image_only) return;
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  ImageView im{h->image.as<uint8_t>(), h->nblocks, h->xnslots, 0};
  size_t nb = h->nblocks;
  h->d_heads.reserve((nb + 2) * 4); h->d_heads64.reserve((nb + 2) * 8);
  decode_blocks_kernel<false><<<cdiv(nb, 128), 128, 0, st>>>(im, h->d_heads.as<uint32_t>(), nullptr, nullptr, nullptr, h->d_err); L++;
  widen_u32_kernel<<<cdiv(nb, 256), 256, 0, st>>>(h->d_heads.as<uint32_t>(), h->d_heads64.as<uint64_t>(), nb); L++;
  exclusive_scan_inplace<SumU64>(h->d_heads64.as<uint64_t>(), nb, h->d_scan_scratch, st, h->d_heads64.as<uint64_t>() + nb, &L);
  uint64_t n = read_scalar(h, h->d_heads64.as<uint64_t>() + nb);
  h->st_keys.reserve((n + 4) * 8); h->st_counts.reserve((n + 4) * 4);
  decode_blocks_kernel<true><<<cdiv(nb, 128), 128, 0, st>>>(im, nullptr, h->d_heads64.as<uint64_t>(), h->st_keys.as<uint64_t>(),
                                                          h->st_counts.as<uint32_t>(), h->d_err); L++;
  check_err_flag(h);
  h->n_state = n;
  h->image_only = false;
}

std::string current_date_time() {      // base/global.cpp:14-21
  time_t now = time(nullptr);
  struct tm t; localtime_r(&now, &t);
  char buf[80];
  strftime(buf, sizeof buf, "%Y-%m-%d.%X", &t);
  return buf;
}
void log_denoise(shq_handle_s* h, bool ready) {   // the two progress lines of the scheduler, cqf/CQF_mt.h:867-868, 912-913
  if (!h->log_progress) return;
  if (ready) fprintf(stderr, "%s\n", current_date_time().c_str());
  fprintf(stderr, "%s DeNoise: ndistinct_elts/total_elts.%llu/%llu ndistinct_true_elts.%llu\n", ready ? "Ready for" : "Finished",
          (unsigned long long)h->ndistinct, (unsigned long long)h->nelts, (unsigned long long)h->ndistinct_true);
  if (!ready) fprintf(stderr, "%s\n", current_date_time().c_str());
}

uint64_t denoise_round_impl(shq_handle_s* h);
uint64_t denoise_round(shq_handle_s* h) {
  log_denoise(h, true);
  uint64_t r = denoise_round_impl(h);
  log_denoise(h, false);
  return r;
}
uint64_t denoise_round_impl(shq_handle_s* h) {
  ensure_state(h);
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  Phase ph(h, &h->stats.ms_denoise);
  size_t n = h->n_state;
  h->rounds_done++;
  h->stats.denoise_launches++; h->stats.denoise_entries += n;
  if (n == 0) return 0;
  if (h->nshards != 1) SHQ_FAIL(SHQ_ERR_STATE, "deNoise on a sharded handle is driven by the multi-GPU layer");
  compute_positions(h);
  uint32_t keep_cap = (uint32_t)(h->nslots >> 20) + 8;
  h->d_keep.reserve((size_t)(keep_cap + 4) * 4);
  uint32_t* n_keep = h->d_keep.as<uint32_t>(); uint32_t* keep_idx = n_keep + 4;
  denoise_ranges_kernel<<<1, 32, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), h->d_pos.as<uint64_t>(), n, h->nslots,
                                        h->xnslots, keep_idx, n_keep, keep_cap); L++;
  const size_t ntiles = (n + kScanTile - 1) / kScanTile;
  h->d_flags.reserve((ntiles + 4) * 8);
  uint64_t* tile_cnt = h->d_flags.as<uint64_t>();
  denoise_keep_kernel<<<cdiv(keep_cap, 128), 128, 0, st>>>(keep_idx, n_keep, keep_cap, h->st_counts.as<uint32_t>()); L++;
  denoise_tiles_kernel<false><<<(unsigned)ntiles, kScanThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), n, tile_cnt, nullptr, nullptr); L++;
  exclusive_scan_inplace<SumU64>(tile_cnt, ntiles, h->d_scan_scratch, st, tile_cnt + ntiles, &L);
  uint64_t kept = read_scalar(h, tile_cnt + ntiles);
  h->st_keys2.reserve((kept + 4) * 8); h->st_counts2.reserve((kept + 4) * 4);
  denoise_tiles_kernel<true><<<(unsigned)ntiles, kScanThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), n, tile_cnt,
                                                                     h->st_keys2.as<uint64_t>(), h->st_counts2.as<uint32_t>()); L++;
  SHQ_CUDA(cudaGetLastError());
  std::swap(h->st_keys, h->st_keys2); std::swap(h->st_counts, h->st_counts2);
  uint64_t removed = n - kept;
  h->n_state = kept;
  h->nelts -= removed; h->ndistinct -= removed;    // CQF_mt.h:1037-1038
  h->image_valid = false;
  return removed;
}

// One device batch: text of nchunks reference chunks (or raw keys) -> state, honouring the schedule.
//
// While deNoise rounds are pending ("armed") a round must fire after the first reference chunk that leaves
// ndistinct >= trigger (t = 1 rule, cqf/CQF_mt.h:837).  ndistinct only grows inside a batch, so the batch is first merged
// SPECULATIVELY into the shadow state as a whole: if the merged state stays below the trigger nothing can have fired and
// the shadow is committed -- an armed batch then costs exactly what an unarmed one does.  Only the (at most `rounds`)
// batches that contain a boundary pay for locating it: an aggregation that tracks the first chunk of every key, the
// histogram of new keys per chunk out of the same merge pass, and one split aggregation around the boundary chunk c*.
// When the new-key rate of the previous batch predicts a crossing, the tracking aggregation is run straight away.
void process_batch(shq_handle_s* h, const char* d_text, size_t n_text, const uint64_t* chunk_ends, int nchunks, int K,
                   const uint64_t* d_keys, size_t n_keys_in) {
  ensure_state(h);
  if (h->finished) SHQ_FAIL(SHQ_ERR_STATE, "handle already finished");
  const bool from_text = d_keys == nullptr;
  if (from_text && n_text == 0) {
    // only empty chunks (fastq_read_parts found no record boundary, cut == 0): the reference still evaluates
    // needDeNoise() after each of them (cqf/CQF_mt.h:829-839), so a round fires while ndistinct stays >= trigger
    for (int c = 0; c < nchunks; c++)
      if (h->rounds_left > 0 && h->ndistinct >= h->trigger) { h->rounds_left--; denoise_round(h); }
    return;
  }
  if (!from_text && n_keys_in == 0) return;
  bool armed = from_text && h->rounds_left > 0;
  Run A{&h->d_run_keys, &h->d_run_counts, &h->d_run_chunk}, B{&h->d_run2_keys, &h->d_run2_counts, nullptr};
  if (armed && (size_t)nchunks > batch_chunk_cap(h)) {
    // WIDE batch (batch_limits): more chunks than the packed word has chunk-id bits for.  Binned without chunk ids and
    // merged speculatively as a whole; if the trigger is reached inside after all, the halves are processed on their own
    // (each may be wide again), down to batches that keep their chunk ids.
    bool done = false;
    if (h->ndistinct < h->trigger && h->n_state > 0) {
      BatchBins wb;
      for (int attempt = 0; attempt < 2 && !done; attempt++) {
        try {
          bin_batch(h, d_text, n_text, chunk_ends, nchunks, K, nullptr, 0, false, &wb, 0, attempt == 0);
          if (attempt == 0) h->stats.batches++;
          if (wb.n_keys == 0) break;
          if (wb.n_keys > 4096) h->bytes_per_kmer = (double)n_text / (double)wb.n_keys;
          aggregate_bins(h, wb, 0, 0xFFFFFFFFu, false, 0, &A, nullptr, nullptr);
          const uint64_t n_out = merge_shadow(h, A, 1, nullptr);
          if (h->ndistinct + (n_out - h->n_state) < h->trigger) {
            const uint64_t n_new = commit_shadow(h, n_out);
            h->ndistinct += n_new; h->nelts += A.occ;
            if (A.occ > 4096) h->new_rate = (double)n_new / (double)A.occ;
            done = true;
          }
          break;
        } catch (const BinOverflow&) { h->stats.bin_fallbacks++; }
      }
      if (!done && wb.n_keys == 0) {      // no k-mer in the whole batch: only the per-chunk trigger checks remain (none can fire below the trigger)
        check_err_flag(h);
        return;
      }
      if (!done) { h->stats.spec_rejected++; h->stats.wide_fallbacks++; }
    }
    if (!done) {
      const int m = nchunks / 2;
      const uint64_t cut = chunk_ends[m - 1];
      process_batch(h, d_text, (size_t)cut, chunk_ends, m, K, nullptr, 0);
      std::vector<uint64_t> rest(chunk_ends + m, chunk_ends + nchunks);
      for (auto& e : rest) e -= cut;
      process_batch(h, d_text + cut, n_text - (size_t)cut, rest.data(), nchunks - m, K, nullptr, 0);
    } else {
      h->stats.wide_batches++;
      check_err_flag(h);
    }
    return;
  }
  BatchBins bb;
  bin_batch(h, d_text, n_text, chunk_ends, nchunks, K, d_keys, n_keys_in, armed, &bb, 0, true);
  h->stats.batches++;
  if (bb.n_keys == 0) return;
  if (from_text && bb.n_keys > 4096) h->bytes_per_kmer = (double)n_text / (double)bb.n_keys;
  uint32_t clo = 0;
  const uint32_t last = (uint32_t)(nchunks > 0 ? nchunks - 1 : 0);
  const uint32_t all = 0xFFFFFFFFu;
  // committed keys / occurrences of a run
  auto account = [&](uint64_t n_new, const Run& R) {
    h->ndistinct += n_new; h->nelts += R.occ;
    if (R.occ > 4096) h->new_rate = (double)n_new / (double)R.occ;
  };
  auto fire = [&]() { h->rounds_left--; denoise_round(h); };
  // does the new-key rate seen so far predict that chunks [clo, last] reach the trigger?  (no history: assume they might)
  auto predicts_crossing = [&]() {
    const double frac = (double)(last - clo + 1) / (double)(last + 1);
    double est_new;
    if (bb.new_frac >= 0 && clo == 0) {
      // measured on this batch's sample; 3 sigma of the sample count plus 10 % on top
      est_new = bb.new_frac * (double)bb.n_keys * (1.10 + 3.0 / sqrt((double)bb.sample_new + 1.0));
    } else {
      const double rate = h->new_rate < 0 ? 1.0 : h->new_rate * 1.15;
      est_new = rate * frac * (double)bb.n_keys;
    }
    return (double)h->ndistinct + est_new + 4096.0 >= (double)h->trigger;
  };
  for (;;) try {
    if (!(armed && h->rounds_left > 0)) {                 // no round pending: chunks [clo, last] go in as one run
      aggregate_bins(h, bb, clo, armed ? last : all, false, 0, &A, nullptr, nullptr);
      account(merge_run(h, A), A);
      break;
    }
    if (h->ndistinct >= h->trigger) {                     // still over the trigger after a sweep: the next chunk fires again
      aggregate_bins(h, bb, clo, clo, false, 0, &A, nullptr, nullptr);
      account(merge_run(h, A), A);
      fire();
      if (clo == last) break;
      clo++;
      continue;
    }
    if (!predicts_crossing()) {
      aggregate_bins(h, bb, clo, last, false, 0, &A, nullptr, nullptr);
      if (h->n_state == 0) {                              // nothing to merge with: the run's size is the new ndistinct
        if (h->ndistinct + A.n < h->trigger) { account(merge_run(h, A), A); break; }
      } else {
        const uint64_t n_out = merge_shadow(h, A, 1, nullptr);
        if (h->ndistinct + (n_out - h->n_state) < h->trigger) { account(commit_shadow(h, n_out), A); break; }
      }
      h->stats.spec_rejected++;                           // the trigger is crossed inside [clo, last]: locate the chunk below
    }
    // first chunk of every key + histogram of the new keys per chunk -> c* = first chunk after which ndistinct >= trigger
    std::vector<uint32_t> hist;
    uint64_t n_out = 0;
    const bool empty_state = h->n_state == 0;
    aggregate_bins(h, bb, clo, last, false, 0, &A, nullptr, empty_state ? &hist : nullptr, true);
    if (!empty_state) n_out = merge_shadow(h, A, nchunks, &hist);
    uint64_t nd = h->ndistinct; int cstar = -1;
    for (uint32_t c = clo; c <= last; c++) { nd += hist[c]; if (nd >= h->trigger) { cstar = (int)c; break; } }
    if (cstar < 0 || (uint32_t)cstar == last) {           // no boundary, or at the very end of the batch: the merge stands
      account(empty_state ? merge_run(h, A) : commit_shadow(h, n_out), A);
      if (cstar >= 0) fire();
      break;
    }
    // boundary inside the batch: one split pass yields both sides of c*
    if (aggregate_bins(h, bb, clo, last, true, (uint32_t)cstar, &A, &B, nullptr)) {
      account(merge_run(h, A), A);
      fire();
      if (h->rounds_left == 0 || h->ndistinct >= h->trigger) {
        if (h->rounds_left == 0) { account(merge_run(h, B), B); break; }
      } else {
        // the rest of the batch may hold the next boundary only if it crosses the trigger again: speculate once more
        const uint64_t n2 = merge_shadow(h, B, 1, nullptr);
        if (h->ndistinct + (n2 - h->n_state) < h->trigger) { account(commit_shadow(h, n2), B); break; }
        h->stats.spec_rejected++;
      }
      clo = (uint32_t)cstar + 1;                          // replay chunks (c*, last] with tracking
      continue;
    }
    // split counters overflowed (a key seen >= 65535 times on one side): two plain passes instead
    aggregate_bins(h, bb, clo, (uint32_t)cstar, false, 0, &A, nullptr, nullptr);
    account(merge_run(h, A), A);
    fire();
    clo = (uint32_t)cstar + 1;
  } catch (const BinOverflow&) {
    // thrown by the first aggregation of a one-pass batch, before anything was merged: bin again with exact offsets
    h->stats.bin_fallbacks++;
    bin_batch(h, d_text, n_text, chunk_ends, nchunks, K, d_keys, n_keys_in, armed, &bb, 0, false);
    if (bb.n_keys == 0) return;
  }
  check_err_flag(h);
}

// All keys of the reads in d_text as a device list: read r owns slots [off[r], off[r+1]) (off = scan of len-K+1);
// slots the N rule leaves unused hold ~0.  Returns the number of slots.
size_t hash_to_keylist(shq_handle_s* h, const char* d_text, size_t n, int K, uint64_t* d_out, size_t cap) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  if (K < 1 || K > 1024) SHQ_FAIL(SHQ_ERR_ARG, "K=%d outside [1,1024]", K);
  if (n >= ((size_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "text of %zu bytes exceeds 4 GiB per call", n);
  const char* src = d_text;
  uint32_t n_lines = index_lines(h, src, n);
  if (n_lines & 3) SHQ_FAIL(SHQ_ERR_MALFORMED, "text holds %u lines, not a multiple of 4", n_lines);
  uint32_t n_reads = n_lines / 4;
  if (n_reads == 0) return 0;
  h->d_heads64.reserve((size_t)(n_reads + 2) * 8);
  uint64_t* off = h->d_heads64.as<uint64_t>();
  read_window_counts_kernel<<<cdiv(n_reads, 256), 256, 0, st>>>(h->d_nl_pos.as<uint32_t>(), n_reads, K, off); L++;
  exclusive_scan_inplace<SumU64>(off, n_reads, h->d_scan_scratch, st, off + n_reads, &L);
  uint64_t total = read_scalar(h, off + n_reads);
  if (total > cap) SHQ_FAIL(SHQ_ERR_ARG, "key buffer too small: %llu slots needed, capacity %zu", (unsigned long long)total, cap);
  HashParams hp{};
  hp.text = src; hp.nl_pos = h->d_nl_pos.as<uint32_t>(); hp.n_reads = n_reads; hp.K = K; hp.hb = h->hb;
  roll_keys_kernel<<<cdiv(n_reads, 256), 256, 0, st>>>(hp, off, d_out); L++;
  SHQ_CUDA(cudaGetLastError());
  h->stats.keys_hashed += total;
  return (size_t)total;
}

// Process ONE batch out of the ring: the longest run of chunks at the front of pin_chunks that is contiguous, of one K,
// within the batch limits and -- unless `all` -- already copied; at least kPinMinBatch bytes of it are waited for.
constexpr size_t kPinMinBatch = (size_t)256 << 20;
void pin_process_next(shq_handle_s* h, bool all) {
  if (h->pin_chunks.empty()) return;
  size_t lim_bytes, lim_chunks;
  batch_limits(h, &lim_bytes, &lim_chunks);
  // at least half a full-size batch: every batch pays one pass over the resident state, and many small batches would make
  // the GPU the limiter of a job that the host link should be (measured: 30 batches per step kept the GPU busy for
  // 427 ms of configs[2], 13 full-size ones for 283)
  const size_t min_bytes = std::min(lim_bytes, std::max(kPinMinBatch, lim_bytes / 2));
  const shq_handle_s::PinChunk first = h->pin_chunks.front();
  std::vector<uint64_t> ends;
  size_t bytes = 0, taken = 0;
  for (const auto& c : h->pin_chunks) {
    if (taken) {
      if (c.K != first.K || c.off != first.off + bytes || bytes + c.n > lim_bytes || taken >= lim_chunks) break;
      if (!all && bytes >= min_bytes && cudaEventQuery(c.ev) != cudaSuccess) break;   // not there yet: the next batch takes it
    }
    bytes += c.n; taken++; ends.push_back(bytes);
  }
  const cudaEvent_t last = h->pin_chunks[taken - 1].ev;
  SHQ_CUDA(cudaStreamWaitEvent(h->stream, last, 0));
  SHQ_CUDA(cudaEventSynchronize(last));
  for (size_t i = 0; i < taken; i++) { h->pin_ev_pool.push_back(h->pin_chunks.front().ev); h->pin_chunks.pop_front(); }
  h->pin_unproc -= bytes;
  process_batch(h, h->pin_ring.as<char>() + first.off, bytes, ends.data(), (int)ends.size(), first.K, nullptr, 0);
}
void pin_drain(shq_handle_s* h) {
  while (!h->pin_chunks.empty()) pin_process_next(h, true);
}
// room for n more bytes at the head of the ring without touching unprocessed text?  (may wrap the head to 0)
bool pin_make_room(shq_handle_s* h, size_t n) {
  if (h->pin_chunks.empty()) { h->pin_head = 0; return n <= h->pin_ring_size; }
  const size_t tail = h->pin_chunks.front().off;
  if (h->pin_head >= tail) {                           // unprocessed text is [tail, head)
    if (h->pin_head + n <= h->pin_ring_size) return true;
    if (n < tail) { h->pin_head = 0; return true; }    // wrap: the next batch starts at the front of the ring
    return false;
  }
  return h->pin_head + n < tail;                       // wrapped: unprocessed text is [tail, ...) and [0, head)
}

void flush_stage(shq_handle_s* h) {
  pin_drain(h);
  if (h->h_stage_fill == 0) {
    if (!h->stage_chunk_ends.empty()) {        // nothing but empty chunks
      std::vector<uint64_t> e; e.swap(h->stage_chunk_ends);
      process_batch(h, nullptr, 0, e.data(), (int)e.size(), h->stage_K, nullptr, 0);
    }
    return;
  }
  size_t n = h->h_stage_fill;
  h->d_text.reserve(n + 64);
  SHQ_CUDA(cudaMemcpyAsync(h->d_text.p, h->h_stage, n, cudaMemcpyHostToDevice, h->stream));
  h->stats.h2d_bytes += n;
  std::vector<uint64_t> ends;
  ends.swap(h->stage_chunk_ends);
  h->h_stage_fill = 0;
  process_batch(h, h->d_text.as<char>(), n, ends.data(), (int)ends.size(), h->stage_K, nullptr, 0);
}

int guard(shq_handle_s* h, void (*fn)(shq_handle_s*, void*), void* arg) {
  try {
    if (h) SHQ_CUDA(cudaSetDevice(h->device));
    fn(h, arg);
    return SHQ_OK;
  } catch (const CudaError& e) {
    char b[512]; snprintf(b, sizeof b, "%s: %s (%s:%d)", e.what, cudaGetErrorString(e.err), e.file, e.line);
    if (h) h->last_error = b; else fprintf(stderr, "shq: %s\n", b);
    return SHQ_ERR_CUDA;
  } catch (const StatusError& e) {
    if (h) h->last_error = e.msg; else fprintf(stderr, "shq: %s\n", e.msg.c_str());
    return e.status;
  } catch (const std::bad_alloc&) {
    if (h) h->last_error = "host allocation failed";
    return SHQ_ERR_ARG;
  }
}
#define SHQ_API_BEGIN(h) return guard((h), [](shq_handle_s* h, void* argp) { auto& a = *(decltype(args)*)argp;
#define SHQ_API_END }, &args);

void init_handle(shq_handle_s* h, int qb, int hb, uint32_t seed, int device, int shard, int nshards) {
  if (qb < 1 || qb > 40 || hb - qb != 8) SHQ_FAIL(SHQ_ERR_ARG, "need 1 <= qb <= 40 and hb - qb == 8 (got qb=%d hb=%d)", qb, hb);
  if (nshards < 1 || (nshards & (nshards - 1)) || shard < 0 || shard >= nshards || (1ULL << qb) < (unsigned long long)nshards * 64)
    SHQ_FAIL(SHQ_ERR_ARG, "bad shard %d of %d", shard, nshards);
  if (const char* e = getenv("SHQ_MAX_BATCH_MB")) { long v = atol(e); if (v >= 1 && v <= 3072) { h->max_batch_bytes = (size_t)v << 20; h->batch_fixed = true; } }
  if (const char* e = getenv("SHQ_BATCH_CAP_MB")) { long v = atol(e); if (v >= 64 && v <= 3072) h->batch_cap_bytes = (size_t)v << 20; }
  if (const char* e = getenv("SHQ_ARMED_NEAR")) h->armed_near = atoi(e) != 0;
  if (const char* e = getenv("SHQ_WIDE")) { int v = atoi(e); if (v >= 0 && v <= 2) h->wide_batches = v; }
  if (const char* e = getenv("SHQ_ARMED_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= 1024) h->armed_chunk_cap = v; }
  if (const char* e = getenv("SHQ_SPEC_BINS")) h->spec_bins = atoi(e) != 0;
  if (const char* e = getenv("SHQ_FINE_TARGET")) { double v = atof(e); if (v >= 500 && v <= 7000) h->fine_target = v; }
  if (const char* e = getenv("SHQ_FINE_BITS")) { int v = atoi(e); if (v >= 0 && v <= 5) h->fine_bits_forced = v; }
  if (const char* e = getenv("SHQ_MERGE_LOOKBACK")) h->merge_lookback = atoi(e) != 0;
  if (const char* e = getenv("SHQ_MERGE_TMA")) h->merge_tma = atoi(e) != 0;
  if (const char* e = getenv("SHQ_BIN_BITS")) { int v = atoi(e); if (v >= 4 && v <= 13) h->bin_bits = v; }
  SHQ_CUDA(cudaSetDevice(device));
  h->device = device; h->qb = qb; h->hb = hb; h->seed = seed; h->shard = shard; h->nshards = nshards;
  h->nslots = 1ULL << qb;
  h->xnslots = h->nslots + (uint64_t)(10 * sqrt((double)h->nslots));   // gqf.c:2198
  h->nblocks = (h->xnslots + 63) / 64;
  uint64_t range = 1ULL << hb;
  h->key_lo = range / nshards * shard; h->key_hi = range / nshards * (shard + 1);
  SHQ_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  SHQ_CUDA(cudaEventCreate(&h->ev0)); SHQ_CUDA(cudaEventCreate(&h->ev1));
  SHQ_CUDA(cudaMalloc(&h->d_err, sizeof(int)));
  SHQ_CUDA(cudaMemset(h->d_err, 0, sizeof(int)));
  memset(&h->stats, 0, sizeof h->stats);
}

void free_handle(shq_handle_s* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->unitig_cache) { shq_unitigs_s* u = (shq_unitigs_s*)h->unitig_cache; h->unitig_cache = nullptr; shq_unitigs_release(u); }
  DevBuf* bufs[] = {&h->st_keys, &h->st_counts, &h->st_keys2, &h->st_counts2, &h->image, &h->d_text, &h->d_nl_tiles, &h->d_nl_pos,
                    &h->d_chunk_ends, &h->d_chunk_line_end, &h->d_chunk_keys, &h->d_chunk_hist, &h->d_slow, &h->d_hist, &h->d_tot,
                    &h->d_binoff, &h->d_packed, &h->d_tmp_keys, &h->d_tmp_counts, &h->d_tmp_chunk, &h->d_nout, &h->d_occ,
                    &h->d_dense_off, &h->d_run_keys, &h->d_run_counts, &h->d_run_chunk, &h->d_split, &h->d_heads, &h->d_heads64,
                    &h->d_scan_scratch, &h->d_misc, &h->d_pairs, &h->d_pos, &h->d_flags, &h->d_keep, &h->d_q_in, &h->d_q_cnt, &h->d_q_was,
                    &h->d_tmp_keys2, &h->d_tmp_counts2, &h->d_run2_keys, &h->d_run2_counts, &h->d_seg, &h->d_fine, &h->d_fine_cur, &h->d_packed2, &h->d_bases, &h->d_sample, &h->d_fine2, &h->d_walk, &h->d_aggctl};
  for (DevBuf* b : bufs) b->release();
  h->pin_ring.release();
  for (void* p : h->mg_opened) cudaIpcCloseMemHandle(p);
  for (void* p : h->img_opened) cudaIpcCloseMemHandle(p);
  h->d_img_peer.release();
  for (uint32_t* p : h->mg_send) if (p) cudaFree(p);
  for (auto& c : h->pin_chunks) cudaEventDestroy(c.ev);
  for (auto e : h->pin_ev_pool) cudaEventDestroy(e);
  if (h->cstream) { cudaStreamSynchronize(h->cstream); cudaStreamDestroy(h->cstream); }
  if (h->d_err) cudaFree(h->d_err);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  resolve_phases(h);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// two page-locked staging buffers + events + a FILE*, released on every exit path (shq_save / shq_load)
struct PinnedPair {
  FILE* f = nullptr; uint8_t* pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
  void alloc(size_t piece) {
    for (int i = 0; i < 2; i++) { SHQ_CUDA(cudaMallocHost((void**)&pin[i], piece)); SHQ_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
  }
  int close_file() { int rc = f ? fclose(f) : 0; f = nullptr; return rc; }
  ~PinnedPair() {
    if (f) fclose(f);
    for (int i = 0; i < 2; i++) { if (pin[i]) cudaFreeHost(pin[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
  }
};

void build_header(shq_handle_s* h, uint8_t* hd) {   // qfmetadata, cqf/gqf.h:62-77 (x86-64 layout)
  memset(hd, 0, 128);
  auto put = [&](int off, uint64_t v) { memcpy(hd + off, &v, 8); };
  put(0, h->nblocks * kBlockBytes);
  memcpy(hd + 8, &h->seed, 4);
  put(16, h->nslots); put(24, h->xnslots); put(32, (uint64_t)h->hb); put(40, 0);
  put(48, (uint64_t)(h->hb - h->qb)); put(56, (uint64_t)(h->hb - h->qb));
  unsigned __int128 range = (unsigned __int128)h->nslots << (h->hb - h->qb);
  memcpy(hd + 64, &range, 16);
  put(80, h->nblocks); put(88, h->nelts); put(96, h->ndistinct); put(104, 0);
  put(112, h->xnslots / 65536 + 2);
}

// the image as the query / unitig kernels see it: one array, or the shards' arrays (own + peer-mapped)
ImageView image_view(shq_handle_s* h) {
  ImageView im{h->image.as<uint8_t>(), h->nblocks, h->xnslots, 0};
  if (!h->img_peer.empty()) {
    im.blocks = nullptr;
    im.shard_blocks = h->d_img_peer.as<uint8_t*>();
    im.blocks_per_shard = h->nslots / (uint64_t)h->nshards / 64;
    im.nshards = (uint32_t)h->nshards;
  }
  return im;
}

}  // namespace

// ============================================================== C ABI ========
extern "C" {

const char* shq_strerror(int s) {
  switch (s) {
    case SHQ_OK: return "ok";
    case SHQ_ERR_ARG: return "bad argument";
    case SHQ_ERR_CUDA: return "CUDA error";
    case SHQ_ERR_CAPACITY: return "filter capacity exceeded";
    case SHQ_ERR_MALFORMED: return "malformed FASTQ chunk";
    case SHQ_ERR_IO: return "I/O error";
    case SHQ_ERR_FORMAT: return "unsupported .cqf image";
    case SHQ_ERR_COUNT_OVERFLOW: return "count overflow";
    case SHQ_ERR_STATE: return "invalid state";
  }
  return "unknown";
}
const char* shq_last_error(shq_t h) { return h ? h->last_error.c_str() : ""; }

int shq_create_shard(int qb, int hb, uint32_t seed, int device, int shard, int nshards, shq_t* out) {
  if (!out) return SHQ_ERR_ARG;
  *out = nullptr;
  shq_handle_s* h = new shq_handle_s();
  struct { int qb, hb; uint32_t seed; int device, shard, nshards; } args{qb, hb, seed, device, shard, nshards};
  int rc = guard(h, [](shq_handle_s* h, void* p) {
    auto& a = *(decltype(args)*)p;
    init_handle(h, a.qb, a.hb, a.seed, a.device, a.shard, a.nshards);
  }, &args);
  if (rc != SHQ_OK && rc != SHQ_ERR_CUDA) { fprintf(stderr, "shq_create: %s\n", h->last_error.c_str()); }
  if (rc != SHQ_OK) { if (rc == SHQ_ERR_CUDA) fprintf(stderr, "shq_create: %s\n", h->last_error.c_str()); delete h; return rc; }
  *out = h;
  return SHQ_OK;
}
int shq_create(int qb, int hb, uint32_t seed, int device, shq_t* out) { return shq_create_shard(qb, hb, seed, device, 0, 1, out); }
void shq_destroy(shq_t h) { free_handle(h); }
int shq_reset(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->n_state = 0; h->ndistinct = 0; h->nelts = 0; h->new_rate = -1.0; h->ratio_seen = false; h->distinct_ratio = 0.125;
  h->mg_pending = {};
  h->image_valid = false; h->image_only = false;
  h->trigger = 0; h->rounds_left = 0; h->end_denoise = false; h->rounds_done = 0; h->finished = false;
  h->h_stage_fill = 0; h->stage_chunk_ends.clear();
  for (auto& c : h->pin_chunks) h->pin_ev_pool.push_back(c.ev);
  h->pin_chunks.clear(); h->pin_head = 0; h->pin_unproc = 0;
  if (h->cstream) cudaStreamSynchronize(h->cstream);
  return SHQ_OK;
}

int shq_set_denoise_schedule(shq_t h, uint64_t trigger, uint32_t rounds, int end_denoise) {
  if (!h) return SHQ_ERR_ARG;
  h->trigger = trigger; h->rounds_left = rounds; h->end_denoise = end_denoise != 0;
  return SHQ_OK;
}

int shq_set_progress_log(shq_t h, int enable, uint64_t ndistinct_true_elts) {
  if (!h) return SHQ_ERR_ARG;
  h->log_progress = enable != 0; h->ndistinct_true = ndistinct_true_elts;
  return SHQ_OK;
}

int shq_insert_fastq_chunk(shq_t h, const char* text, size_t n, int K) {
  if (!h || (!text && n)) return SHQ_ERR_ARG;
  struct { const char* text; size_t n; int K; } args{text, n, K};
  SHQ_API_BEGIN(h)
    // the staged path copies through one page-locked host buffer: its batches stay within 1 GiB and the armed chunk cap
    const size_t target = std::min(batch_target_bytes(h), kArmedBatchBytes);
    if ((h->h_stage_fill || !h->stage_chunk_ends.empty()) &&
        (h->stage_K != a.K || h->h_stage_fill + a.n > target || h->stage_chunk_ends.size() >= batch_chunk_cap(h))) flush_stage(h);
    if (a.n >= ((size_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "chunk too large");
    if (h->h_stage_cap < h->h_stage_fill + a.n) {
      size_t want = std::max(h->h_stage_fill + a.n, std::min(target, std::max((size_t)64 << 20, 2 * h->h_stage_cap)));
      char* p = nullptr;
      SHQ_CUDA(cudaMallocHost((void**)&p, want));
      if (h->h_stage_fill) memcpy(p, h->h_stage, h->h_stage_fill);
      if (h->h_stage) cudaFreeHost(h->h_stage);
      h->h_stage = p; h->h_stage_cap = want;
    }
    memcpy(h->h_stage + h->h_stage_fill, a.text, a.n);
    h->h_stage_fill += a.n;
    h->stage_chunk_ends.push_back(h->h_stage_fill);
    h->stage_K = a.K;
  SHQ_API_END
}

int shq_insert_fastq_pinned(shq_t h, const char* text, size_t n, int K) {
  if (!h || (!text && n)) return SHQ_ERR_ARG;
  if (n == 0) return SHQ_OK;
  struct { const char* text; size_t n; int K; } args{text, n, K};
  SHQ_API_BEGIN(h)
    if (h->h_stage_fill || !h->stage_chunk_ends.empty()) flush_stage(h);     // keep chunk order if the staged path was used before
    if (a.n >= ((size_t)1 << 31)) SHQ_FAIL(SHQ_ERR_ARG, "chunk too large");
    if (!h->cstream) SHQ_CUDA(cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
    // how far the copies may run ahead of the GPU: one full-size batch plus 1 GiB (so that the copy engine stays busy
    // while process_batch blocks this thread); the ring holds twice that, a batch never wraps
    const size_t cap = h->batch_fixed ? h->max_batch_bytes : h->batch_cap_bytes;
    const size_t ahead = cap + std::min(cap, kArmedBatchBytes);
    if (h->pin_ring_size < 2 * ahead + a.n || h->pin_ring_size < 4 * a.n) {
      pin_drain(h);
      SHQ_CUDA(cudaStreamSynchronize(h->cstream));
      h->pin_ring_size = std::max(2 * ahead + std::max(a.n, (size_t)64 << 20), 4 * a.n);
      h->pin_ring.reserve(h->pin_ring_size + 64);
      h->pin_head = 0;
    }
    while (!pin_make_room(h, a.n)) pin_process_next(h, false);
    cudaEvent_t ev;
    if (h->pin_ev_pool.empty()) SHQ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    else { ev = h->pin_ev_pool.back(); h->pin_ev_pool.pop_back(); }
    SHQ_CUDA(cudaMemcpyAsync(h->pin_ring.as<char>() + h->pin_head, a.text, a.n, cudaMemcpyHostToDevice, h->cstream));
    SHQ_CUDA(cudaEventRecord(ev, h->cstream));
    h->stats.h2d_bytes += a.n;
    h->pin_chunks.push_back({h->pin_head, a.n, a.K, ev});
    h->pin_head += a.n; h->pin_unproc += a.n;
    // the GPU takes its next batch when the copies are a full window ahead (fast producer: the batch is what has arrived
    // by now, so compute and copies balance out) or when a full-size batch has arrived (slow producer)
    size_t lim_bytes, lim_chunks;
    batch_limits(h, &lim_bytes, &lim_chunks);
    if (h->pin_unproc >= ahead) pin_process_next(h, false);
    else if (h->pin_unproc >= lim_bytes && cudaEventQuery(h->pin_chunks.back().ev) == cudaSuccess) pin_process_next(h, false);
  SHQ_API_END
}

int shq_insert_fastq_pinned_chunks(shq_t h, const char* pinned_text, size_t n, const uint64_t* chunk_ends, size_t nchunks, int K) {
  if (!h || (!pinned_text && n) || (!chunk_ends && nchunks)) return SHQ_ERR_ARG;
  if (nchunks && chunk_ends[nchunks - 1] != n) return SHQ_ERR_ARG;
  uint64_t b = 0;
  for (size_t c = 0; c < nchunks; c++) {
    int rc = shq_insert_fastq_pinned(h, pinned_text + b, (size_t)(chunk_ends[c] - b), K);
    if (rc) return rc;
    b = chunk_ends[c];
  }
  return SHQ_OK;
}

int shq_insert_fastq_device(shq_t h, const char* d_text, size_t n, const uint64_t* chunk_ends, size_t nchunks, int K) {
  if (!h || (!d_text && n) || (!chunk_ends && nchunks)) return SHQ_ERR_ARG;
  struct { const char* t; size_t n; const uint64_t* ce; size_t nc; int K; } args{d_text, n, chunk_ends, nchunks, K};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    if (a.nc && a.ce[a.nc - 1] != a.n) SHQ_FAIL(SHQ_ERR_ARG, "last chunk end must equal n");
    // split into device batches on chunk boundaries
    size_t c0 = 0;
    while (c0 < a.nc) {
      uint64_t base = c0 ? a.ce[c0 - 1] : 0;
      size_t c1 = c0;
      size_t target, cmax;
      batch_limits(h, &target, &cmax);
      while (c1 < a.nc && a.ce[c1] - base <= target && c1 - c0 < cmax) c1++;
      if (c1 == c0) c1 = c0 + 1;
      std::vector<uint64_t> ends(a.ce + c0, a.ce + c1);
      for (auto& e : ends) e -= base;
      const char* src = a.t + base;          // any byte offset: the line index copes with unaligned batch starts
      process_batch(h, src, ends.back(), ends.data(), (int)ends.size(), a.K, nullptr, 0);
      c0 = c1;
    }
  SHQ_API_END
}

int shq_insert_keys_device(shq_t h, const uint64_t* d_keys, size_t n) {
  if (!h || (!d_keys && n)) return SHQ_ERR_ARG;
  struct { const uint64_t* k; size_t n; } args{d_keys, n};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    const size_t step = (size_t)1 << 28;
    for (size_t o = 0; o < a.n; o += step) process_batch(h, nullptr, 0, nullptr, 0, 0, a.k + o, std::min(step, a.n - o));
  SHQ_API_END
}

int shq_insert_keys(shq_t h, const uint64_t* keys, size_t n) {
  if (!h || (!keys && n)) return SHQ_ERR_ARG;
  struct { const uint64_t* k; size_t n; } args{keys, n};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    const size_t step = (size_t)1 << 27;
    for (size_t o = 0; o < a.n; o += step) {
      size_t m = std::min(step, a.n - o);
      h->d_q_in.reserve(m * 8);
      SHQ_CUDA(cudaMemcpyAsync(h->d_q_in.p, a.k + o, m * 8, cudaMemcpyHostToDevice, h->stream));
      h->stats.h2d_bytes += m * 8;
      process_batch(h, nullptr, 0, nullptr, 0, 0, h->d_q_in.as<uint64_t>(), m);
    }
  SHQ_API_END
}

int shq_sync(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  int args = 0;
  SHQ_API_BEGIN(h)
    (void)a;
    flush_stage(h);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}

int shq_sync_copies(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  if (!h->cstream) return SHQ_OK;
  cudaSetDevice(h->device);
  return cudaStreamSynchronize(h->cstream) == cudaSuccess ? SHQ_OK : SHQ_ERR_CUDA;
}

int shq_finish(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  int args = 0;
  SHQ_API_BEGIN(h)
    (void)a;
    flush_stage(h);
    // EOF rule of fastq_to_uint64kmers_prod (CQF_mt.h:858-866): one more round if the trigger still
    // holds with rounds left, or if --endDeNoise was given.
    if (!h->finished) {
      if ((h->rounds_left > 0 && h->ndistinct >= h->trigger) || h->end_denoise) {
        if (h->rounds_left > 0) h->rounds_left--;
        denoise_round(h);
      }
      h->finished = true;
    }
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}

int shq_counts(shq_t h, uint64_t* ndistinct, uint64_t* nelts) {
  int rc = shq_sync(h);
  if (rc) return rc;
  if (ndistinct) *ndistinct = h->ndistinct;
  if (nelts) *nelts = h->nelts;
  return SHQ_OK;
}
int shq_rounds_done(shq_t h, uint32_t* rounds) { if (!h || !rounds) return SHQ_ERR_ARG; *rounds = h->rounds_done; return SHQ_OK; }

int shq_denoise(shq_t h, uint64_t* removed) {
  if (!h) return SHQ_ERR_ARG;
  struct { uint64_t* r; } args{removed};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    uint64_t r = denoise_round(h);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    if (a.r) *a.r = r;
  SHQ_API_END
}

int shq_geometry(shq_t h, int* qb, int* hb, uint64_t* nslots, uint64_t* xnslots, uint64_t* nblocks) {
  if (!h) return SHQ_ERR_ARG;
  if (qb) *qb = h->qb; if (hb) *hb = h->hb;
  if (nslots) *nslots = h->nslots; if (xnslots) *xnslots = h->xnslots; if (nblocks) *nblocks = h->nblocks;
  return SHQ_OK;
}
int shq_image_size(shq_t h, size_t* bytes) { if (!h || !bytes) return SHQ_ERR_ARG; *bytes = 128 + (size_t)h->nblocks * kBlockBytes; return SHQ_OK; }

int shq_export_cqf(shq_t h, void* dst, size_t cap, size_t* written) {
  if (!h || !dst) return SHQ_ERR_ARG;
  struct { void* dst; size_t cap; size_t* w; } args{dst, cap, written};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    size_t total = 128 + (size_t)h->nblocks * kBlockBytes;
    if (a.cap < total) SHQ_FAIL(SHQ_ERR_ARG, "buffer of %zu bytes < image of %zu", a.cap, total);
    if (h->nshards != 1) SHQ_FAIL(SHQ_ERR_STATE, "export of a sharded handle goes through the multi-GPU layer");
    if (!h->image_only) materialise(h);
    uint8_t* out = (uint8_t*)a.dst;
    build_header(h, out);
    SHQ_CUDA(cudaMemcpyAsync(out + 128, h->image.p, total - 128, cudaMemcpyDeviceToHost, h->stream));
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += total - 128;
    if (a.w) *a.w = total;
  SHQ_API_END
}

int shq_save(shq_t h, const char* path) {
  if (!h || !path) return SHQ_ERR_ARG;
  struct { const char* path; } args{path};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    if (h->nshards != 1) SHQ_FAIL(SHQ_ERR_STATE, "export of a sharded handle goes through the multi-GPU layer");
    if (!h->image_only) materialise(h);
    // the image leaves the device through two page-locked buffers: the copy of piece i+1 runs under the write of piece i
    PinnedPair pp;                       // releases file, buffers and events on every exit path
    pp.f = fopen(a.path, "wb+");
    if (!pp.f) SHQ_FAIL(SHQ_ERR_IO, "cannot write %s", a.path);
    uint8_t hd[128];
    build_header(h, hd);
    bool ok = fwrite(hd, 1, 128, pp.f) == 128;
    const size_t total = (size_t)h->nblocks * kBlockBytes, piece = (size_t)32 << 20;
    pp.alloc(piece);
    size_t off = 0; int cur = 0;
    auto issue = [&](int b, size_t o) {
      size_t m = std::min(piece, total - o);
      SHQ_CUDA(cudaMemcpyAsync(pp.pin[b], h->image.as<uint8_t>() + o, m, cudaMemcpyDeviceToHost, h->stream));
      SHQ_CUDA(cudaEventRecord(pp.ev[b], h->stream));
    };
    if (total) issue(0, 0);
    while (off < total && ok) {
      size_t m = std::min(piece, total - off);
      if (off + m < total) issue(cur ^ 1, off + m);
      SHQ_CUDA(cudaEventSynchronize(pp.ev[cur]));
      ok = fwrite(pp.pin[cur], 1, m, pp.f) == m;
      off += m; cur ^= 1;
    }
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += total;
    if (pp.close_file() != 0) ok = false;
    if (!ok) SHQ_FAIL(SHQ_ERR_IO, "cannot write %s", a.path);
  SHQ_API_END
}

int shq_load_image(const void* image, size_t n, int device, shq_t* out) {
  if (!image || !out || n < 128) return SHQ_ERR_FORMAT;
  *out = nullptr;
  const uint8_t* p = (const uint8_t*)image;
  auto get = [&](int off) { uint64_t v; memcpy(&v, p + off, 8); return v; };
  uint64_t size = get(0), nslots = get(16), xnslots = get(24), key_bits = get(32), bps = get(56), nblocks = get(80);
  uint32_t seed; memcpy(&seed, p + 8, 4);
  int qb = 0; while (qb < 63 && (1ULL << qb) < nslots) qb++;
  if (bps != 8 || (1ULL << qb) != nslots || key_bits != (uint64_t)qb + 8 || size != nblocks * kBlockBytes || n < 128 + size)
    return SHQ_ERR_FORMAT;
  shq_t h = nullptr;
  int rc = shq_create(qb, (int)key_bits, seed, device, &h);
  if (rc) return rc;
  if (h->xnslots != xnslots || h->nblocks != nblocks) { shq_destroy(h); return SHQ_ERR_FORMAT; }
  struct { const uint8_t* p; size_t size; } args{p, (size_t)size};
  rc = guard(h, [](shq_handle_s* h, void* ap) {
    auto& a = *(decltype(args)*)ap;
    size_t tiles = (h->nblocks + kTileBlocks - 1) / kTileBlocks;
    h->image.reserve(tiles * kTileBytes + 16);
    SHQ_CUDA(cudaMemcpyAsync(h->image.p, a.p + 128, a.size, cudaMemcpyHostToDevice, h->stream));
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.h2d_bytes += a.size;
  }, &args);
  if (rc) { shq_destroy(h); return rc; }
  h->nelts = get(88); h->ndistinct = get(96);
  h->image_valid = true; h->image_only = true;
  *out = h;
  return SHQ_OK;
}

// header checks shared by shq_load_image / shq_load: geometry out of the 128-byte qfmetadata
static int parse_header(const uint8_t* p, size_t avail, int* qb, int* hb, uint32_t* seed, uint64_t* size, uint64_t* nelts, uint64_t* ndistinct,
                        uint64_t* xnslots_out, uint64_t* nblocks_out) {
  auto get = [&](int off) { uint64_t v; memcpy(&v, p + off, 8); return v; };
  uint64_t sz = get(0), nslots = get(16), xnslots = get(24), key_bits = get(32), bps = get(56), nblocks = get(80);
  memcpy(seed, p + 8, 4);
  int q = 0; while (q < 63 && (1ULL << q) < nslots) q++;
  if (bps != 8 || (1ULL << q) != nslots || key_bits != (uint64_t)q + 8 || sz != nblocks * kBlockBytes || avail < 128 + sz) return SHQ_ERR_FORMAT;
  *qb = q; *hb = (int)key_bits; *size = sz; *nelts = get(88); *ndistinct = get(96); *xnslots_out = xnslots; *nblocks_out = nblocks;
  return SHQ_OK;
}

// qf_deserialize (cqf/gqf.c:2396-2420) without a whole-file host copy: the block array streams through two page-locked
// pieces, the read of piece i+1 overlapping the H2D copy of piece i (a 22 GiB human table never sits in pageable memory)
int shq_load(const char* path, int device, shq_t* out) {
  if (!path || !out) return SHQ_ERR_ARG;
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return SHQ_ERR_IO;
  uint8_t hd[128];
  if (fread(hd, 1, 128, f) != 128) { fclose(f); return SHQ_ERR_FORMAT; }
  if (fseeko(f, 0, SEEK_END) != 0) { fclose(f); return SHQ_ERR_IO; }
  const off_t fsz = ftello(f);
  if (fseeko(f, 128, SEEK_SET) != 0) { fclose(f); return SHQ_ERR_IO; }
  int qb, hb; uint32_t seed; uint64_t size, nelts, ndistinct, xnslots, nblocks;
  int rc = parse_header(hd, fsz < 0 ? 0 : (size_t)fsz, &qb, &hb, &seed, &size, &nelts, &ndistinct, &xnslots, &nblocks);
  if (rc) { fclose(f); return rc; }
  shq_t h = nullptr;
  rc = shq_create(qb, hb, seed, device, &h);
  if (rc) { fclose(f); return rc; }
  if (h->xnslots != xnslots || h->nblocks != nblocks) { fclose(f); shq_destroy(h); return SHQ_ERR_FORMAT; }
  struct { FILE* f; size_t size; } args{f, (size_t)size};
  rc = guard(h, [](shq_handle_s* h, void* ap) {
    auto& a = *(decltype(args)*)ap;
    PinnedPair pp; pp.f = a.f;
    const size_t piece = (size_t)32 << 20;
    pp.alloc(piece);
    size_t tiles = (h->nblocks + kTileBlocks - 1) / kTileBlocks;
    h->image.reserve(tiles * kTileBytes + 16);
    int cur = 0; bool used[2] = {false, false};
    for (size_t off = 0; off < a.size; off += piece, cur ^= 1) {
      const size_t m = std::min(piece, a.size - off);
      if (used[cur]) SHQ_CUDA(cudaEventSynchronize(pp.ev[cur]));       // the copy that last read this buffer is done
      if (fread(pp.pin[cur], 1, m, pp.f) != m) SHQ_FAIL(SHQ_ERR_IO, "short read at image offset %zu", off);
      SHQ_CUDA(cudaMemcpyAsync(h->image.as<uint8_t>() + off, pp.pin[cur], m, cudaMemcpyHostToDevice, h->stream));
      SHQ_CUDA(cudaEventRecord(pp.ev[cur], h->stream));
      used[cur] = true;
    }
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.h2d_bytes += a.size;
  }, &args);
  if (rc) { shq_destroy(h); return rc; }
  memcpy(h->header, hd, 128); h->has_header = true;
  h->nelts = nelts; h->ndistinct = ndistinct;
  h->image_valid = true; h->image_only = true;
  *out = h;
  return SHQ_OK;
}

int shq_iterate(shq_t h, uint64_t* keys, uint64_t* counts, size_t cap, size_t* n) {
  if (!h || !n) return SHQ_ERR_ARG;
  struct { uint64_t* k; uint64_t* c; size_t cap; size_t* n; } args{keys, counts, cap, n};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    ensure_state(h);
    *a.n = h->n_state;
    size_t m = std::min(a.cap, h->n_state);
    if (m && a.k) { SHQ_CUDA(cudaMemcpyAsync(a.k, h->st_keys.p, m * 8, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += m * 8; }
    if (m && a.c) {
      std::vector<uint32_t> tmp(m);
      SHQ_CUDA(cudaMemcpyAsync(tmp.data(), h->st_counts.p, m * 4, cudaMemcpyDeviceToHost, h->stream));
      SHQ_CUDA(cudaStreamSynchronize(h->stream));
      h->stats.d2h_bytes += m * 4;
      for (size_t i = 0; i < m; i++) a.c[i] = tmp[i];
    }
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}

static int query_impl(shq_t h, const uint64_t* keys, size_t n, uint64_t* counts, uint8_t* was, bool set) {
  if (!h || (!keys && n) || (!counts && n)) return SHQ_ERR_ARG;
  struct { const uint64_t* k; size_t n; uint64_t* c; uint8_t* w; bool set; } args{keys, n, counts, was, set};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    if (a.n == 0) return;
    if (!h->image_only) materialise(h);
    ImageView im{h->image.as<uint8_t>(), h->nblocks, h->xnslots, 0};
    h->d_q_in.reserve(a.n * 8); h->d_q_cnt.reserve(a.n * 8); h->d_q_was.reserve(a.n + 16);
    SHQ_CUDA(cudaMemcpyAsync(h->d_q_in.p, a.k, a.n * 8, cudaMemcpyHostToDevice, h->stream));
    if (a.set) query_kernel<true><<<cdiv(a.n, 128), 128, 0, h->stream>>>(im, h->d_q_in.as<uint64_t>(), a.n, h->d_q_cnt.as<uint64_t>(), h->d_q_was.as<uint8_t>());
    else query_kernel<false><<<cdiv(a.n, 128), 128, 0, h->stream>>>(im, h->d_q_in.as<uint64_t>(), a.n, h->d_q_cnt.as<uint64_t>(), nullptr);
    h->stats.kernels_launched++;
    SHQ_CUDA(cudaGetLastError());
    SHQ_CUDA(cudaMemcpyAsync(a.c, h->d_q_cnt.p, a.n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (a.set && a.w) SHQ_CUDA(cudaMemcpyAsync(a.w, h->d_q_was.p, a.n, cudaMemcpyDeviceToHost, h->stream));
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.h2d_bytes += a.n * 8; h->stats.d2h_bytes += a.n * 9;
  SHQ_API_END
}
int shq_query_set_traveled(shq_t h, const uint64_t* keys, size_t n, uint64_t* counts, uint8_t* was) { return query_impl(h, keys, n, counts, was, true); }
int shq_query(shq_t h, const uint64_t* keys, size_t n, uint64_t* counts) { return query_impl(h, keys, n, counts, nullptr, false); }

int shq_hash_fastq_chunk(shq_t h, const char* text, size_t n, int K, uint64_t* keys, size_t cap, size_t* nkeys) {
  if (!h || !nkeys || (!text && n)) return SHQ_ERR_ARG;
  struct { const char* t; size_t n; int K; uint64_t* keys; size_t cap; size_t* nk; } args{text, n, K, keys, cap, nkeys};
  SHQ_API_BEGIN(h)
    cudaStream_t st = h->stream;
    *a.nk = 0;
    if (a.n == 0) return;
    if (a.t[a.n - 1] != '\n') SHQ_FAIL(SHQ_ERR_MALFORMED, "chunk has no trailing newline");
    h->d_text.reserve(a.n + 64);
    SHQ_CUDA(cudaMemcpyAsync(h->d_text.p, a.t, a.n, cudaMemcpyHostToDevice, st));
    size_t slots_cap = a.n;
    h->d_q_cnt.reserve(slots_cap * 8 + 64);
    size_t slots = hash_to_keylist(h, h->d_text.as<char>(), a.n, a.K, h->d_q_cnt.as<uint64_t>(), slots_cap);
    std::vector<uint64_t> tmp(slots);
    if (slots) SHQ_CUDA(cudaMemcpyAsync(tmp.data(), h->d_q_cnt.p, slots * 8, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    size_t m = 0;
    for (size_t i = 0; i < slots; i++) if (tmp[i] != ~0ULL) { if (m < a.cap && a.keys) a.keys[m] = tmp[i]; m++; }
    *a.nk = m;
    h->stats.h2d_bytes += a.n; h->stats.d2h_bytes += slots * 8;
  SHQ_API_END
}

int shq_materialize(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  int args = 0;
  SHQ_API_BEGIN(h)
    (void)a;
    flush_stage(h);
    if (!h->image_only) materialise(h);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}

int shq_hash_fastq_device(shq_t h, const char* d_text, size_t n, int K, uint64_t* d_keys, size_t cap, size_t* nkeys) {
  if (!h || !nkeys || (!d_text && n)) return SHQ_ERR_ARG;
  struct { const char* t; size_t n; int K; uint64_t* keys; size_t cap; size_t* nk; } args{d_text, n, K, d_keys, cap, nkeys};
  SHQ_API_BEGIN(h)
    *a.nk = 0;
    if (a.n == 0) return;
    *a.nk = hash_to_keylist(h, a.t, a.n, a.K, a.keys, a.cap);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}

int shq_partition_keys_device(shq_t h, const uint64_t* d_in, size_t n, int nparts, uint64_t* d_out, uint64_t* counts) {
  if (!h || !counts || nparts < 1 || nparts > 16 || (nparts & (nparts - 1)) || (n && (!d_in || !d_out))) return SHQ_ERR_ARG;
  struct { const uint64_t* in; size_t n; int np; uint64_t* out; uint64_t* counts; } args{d_in, n, nparts, d_out, counts};
  SHQ_API_BEGIN(h)
    cudaStream_t st = h->stream;
    uint64_t& L = h->stats.kernels_launched;
    for (int i = 0; i < a.np; i++) a.counts[i] = 0;
    if (a.n == 0) return;
    int pbits = 0; while ((1 << pbits) < a.np) pbits++;
    int shift = h->hb - pbits;
    size_t nthreads = (a.n + kOwnerStrip - 1) / kOwnerStrip;
    size_t cells = nthreads * a.np;
    h->d_tmp_keys.reserve((cells + 4) * 8);
    uint64_t* mat = h->d_tmp_keys.as<uint64_t>();
    owner_partition_kernel<false><<<cdiv(nthreads, 128), 128, 0, st>>>(a.in, a.n, shift, a.np, nthreads, mat, nullptr); L++;
    exclusive_scan_inplace<SumU64>(mat, cells, h->d_scan_scratch, st, mat + cells, &L);
    owner_partition_kernel<true><<<cdiv(nthreads, 128), 128, 0, st>>>(a.in, a.n, shift, a.np, nthreads, mat, a.out); L++;
    SHQ_CUDA(cudaGetLastError());
    // owner o starts at mat[o * nthreads]
    std::vector<uint64_t> starts(a.np + 1);
    for (int o = 0; o < a.np; o++) SHQ_CUDA(cudaMemcpyAsync(&starts[o], mat + (size_t)o * nthreads, 8, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    starts[a.np] = a.n;
    for (int o = 0; o < a.np; o++) a.counts[o] = starts[o + 1] - starts[o];
  SHQ_API_END
}

int shq_mg_bin_fastq_device(shq_t h, const char* d_text, size_t n, int K, int coarse_bits, const uint32_t** d_words,
                            const uint32_t** d_bin_off, size_t* n_words) {
  if (!h || (!d_text && n) || !d_words || !d_bin_off || !n_words || coarse_bits < 1 || coarse_bits > 12) return SHQ_ERR_ARG;
  struct { const char* t; size_t n; int K, cb; const uint32_t** w; const uint32_t** o; size_t* nw; } args{d_text, n, K, coarse_bits, d_words, d_bin_off, n_words};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    *a.nw = 0; *a.w = nullptr; *a.o = nullptr;
    if (a.n == 0) return;
    const char* src = a.t;
    uint64_t end = a.n;
    BatchBins bb;
    bin_batch(h, src, a.n, &end, 1, a.K, nullptr, 0, false, &bb, a.cb);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.batches++;
    *a.nw = bb.n_keys; *a.w = h->d_packed.as<uint32_t>(); *a.o = h->d_binoff.as<uint32_t>();
  SHQ_API_END
}

int shq_mg_insert_binned(shq_t h, const uint32_t* d_recv, const uint64_t* seg_start, const uint32_t* d_counts, int world,
                         int coarse_bits, int fine_bits, size_t n_words) {
  if (!h || world < 1 || world > 64 || !seg_start || (n_words && (!d_recv || !d_counts)) || fine_bits < 0 || fine_bits > 5) return SHQ_ERR_ARG;
  struct { const uint32_t* r; const uint64_t* ss; const uint32_t* c; int world, cb, fb; size_t n; } args{d_recv, seg_start, d_counts, world, coarse_bits, fine_bits, n_words};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    ensure_state(h);
    if (a.n == 0) return;
    if (a.world != h->nshards) SHQ_FAIL(SHQ_ERR_ARG, "world %d != shards of this handle %d", a.world, h->nshards);
    if (a.n >= ((size_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "more than 2^32 words in one exchange batch");
    cudaStream_t st = h->stream;
    uint64_t& L = h->stats.kernels_launched;
    int shard_bits = 0; while ((1 << shard_bits) < h->nshards) shard_bits++;
    const int cps = 1 << (a.cb - shard_bits);            // coarse bins per shard
    const int lbc = h->hb - a.cb, lbf = lbc - a.fb;
    if (a.cb < shard_bits || lbf < 1 || lbc > 31) SHQ_FAIL(SHQ_ERR_ARG, "bad exchange geometry (coarse %d, fine %d bits)", a.cb, a.fb);
    const int nfine = cps << a.fb;
    BatchBins bb;
    {
      Phase ph(h, &h->stats.ms_scatter);
      h->d_seg.reserve(((size_t)a.world * (cps + 1) + a.world + 8) * 8);
      uint64_t* d_seg_start = h->d_seg.as<uint64_t>();
      uint64_t* seg_beg = d_seg_start + a.world + 4;
      SHQ_CUDA(cudaMemcpyAsync(d_seg_start, a.ss, (size_t)a.world * 8, cudaMemcpyHostToDevice, st));
      seg_table_kernel<<<a.world, kScanThreads, 0, st>>>(a.c, d_seg_start, a.world, cps, seg_beg); L++;
      h->d_fine.reserve(((size_t)nfine + 2) * 4);
      h->d_binoff.reserve(((size_t)nfine + 2) * 4);
      SHQ_CUDA(cudaMemsetAsync(h->d_fine.p, 0, ((size_t)nfine + 1) * 4, st));
      int parts = 1; while (cps * parts < 4 * kNumSMs) parts *= 2;
      h->d_bases.reserve((size_t)a.world * 8 + 64);
      { std::vector<const uint32_t*> b((size_t)a.world, a.r);
        SHQ_CUDA(cudaMemcpyAsync(h->d_bases.p, b.data(), (size_t)a.world * 8, cudaMemcpyHostToDevice, st));
        SHQ_CUDA(cudaStreamSynchronize(st)); }
      const uint32_t* const* bases = h->d_bases.as<const uint32_t*>();
      resplit_kernel<true><<<cps * parts, kSplitThreads, 0, st>>>(bases, seg_beg, a.world, cps, parts, lbc, a.fb, h->d_fine.as<uint32_t>(), nullptr); L++;
      SHQ_CUDA(cudaMemcpyAsync(h->d_binoff.p, h->d_fine.p, (size_t)nfine * 4, cudaMemcpyDeviceToDevice, st));
      exclusive_scan_inplace<SumU32>(h->d_binoff.as<uint32_t>(), nfine, h->d_scan_scratch, st, h->d_binoff.as<uint32_t>() + nfine, &L);
      SHQ_CUDA(cudaMemcpyAsync(h->d_fine.p, h->d_binoff.p, (size_t)nfine * 4, cudaMemcpyDeviceToDevice, st));
      h->d_packed.reserve((a.n + 4) * 4);
      resplit_kernel<false><<<cps * parts, kSplitThreads, 0, st>>>(bases, seg_beg, a.world, cps, parts, lbc, a.fb, h->d_fine.as<uint32_t>(),
                                                                 h->d_packed.as<uint32_t>()); L++;
      SHQ_CUDA(cudaGetLastError());
    }
    bb.n_bins = nfine; bb.lb = lbf; bb.n_keys = (uint32_t)a.n; bb.armed = false; bb.nchunks = 1;
    bb.words = h->d_packed.as<uint32_t>(); bb.off = h->d_binoff.as<uint32_t>(); bb.end = nullptr;
    Run A{&h->d_run_keys, &h->d_run_counts, &h->d_run_chunk};
    aggregate_bins(h, bb, 0, 0xFFFFFFFFu, false, 0, &A, nullptr, nullptr);
    if (A.occ != a.n) SHQ_FAIL(SHQ_ERR_STATE, "exchange batch: %llu words aggregated, %zu received", (unsigned long long)A.occ, a.n);
    uint64_t n_new = merge_run(h, A);
    h->ndistinct += n_new; h->nelts += A.occ;
    check_err_flag(h);
  SHQ_API_END
}

// ---- pull exchange: owners read their coarse bins out of the sources' binned arrays (peer memory over NVLink) ----
constexpr int kMgCoarseBits = 11;      // coarse bins over the whole key space (level-1 scatter)
constexpr int kMgFineBits = 12;        // fine bins per shard (what the aggregator wants)

int shq_mg_setup(shq_t h, size_t max_words, void* ipc_handles_out) {
  if (!h || !ipc_handles_out || max_words == 0 || max_words >= ((size_t)1 << 32)) return SHQ_ERR_ARG;
  struct { size_t n; void* out; } args{max_words, ipc_handles_out};
  SHQ_API_BEGIN(h)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int i = 0; i < 2; i++) {
      if (h->mg_send[i]) { cudaFree(h->mg_send[i]); h->mg_send[i] = nullptr; }
      SHQ_CUDA(cudaMalloc((void**)&h->mg_send[i], (a.n + 64) * 4));
      cudaIpcMemHandle_t hd;
      SHQ_CUDA(cudaIpcGetMemHandle(&hd, h->mg_send[i]));
      memcpy((char*)a.out + 64 * i, &hd, 64);
    }
    h->mg_cap_words = a.n;
  SHQ_API_END
}

int shq_mg_open_peers(shq_t h, const void* all_handles, int world) {
  if (!h || !all_handles || world != h->nshards) return SHQ_ERR_ARG;
  struct { const void* hs; int world; } args{all_handles, world};
  SHQ_API_BEGIN(h)
    if (!h->mg_send[0]) SHQ_FAIL(SHQ_ERR_STATE, "shq_mg_setup first");
    for (void* p : h->mg_opened) cudaIpcCloseMemHandle(p);
    h->mg_opened.clear();
    h->mg_peer.assign((size_t)a.world * 2, nullptr);
    for (int r = 0; r < a.world; r++) for (int i = 0; i < 2; i++) {
      if (r == h->shard) { h->mg_peer[r * 2 + i] = h->mg_send[i]; continue; }
      cudaIpcMemHandle_t hd;
      memcpy(&hd, (const char*)a.hs + ((size_t)r * 2 + i) * 64, 64);
      void* p = nullptr;
      SHQ_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      h->mg_opened.push_back(p);
      h->mg_peer[r * 2 + i] = (const uint32_t*)p;
    }
  SHQ_API_END
}

int shq_mg_bin_local(shq_t h, const char* d_text, size_t n, int K, int parity, uint32_t* d_fine_hist, size_t* n_words) {
  if (!h || (!d_text && n) || !d_fine_hist || !n_words || parity < 0 || parity > 1) return SHQ_ERR_ARG;
  struct { const char* t; size_t n; int K, par; uint32_t* fh; size_t* nw; } args{d_text, n, K, parity, d_fine_hist, n_words};
  SHQ_API_BEGIN(h)
    // Bins this rank's text over the WHOLE key space into 2^cb coarse bins inside its exchange buffer and publishes the
    // layout: meta[b] = first word of coarse bin b, meta[2^cb + b] = one past its last (2^(cb+1) words; the caller
    // all-gathers them).  One hashing pass into fixed-capacity bins; if a bin overflows (skewed keys) the batch is binned
    // again with a counting pass and exact offsets -- the owners only ever see (begin, end) pairs.
    flush_stage(h);
    if (!h->mg_send[0]) SHQ_FAIL(SHQ_ERR_STATE, "shq_mg_setup first");
    cudaStream_t st = h->stream;
    uint64_t& L = h->stats.kernels_launched;
    *a.nw = 0;
    int shard_bits = 0; while ((1 << shard_bits) < h->nshards) shard_bits++;
    const int cb = kMgCoarseBits > shard_bits ? kMgCoarseBits : shard_bits + 1;
    const int ncoarse = 1 << cb;
    const int lbc = h->hb - cb;
    if (lbc < 1 || lbc > 31) SHQ_FAIL(SHQ_ERR_ARG, "hb=%d does not fit the exchange geometry of %d shards", h->hb, h->nshards);
    SHQ_CUDA(cudaMemsetAsync(a.fh, 0, (size_t)ncoarse * 8, st));
    if (a.n == 0) { SHQ_CUDA(cudaStreamSynchronize(st)); return; }
    if (a.K < 1 || a.K > 1024) SHQ_FAIL(SHQ_ERR_ARG, "K=%d outside [1,1024]", a.K);
    if (a.n >= ((size_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "batch text of %zu bytes exceeds 4 GiB", a.n);
    const char* src = a.t;
    uint32_t n_lines = index_lines(h, src, a.n);
    h->d_chunk_ends.reserve(16); h->d_chunk_line_end.reserve(16);
    uint64_t end = a.n;
    SHQ_CUDA(cudaMemcpyAsync(h->d_chunk_ends.p, &end, 8, cudaMemcpyHostToDevice, st));
    chunk_validate_kernel<<<1, 128, 0, st>>>(src, h->d_nl_pos.as<uint32_t>(), n_lines, h->d_chunk_ends.as<uint64_t>(), 1, h->d_chunk_line_end.as<uint32_t>(), h->d_err); L++;
    if (n_lines & 3) SHQ_FAIL(SHQ_ERR_MALFORMED, "batch holds %u lines, not a multiple of 4", n_lines);
    HashParams hp{};
    hp.text = src; hp.nl_pos = h->d_nl_pos.as<uint32_t>(); hp.n_reads = n_lines / 4; hp.K = a.K;
    hp.hb = h->hb; hp.key_lo = 0; hp.key_hi = 1ULL << h->hb; hp.nchunks = 1;
    hp.lb = lbc; hp.n_bins = ncoarse;
    h->d_hist.reserve((size_t)(ncoarse + 1) * 4); h->d_binoff.reserve((size_t)(ncoarse + 2) * 4); h->d_tot.reserve((size_t)(ncoarse + 2) * 4);
    h->d_misc.reserve(256);
    unsigned long long* d_ub = h->d_misc.as<unsigned long long>() + 12;       // [12] window total, [13] words held
    int* d_ovf = h->d_misc.as<int>() + 20;
    const size_t smem_scatter = sizeof(RollTables) + ((size_t)3 * ncoarse + 1) * 4 + (size_t)kScatThreads * 16 * 14 + 16;
    SHQ_CUDA(cudaFuncSetAttribute(roll_scatter_kernel<kScatThreads, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
    unsigned long long n_keys = 0;
    bool done = false;
    if (h->spec_bins) {
      SHQ_CUDA(cudaMemsetAsync(d_ub, 0, 16, st));
      SHQ_CUDA(cudaMemsetAsync(d_ovf, 0, 4, st));
      window_total_kernel<<<kNumSMs * 4, 256, 0, st>>>(hp.nl_pos, hp.n_reads, a.K, d_ub); L++;
      unsigned long long n_ub = 0; int dev_err = 0;
      read_pair(h, d_ub, &n_ub, h->d_err, &dev_err);
      if (dev_err) { SHQ_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), st)); SHQ_FAIL(dev_err, "device reported: %s", shq_strerror(dev_err)); }
      const uint64_t capbin = bin_capacity(h, n_ub / (uint64_t)ncoarse);
      const uint64_t span = capbin * (uint64_t)ncoarse, round = (uint64_t)kScatThreads * 16;
      if (n_ub > 0 && span + round + 64 <= h->mg_cap_words) {
        fixed_bin_offsets_kernel<<<cdiv(ncoarse + 1, 256), 256, 0, st>>>(h->d_binoff.as<uint32_t>(), h->d_tot.as<uint32_t>(), ncoarse, (uint32_t)capbin); L++;
        {
          Phase ph(h, &h->stats.ms_scatter);
          roll_scatter_kernel<kScatThreads, 16, true><<<kNumSMs, kScatThreads, smem_scatter, st>>>(hp, h->d_tot.as<uint32_t>(), h->mg_send[a.par],
                                                                                               (uint32_t)capbin, (uint32_t)span, d_ovf); L++;
        }
        layout_meta_kernel<<<cdiv(ncoarse, 256), 256, 0, st>>>(h->d_binoff.as<uint32_t>(), h->d_tot.as<uint32_t>(), ncoarse, a.fh, d_ub + 1); L++;
        SHQ_CUDA(cudaGetLastError());
        int ovf = 0;
        read_pair(h, d_ub + 1, &n_keys, d_ovf, &ovf);      // also: the buffer is complete once this returns
        if (!ovf) done = true; else h->stats.bin_fallbacks++;
      } else if (n_ub == 0) {
        done = true;
      }
    }
    if (!done) {
      {
        Phase ph(h, &h->stats.ms_hash_count);
        SHQ_CUDA(cudaMemsetAsync(h->d_hist.p, 0, (size_t)(ncoarse + 1) * 4, st));
        size_t smem_count = sizeof(RollTables) + (size_t)ncoarse * 4;
        SHQ_CUDA(cudaFuncSetAttribute(roll_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
        roll_count_kernel<true><<<kNumSMs * 2, kRollThreads, smem_count, st>>>(hp, h->d_hist.as<uint32_t>()); L++;
        SHQ_CUDA(cudaMemcpyAsync(h->d_binoff.p, h->d_hist.p, (size_t)ncoarse * 4, cudaMemcpyDeviceToDevice, st));
        exclusive_scan_inplace<SumU32>(h->d_binoff.as<uint32_t>(), ncoarse, h->d_scan_scratch, st, h->d_binoff.as<uint32_t>() + ncoarse, &L);
        SHQ_CUDA(cudaMemcpyAsync(h->d_tot.p, h->d_binoff.p, (size_t)ncoarse * 4, cudaMemcpyDeviceToDevice, st));
      }
      uint32_t nk = 0; int dev_err = 0;
      read_pair(h, h->d_binoff.as<uint32_t>() + ncoarse, &nk, h->d_err, &dev_err);
      if (dev_err) { SHQ_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), st)); SHQ_FAIL(dev_err, "device reported: %s", shq_strerror(dev_err)); }
      if (nk > h->mg_cap_words) SHQ_FAIL(SHQ_ERR_CAPACITY, "%u k-mers in one batch, exchange buffers hold %zu", nk, h->mg_cap_words);
      {
        Phase ph(h, &h->stats.ms_scatter);
        roll_scatter_kernel<kScatThreads, 16, true><<<kNumSMs, kScatThreads, smem_scatter, st>>>(hp, h->d_tot.as<uint32_t>(), h->mg_send[a.par]); L++;
      }
      SHQ_CUDA(cudaMemsetAsync(d_ub + 1, 0, 8, st));
      layout_meta_kernel<<<cdiv(ncoarse, 256), 256, 0, st>>>(h->d_binoff.as<uint32_t>(), nullptr, ncoarse, a.fh, d_ub + 1); L++;
      SHQ_CUDA(cudaGetLastError());
      SHQ_CUDA(cudaStreamSynchronize(st));
      n_keys = nk;
    }
    h->stats.batches++; h->stats.keys_hashed += n_keys;
    h->stats.scatter_launches++; h->stats.scatter_text_bytes += a.n; h->stats.scatter_words += n_keys;
    *a.nw = (size_t)n_keys;
  SHQ_API_END
}

// Owner side of the pull exchange: from the all-gathered layouts (row r = shq_mg_bin_local's metadata of rank r, row
// stride 2^(shard_bits + kMgFineBits) words) pull my coarse bins out of every source's buffer and split them into my
// kMgFineBits fine bins in one kernel.  The fine bins have a fixed capacity (the total is known, ntHash keys are uniform);
// exact = true runs a counting pass over the same peer memory first and uses exact offsets instead.
// Returns the number of words pulled; bb describes the fine bins.
static uint64_t mg_pull_split(shq_handle_s* h, const uint32_t* d_all_meta, int parity, bool exact, BatchBins* bb) {
  cudaStream_t st = h->stream;
  uint64_t& L = h->stats.kernels_launched;
  const int world = h->nshards;
  int shard_bits = 0; while ((1 << shard_bits) < world) shard_bits++;
  const int cb = kMgCoarseBits > shard_bits ? kMgCoarseBits : shard_bits + 1;
  const int gfb = shard_bits + kMgFineBits;
  const int nfg = 1 << gfb, ncoarse = 1 << cb, cps = 1 << (cb - shard_bits);
  const int lbc = h->hb - cb;
  int fb = gfb - cb;                           // kMgFineBits fine bins per shard ...
  if (lbc - fb < 1 || lbc > 31 || fb < 0 || fb > 5) SHQ_FAIL(SHQ_ERR_ARG, "hb=%d does not fit the exchange geometry of %d shards", h->hb, h->nshards);
  Phase ph(h, &h->stats.ms_split);
  h->d_bases.reserve((size_t)world * 8 + 64);
  std::vector<const uint32_t*> b((size_t)world);
  for (int r = 0; r < world; r++) b[r] = h->mg_peer[r * 2 + parity];
  SHQ_CUDA(cudaMemcpyAsync(h->d_bases.p, b.data(), (size_t)world * 8, cudaMemcpyHostToDevice, st));
  h->d_seg.reserve(((size_t)world * cps * 2 + 8) * 8);
  uint64_t* seg_beg = h->d_seg.as<uint64_t>(); uint64_t* seg_end = seg_beg + (size_t)world * cps;
  h->d_misc.reserve(256);
  unsigned long long* d_tot = h->d_misc.as<unsigned long long>() + 14;
  int* d_ovf = h->d_misc.as<int>() + 20;
  SHQ_CUDA(cudaMemsetAsync(d_tot, 0, 8, st));
  SHQ_CUDA(cudaMemsetAsync(d_ovf, 0, 4, st));
  pull_layout_kernel<<<cdiv(world * cps, 256), 256, 0, st>>>(d_all_meta, (size_t)nfg, world, h->shard, cps, ncoarse, seg_beg, seg_end, d_tot); L++;
  const uint64_t n = read_scalar(h, (const uint64_t*)d_tot);
  // ... or more, when the distinct ratio of the previous batch says that 4096 bins would each hold more keys than the
  // aggregator's table takes: the pull then splits all the way (up to 128 fine bins per coarse bin) and the second
  // split pass over the received words is not needed
  if (h->ratio_seen && h->fine_bits_forced < 0)
    while (fb < 7 && lbc - fb > 2 && (double)n * std::min(1.0, h->distinct_ratio * 1.1) / (double)((uint64_t)cps << fb) > h->fine_target) fb++;   // refine_bins' rule
  const int F = cps << fb, lbf = lbc - fb;
  bb->n_bins = F; bb->lb = lbf; bb->n_keys = (uint32_t)n; bb->armed = false; bb->nchunks = 1; bb->fine_bits = 0;
  bb->spec = false; bb->level1_spec = false; bb->span = 0; bb->end = nullptr;
  if (n == 0) return 0;
  if (n >= ((uint64_t)1 << 32)) SHQ_FAIL(SHQ_ERR_ARG, "%llu words in one exchange batch", (unsigned long long)n);
  h->d_fine.reserve(((size_t)F + 2) * 4); h->d_binoff.reserve(((size_t)F + 2) * 4);
  int parts = 1; while (cps * parts < 4 * kNumSMs) parts *= 2;
  const uint64_t capfine = bin_capacity(h, n / (uint64_t)F);
  const uint64_t span = capfine * (uint64_t)F;
  if (!exact && span + kSplitTile + 64 < ((uint64_t)1 << 32)) {
    h->d_packed.reserve((size_t)(span + kSplitTile + 64) * 4);
    fine_bin_offsets_kernel<<<cdiv(F + 1, 256), 256, 0, st>>>(h->d_binoff.as<uint32_t>(), h->d_fine.as<uint32_t>(), (uint32_t)F, (uint32_t)capfine); L++;
    resplit_kernel<false><<<cps * parts, kSplitThreads, 0, st>>>(h->d_bases.as<const uint32_t*>(), seg_beg, world, cps, parts, lbc, fb,
                                                               h->d_fine.as<uint32_t>(), h->d_packed.as<uint32_t>(), seg_end, (uint32_t)capfine,
                                                               (uint32_t)span, d_ovf); L++;
    bb->spec = true; bb->level1_spec = true; bb->span = (uint32_t)span; bb->end = h->d_fine.as<uint32_t>();
  } else {
    SHQ_CUDA(cudaMemsetAsync(h->d_binoff.p, 0, ((size_t)F + 2) * 4, st));
    resplit_kernel<true><<<cps * parts, kSplitThreads, 0, st>>>(h->d_bases.as<const uint32_t*>(), seg_beg, world, cps, parts, lbc, fb,
                                                              h->d_binoff.as<uint32_t>(), nullptr, seg_end); L++;
    exclusive_scan_inplace<SumU32>(h->d_binoff.as<uint32_t>(), F, h->d_scan_scratch, st, h->d_binoff.as<uint32_t>() + F, &L);
    SHQ_CUDA(cudaMemcpyAsync(h->d_fine.p, h->d_binoff.p, (size_t)F * 4, cudaMemcpyDeviceToDevice, st));
    h->d_packed.reserve((n + 4) * 4);
    resplit_kernel<false><<<cps * parts, kSplitThreads, 0, st>>>(h->d_bases.as<const uint32_t*>(), seg_beg, world, cps, parts, lbc, fb,
                                                               h->d_fine.as<uint32_t>(), h->d_packed.as<uint32_t>(), seg_end); L++;
  }
  SHQ_CUDA(cudaGetLastError());
  bb->words = h->d_packed.as<uint32_t>(); bb->off = h->d_binoff.as<uint32_t>();
  h->stats.split_launches++; h->stats.split_words += n;
  return n;
}
// Pull + split + aggregate into run A.  The peers' buffers of this parity have been read completely once this returns
// (aggregate_bins ends with a host round trip): a source may refill its buffer two batches later only because every
// owner passed this point (and the all-gather in between) before.
static uint64_t mg_pull_aggregate(shq_handle_s* h, const uint32_t* d_all_meta, int parity, Run* A) {
  uint64_t n = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    BatchBins bb;
    n = mg_pull_split(h, d_all_meta, parity, attempt == 1, &bb);
    if (n == 0) { SHQ_CUDA(cudaStreamSynchronize(h->stream)); return 0; }
    const int forced = h->fine_bits_forced;
    if (attempt == 1) h->fine_bits_forced = 0;        // fixed-capacity bins overflowed: exact offsets, bins aggregated as received
    try {
      refine_bins(h, &bb, (uint32_t)std::min<uint64_t>(n, (n / (uint64_t)bb.n_bins) * 4 + 65536));
      h->fine_bits_forced = forced;
      aggregate_bins(h, bb, 0, 0xFFFFFFFFu, false, 0, A, nullptr, nullptr);
      break;
    } catch (const BinOverflow&) {
      h->fine_bits_forced = forced;
      h->stats.bin_fallbacks++;
      if (attempt == 1) SHQ_FAIL(SHQ_ERR_STATE, "binning overflow without fixed-capacity bins");
    }
  }
  if (A->occ != n) SHQ_FAIL(SHQ_ERR_STATE, "exchange batch: %llu words aggregated, %llu expected", (unsigned long long)A->occ, (unsigned long long)n);
  return n;
}

int shq_mg_insert_pull(shq_t h, const uint32_t* d_all_fine_hist, int parity) {
  if (!h || !d_all_fine_hist || parity < 0 || parity > 1) return SHQ_ERR_ARG;
  struct { const uint32_t* fh; int par; } args{d_all_fine_hist, parity};
  SHQ_API_BEGIN(h)
    flush_stage(h); ensure_state(h);
    if (h->mg_peer.empty()) SHQ_FAIL(SHQ_ERR_STATE, "shq_mg_open_peers first");
    Run A{&h->d_run_keys, &h->d_run_counts, &h->d_run_chunk};
    if (mg_pull_aggregate(h, a.fh, a.par, &A) == 0) return;
    uint64_t n_new = merge_run(h, A);
    h->ndistinct += n_new; h->nelts += A.occ;
    check_err_flag(h);
  SHQ_API_END
}

// ---- staged exchange batches + sharded deNoise: the building blocks of "the same job at every N" -----------------------
// While deNoise rounds are pending, a batch may only be committed if the GLOBAL number of distinct keys stays below
// the trigger afterwards.  shq_mg_stage_pull does everything up to that decision for this shard -- pull + split of the
// peers' binned words, aggregation, and the count of the batch's keys that are new to the shard (by a merge into the
// shadow state, or, for a batch much smaller than the state, by binary searches) -- and reports the shard's numbers.
// The caller sums them over the ranks and then calls shq_mg_commit or shq_mg_discard on every rank.
int shq_mg_stage_pull(shq_t h, const uint32_t* d_all_fine_hist, int parity, uint64_t* n_new, uint64_t* n_occ) {
  if (!h || !d_all_fine_hist || parity < 0 || parity > 1 || !n_new || !n_occ) return SHQ_ERR_ARG;
  struct { const uint32_t* fh; int par; uint64_t* nn; uint64_t* no; } args{d_all_fine_hist, parity, n_new, n_occ};
  SHQ_API_BEGIN(h)
    flush_stage(h); ensure_state(h);
    if (h->mg_peer.empty()) SHQ_FAIL(SHQ_ERR_STATE, "shq_mg_open_peers first");
    if (h->mg_pending.valid) SHQ_FAIL(SHQ_ERR_STATE, "a staged batch is waiting for shq_mg_commit / shq_mg_discard");
    cudaStream_t st = h->stream;
    uint64_t& L = h->stats.kernels_launched;
    *a.nn = 0; *a.no = 0;
    Run A{&h->d_run_keys, &h->d_run_counts, &h->d_run_chunk};
    const uint64_t n = mg_pull_aggregate(h, a.fh, a.par, &A);
    h->stats.batches++;
    h->mg_pending = {};
    h->mg_pending.valid = true;
    if (n == 0) return;
    h->mg_pending.run_n = A.n; h->mg_pending.run_occ = A.occ;
    if (h->n_state == 0) {
      h->mg_pending.n_new = A.n;
    } else if ((uint64_t)A.n * 24 < (uint64_t)h->n_state) {
      // a batch much smaller than the state (locating a deNoise boundary): count its new keys by binary search
      h->d_sample.reserve(64);
      uint32_t* d_cnt = h->d_sample.as<uint32_t>() + 8;     // [8] = n keys, [9] = new
      uint32_t hv[2] = {(uint32_t)A.n, 0};
      SHQ_CUDA(cudaMemcpyAsync(d_cnt, hv, 8, cudaMemcpyHostToDevice, st));
      count_new_kernel<<<cdiv(A.n, 256), 256, 0, st>>>(A.keys->as<uint64_t>(), d_cnt, h->st_keys.as<uint64_t>(), h->n_state, d_cnt + 1); L++;
      h->mg_pending.n_new = read_scalar(h, d_cnt + 1);
    } else {
      h->mg_pending.n_out = merge_shadow(h, A, 1, nullptr);
      h->mg_pending.has_shadow = true;
      h->mg_pending.n_new = h->mg_pending.n_out - h->n_state;
    }
    check_err_flag(h);
    *a.nn = h->mg_pending.n_new; *a.no = h->mg_pending.run_occ;
  SHQ_API_END
}

int shq_mg_commit(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  int args = 0;
  SHQ_API_BEGIN(h)
    (void)a;
    if (!h->mg_pending.valid) SHQ_FAIL(SHQ_ERR_STATE, "nothing staged");
    auto pd = h->mg_pending;
    h->mg_pending = {};
    if (pd.run_n == 0) return;
    uint64_t n_new;
    if (pd.has_shadow) n_new = commit_shadow(h, pd.n_out);
    else { Run A{&h->d_run_keys, &h->d_run_counts, &h->d_run_chunk}; A.n = pd.run_n; A.occ = pd.run_occ; n_new = merge_run(h, A); }
    if (n_new != pd.n_new) SHQ_FAIL(SHQ_ERR_STATE, "staged batch: %llu new keys counted, %llu merged", (unsigned long long)pd.n_new, (unsigned long long)n_new);
    h->ndistinct += n_new; h->nelts += pd.run_occ;
    check_err_flag(h);
  SHQ_API_END
}

int shq_mg_discard(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  h->mg_pending = {};
  return SHQ_OK;
}

// One sharded deNoise sweep (cqf/CQF_mt.h:884-922) in three calls:
//  shq_mg_denoise_walk : this shard's entries laid out behind `carry` (composition of the previous shards' layout maps,
//      see shq_mg_layout_aggregate), then this shard's part of the range walk.  state (16 words, host) comes from the
//      previous rank (zeros for rank 0) and goes on to the next one.
//  shq_mg_denoise_finalize : after the last rank, closes the range that was open at the end of all entries (host only).
//  shq_mg_denoise_apply : drops the count-1 entries except the protected ones; protect_last = bit `rank` of state[11].
int shq_mg_denoise_positions(shq_t h, const int64_t carry[2]) {
  if (!h || !carry) return SHQ_ERR_ARG;
  struct { const int64_t* c; } args{carry};
  SHQ_API_BEGIN(h)
    // global slot positions of this shard's entries behind the carry of its predecessors -- all ranks at once, BEFORE the
    // chain walk goes from rank to rank (shq_mg_denoise_walk then only runs the walk; it computes the positions itself
    // when this was not called)
    flush_stage(h); ensure_state(h);
    if (h->mg_pending.valid) SHQ_FAIL(SHQ_ERR_STATE, "a staged batch is waiting for shq_mg_commit / shq_mg_discard");
    h->walk_pos_ready = false;
    if (h->n_state == 0) return;
    Phase ph(h, &h->stats.ms_denoise);
    MaxPlus init{a.c[0], a.c[1]};
    compute_positions(h, 0, init);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->walk_pos_ready = true; h->walk_pos_n = h->n_state; h->walk_pos_carry[0] = a.c[0]; h->walk_pos_carry[1] = a.c[1];
  SHQ_API_END
}

int shq_mg_denoise_walk(shq_t h, const int64_t carry[2], uint64_t state[16]) {
  if (!h || !carry || !state) return SHQ_ERR_ARG;
  struct { const int64_t* c; uint64_t* s; } args{carry, state};
  SHQ_API_BEGIN(h)
    flush_stage(h); ensure_state(h);
    if (h->mg_pending.valid) SHQ_FAIL(SHQ_ERR_STATE, "a staged batch is waiting for shq_mg_commit / shq_mg_discard");
    cudaStream_t st = h->stream;
    uint64_t& L = h->stats.kernels_launched;
    Phase ph(h, &h->stats.ms_denoise);
    a.s[12] = h->nslots; a.s[13] = (uint64_t)h->shard;
    const size_t n = h->n_state;
    uint32_t keep_cap = (uint32_t)(h->nslots >> 20) + 8;
    h->d_keep.reserve((size_t)(keep_cap + 4) * 4);
    uint32_t* n_keep = h->d_keep.as<uint32_t>(); uint32_t* keep_idx = n_keep + 4;
    SHQ_CUDA(cudaMemsetAsync(n_keep, 0, 16, st));
    if (n == 0) return;       // nothing to walk: the chain passes through unchanged
    const bool ready = h->walk_pos_ready && h->walk_pos_n == n && h->walk_pos_carry[0] == a.c[0] && h->walk_pos_carry[1] == a.c[1];
    h->walk_pos_ready = false;
    if (!ready) { MaxPlus init{a.c[0], a.c[1]}; compute_positions(h, 0, init); }
    h->d_walk.reserve(256);
    SHQ_CUDA(cudaMemcpyAsync(h->d_walk.p, a.s, 128, cudaMemcpyHostToDevice, st));
    denoise_walk_kernel<<<1, 32, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), h->d_pos.as<uint64_t>(), n,
                                          h->d_walk.as<WalkState>(), keep_idx, n_keep, keep_cap); L++;
    SHQ_CUDA(cudaGetLastError());
    uint64_t* p = (uint64_t*)pinned(h, 128);
    SHQ_CUDA(cudaMemcpyAsync(p, h->d_walk.p, 128, cudaMemcpyDeviceToHost, st));
    SHQ_CUDA(cudaStreamSynchronize(st));
    h->stats.host_syncs++;
    memcpy(a.s, p, 128);
  SHQ_API_END
}

int shq_mg_denoise_finalize(uint64_t state[16]) {
  if (!state) return SHQ_ERR_ARG;
  // the range that is open when the entries run out: the only candidate is the very last entry (state[5..10])
  const uint64_t phase = state[0], e0 = state[2], cur_end = state[4];
  const bool valid = state[5] != 0;
  const uint64_t pos = state[6], end = state[7], count = state[8], gap = state[9], owner = state[10];
  if (!valid || phase == 0) return SHQ_OK;
  uint64_t range_end;
  if (phase == 1) range_end = end >= e0 ? end : e0 - 1;     // slot e0 covered by the last entry: its cluster ends the range
  else range_end = cur_end;
  if (pos == range_end && end == range_end && count == 1 && gap) state[11] |= 1ULL << owner;
  state[0] = 0;
  return SHQ_OK;
}

int shq_mg_denoise_apply(shq_t h, int protect_last, uint64_t* removed) {
  if (!h) return SHQ_ERR_ARG;
  struct { int pl; uint64_t* r; } args{protect_last, removed};
  SHQ_API_BEGIN(h)
    cudaStream_t st = h->stream;
    uint64_t& L = h->stats.kernels_launched;
    Phase ph(h, &h->stats.ms_denoise);
    const size_t n = h->n_state;
    h->rounds_done++;
    h->stats.denoise_launches++; h->stats.denoise_entries += n;
    if (a.r) *a.r = 0;
    if (n == 0) return;
    uint32_t keep_cap = (uint32_t)(h->nslots >> 20) + 8;
    uint32_t* n_keep = h->d_keep.as<uint32_t>(); uint32_t* keep_idx = n_keep + 4;
    denoise_keep_kernel<<<cdiv(keep_cap, 128), 128, 0, st>>>(keep_idx, n_keep, keep_cap, h->st_counts.as<uint32_t>()); L++;
    if (a.pl) {   // our last entry is the one-slot cluster a later shard's range ended on
      uint32_t c = 0;
      SHQ_CUDA(cudaMemcpyAsync(&c, h->st_counts.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
      SHQ_CUDA(cudaStreamSynchronize(st));
      if (c == 1) { uint32_t z = 0; SHQ_CUDA(cudaMemcpyAsync(h->st_counts.as<uint32_t>() + (n - 1), &z, 4, cudaMemcpyHostToDevice, st)); SHQ_CUDA(cudaStreamSynchronize(st)); }
    }
    const size_t ntiles = (n + kScanTile - 1) / kScanTile;
    h->d_flags.reserve((ntiles + 4) * 8);
    uint64_t* tile_cnt = h->d_flags.as<uint64_t>();
    denoise_tiles_kernel<false><<<(unsigned)ntiles, kScanThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), n, tile_cnt, nullptr, nullptr); L++;
    exclusive_scan_inplace<SumU64>(tile_cnt, ntiles, h->d_scan_scratch, st, tile_cnt + ntiles, &L);
    uint64_t kept = read_scalar(h, tile_cnt + ntiles);
    h->st_keys2.reserve((kept + 4) * 8); h->st_counts2.reserve((kept + 4) * 4);
    denoise_tiles_kernel<true><<<(unsigned)ntiles, kScanThreads, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), n, tile_cnt,
                                                                       h->st_keys2.as<uint64_t>(), h->st_counts2.as<uint32_t>()); L++;
    SHQ_CUDA(cudaGetLastError());
    std::swap(h->st_keys, h->st_keys2); std::swap(h->st_counts, h->st_counts2);
    const uint64_t rem = n - kept;
    h->n_state = kept;
    h->nelts -= rem; h->ndistinct -= rem;
    h->image_valid = false;
    if (a.r) *a.r = rem;
  SHQ_API_END
}


// Sharded image across processes: every rank exports the CUDA IPC handle of its block array (after
// shq_mg_materialize_shard); a rank that opens all of them can query the WHOLE filter -- and run the unitig build over
// it -- with loads (and traveled-bit atomics) that go straight to the owners' HBM.  No rank holds the whole image.
int shq_mg_image_ipc(shq_t h, void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return SHQ_ERR_ARG;
  struct { void* out; } args{ipc_handle_out};
  SHQ_API_BEGIN(h)
    if (!h->image.p) SHQ_FAIL(SHQ_ERR_STATE, "shq_mg_materialize_shard first");
    cudaIpcMemHandle_t hd;
    SHQ_CUDA(cudaIpcGetMemHandle(&hd, h->image.p));
    memcpy(a.out, &hd, 64);
  SHQ_API_END
}
int shq_mg_open_image_peers(shq_t h, const void* all_handles, int world) {
  if (!h || !all_handles || world != h->nshards) return SHQ_ERR_ARG;
  struct { const void* hs; int world; } args{all_handles, world};
  SHQ_API_BEGIN(h)
    for (void* p : h->img_opened) cudaIpcCloseMemHandle(p);
    h->img_opened.clear();
    h->img_peer.assign((size_t)a.world, nullptr);
    for (int r = 0; r < a.world; r++) {
      if (r == h->shard) { h->img_peer[r] = h->image.as<uint8_t>(); continue; }
      cudaIpcMemHandle_t hd;
      memcpy(&hd, (const char*)a.hs + (size_t)r * 64, 64);
      void* p = nullptr;
      SHQ_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      h->img_opened.push_back(p);
      h->img_peer[r] = (uint8_t*)p;
    }
    h->d_img_peer.reserve((size_t)a.world * 8 + 64);
    SHQ_CUDA(cudaMemcpyAsync(h->d_img_peer.p, h->img_peer.data(), (size_t)a.world * 8, cudaMemcpyHostToDevice, h->stream));
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}
int shq_mg_close_image_peers(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  cudaSetDevice(h->device);
  for (void* p : h->img_opened) cudaIpcCloseMemHandle(p);
  h->img_opened.clear(); h->img_peer.clear();
  return SHQ_OK;
}

// ---- sharded image: layout with a carry from the previous shards, spill hand-over, per-shard blocks ----
int shq_mg_layout_aggregate(shq_t h, int64_t agg[2]) {
  if (!h || !agg) return SHQ_ERR_ARG;
  struct { int64_t* g; } args{agg};
  SHQ_API_BEGIN(h)
    flush_stage(h); ensure_state(h);
    MaxPlus m;
    compute_positions(h, 0, MaxPlusOp::id(), &m, false);      // global slot coordinates
    a.g[0] = m.a; a.g[1] = m.b;
  SHQ_API_END
}

int shq_mg_layout_tail(shq_t h, const int64_t carry[2], size_t cap, uint64_t* d_keys, uint32_t* d_counts, uint64_t* d_pos, size_t* n_tail) {
  if (!h || !carry || !n_tail) return SHQ_ERR_ARG;
  struct { const int64_t* c; size_t cap; uint64_t* k; uint32_t* cn; uint64_t* p; size_t* nt; } args{carry, cap, d_keys, d_counts, d_pos, n_tail};
  SHQ_API_BEGIN(h)
    flush_stage(h); ensure_state(h);
    cudaStream_t st = h->stream;
    *a.nt = 0;
    MaxPlus init{a.c[0], a.c[1]};
    uint64_t used = compute_positions(h, 0, init);
    const uint64_t boundary = (h->key_hi >> 8);                // first slot of the next shard
    const bool last = h->shard == h->nshards - 1;
    if (last) {
      if (used > h->xnslots) SHQ_FAIL(SHQ_ERR_CAPACITY, "filter over-full: layout needs slot %llu but xnslots = %llu",
                                      (unsigned long long)used, (unsigned long long)h->xnslots);
      return;
    }
    if (h->n_state == 0 || used <= boundary) return;
    tail_start_kernel<<<1, 32, 0, st>>>(h->st_keys.as<uint64_t>(), h->st_counts.as<uint32_t>(), h->d_pos.as<uint64_t>(), h->n_state, boundary,
                                        h->d_misc.as<uint64_t>() + 8);
    h->stats.kernels_launched++;
    uint64_t k0 = read_scalar(h, h->d_misc.as<uint64_t>() + 8);
    size_t nt = h->n_state - (size_t)k0;
    if (nt > a.cap) SHQ_FAIL(SHQ_ERR_CAPACITY, "%zu entries spill into the next shard (buffer holds %zu)", nt, a.cap);
    if (nt) {
      SHQ_CUDA(cudaMemcpyAsync(a.k, h->st_keys.as<uint64_t>() + k0, nt * 8, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(a.cn, h->st_counts.as<uint32_t>() + k0, nt * 4, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(a.p, h->d_pos.as<uint64_t>() + k0, nt * 8, cudaMemcpyDeviceToDevice, st));
    }
    SHQ_CUDA(cudaStreamSynchronize(st));
    *a.nt = nt;
  SHQ_API_END
}

int shq_mg_materialize_shard(shq_t h, const uint64_t* d_spill_keys, const uint32_t* d_spill_counts, const uint64_t* d_spill_pos,
                             size_t n_spill, const uint8_t** d_image, uint64_t* first_block, uint64_t* n_blocks) {
  if (!h || (n_spill && (!d_spill_keys || !d_spill_counts || !d_spill_pos))) return SHQ_ERR_ARG;
  struct { const uint64_t* k; const uint32_t* c; const uint64_t* p; size_t n; const uint8_t** img; uint64_t* fb; uint64_t* nb; }
      args{d_spill_keys, d_spill_counts, d_spill_pos, n_spill, d_image, first_block, n_blocks};
  SHQ_API_BEGIN(h)
    // positions (global coordinates) must be current: shq_mg_layout_tail has just run on this state
    cudaStream_t st = h->stream;
    Phase ph(h, &h->stats.ms_layout);
    const uint64_t per = h->nslots / (uint64_t)h->nshards / 64;      // blocks per shard
    const uint64_t b_first = per * (uint64_t)h->shard;
    const uint64_t nb = h->shard == h->nshards - 1 ? h->nblocks - b_first : per;
    if (per % kTileBlocks) SHQ_FAIL(SHQ_ERR_ARG, "shard of %llu blocks is not a multiple of %d", (unsigned long long)per, kTileBlocks);
    const size_t n = h->n_state, m = a.n + n;
    // entries that can touch this shard's blocks: the previous shard's spill, then our own state
    h->d_tmp_keys.reserve((m + 4) * 8); h->d_tmp_counts.reserve((m + 4) * 4); h->d_run_keys.reserve((m + 4) * 8);
    uint64_t* ck = h->d_tmp_keys.as<uint64_t>(); uint32_t* cc = h->d_tmp_counts.as<uint32_t>(); uint64_t* cp = h->d_run_keys.as<uint64_t>();
    if (a.n) {
      SHQ_CUDA(cudaMemcpyAsync(ck, a.k, a.n * 8, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(cc, a.c, a.n * 4, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(cp, a.p, a.n * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (n) {
      SHQ_CUDA(cudaMemcpyAsync(ck + a.n, h->st_keys.p, n * 8, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(cc + a.n, h->st_counts.p, n * 4, cudaMemcpyDeviceToDevice, st));
      SHQ_CUDA(cudaMemcpyAsync(cp + a.n, h->d_pos.p, n * 8, cudaMemcpyDeviceToDevice, st));
    }
    size_t tiles = (nb + kTileBlocks - 1) / kTileBlocks;
    h->image.reserve(tiles * kTileBytes + 16);
    image_tiles_kernel<<<(unsigned)tiles, kImgThreads, 0, st>>>(ck, cc, cp, m, 0, nb, h->image.as<uint8_t>(), b_first);
    h->stats.kernels_launched++;
    SHQ_CUDA(cudaGetLastError());
    SHQ_CUDA(cudaStreamSynchronize(st));
    if (a.img) *a.img = h->image.as<uint8_t>();
    if (a.fb) *a.fb = b_first;
    if (a.nb) *a.nb = nb;
  SHQ_API_END
}

int shq_header(shq_t h, uint64_t ndistinct, uint64_t nelts, uint8_t out[128]) {
  if (!h || !out) return SHQ_ERR_ARG;
  uint64_t sd = h->ndistinct, se = h->nelts;
  h->ndistinct = ndistinct; h->nelts = nelts;
  build_header(h, out);
  h->ndistinct = sd; h->nelts = se;
  return SHQ_OK;
}

int shq_state_view(shq_t h, const uint64_t** d_keys, const uint32_t** d_counts, size_t* n) {
  if (!h || !n) return SHQ_ERR_ARG;
  struct { const uint64_t** k; const uint32_t** c; size_t* n; } args{d_keys, d_counts, n};
  SHQ_API_BEGIN(h)
    flush_stage(h);
    ensure_state(h);
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    if (a.k) *a.k = h->st_keys.as<uint64_t>();
    if (a.c) *a.c = h->st_counts.as<uint32_t>();
    *a.n = h->n_state;
  SHQ_API_END
}

int shq_adopt_state(shq_t h, const uint64_t* d_keys, const uint32_t* d_counts, size_t n, uint64_t ndistinct, uint64_t nelts) {
  if (!h || (n && (!d_keys || !d_counts))) return SHQ_ERR_ARG;
  struct { const uint64_t* k; const uint32_t* c; size_t n; uint64_t nd, ne; } args{d_keys, d_counts, n, ndistinct, nelts};
  SHQ_API_BEGIN(h)
    h->h_stage_fill = 0; h->stage_chunk_ends.clear();
    h->st_keys.reserve((a.n + 4) * 8); h->st_counts.reserve((a.n + 4) * 4);
    if (a.n) {
      SHQ_CUDA(cudaMemcpyAsync(h->st_keys.p, a.k, a.n * 8, cudaMemcpyDeviceToDevice, h->stream));
      SHQ_CUDA(cudaMemcpyAsync(h->st_counts.p, a.c, a.n * 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
    h->n_state = a.n; h->ndistinct = a.nd; h->nelts = a.ne;
    h->image_valid = false; h->image_only = false;
  SHQ_API_END
}

int shq_copy_device(shq_t h, void* dst, const void* src, size_t bytes) {
  if (!h || (bytes && (!dst || !src))) return SHQ_ERR_ARG;
  struct { void* d; const void* s; size_t n; } args{dst, src, bytes};
  SHQ_API_BEGIN(h)
    if (a.n) SHQ_CUDA(cudaMemcpyAsync(a.d, a.s, a.n, cudaMemcpyDeviceToDevice, h->stream));
    SHQ_CUDA(cudaStreamSynchronize(h->stream));
  SHQ_API_END
}

int shq_set_counts(shq_t h, uint64_t ndistinct, uint64_t nelts) {
  if (!h) return SHQ_ERR_ARG;
  h->ndistinct = ndistinct; h->nelts = nelts;
  return SHQ_OK;
}

int shq_get_stats(shq_t h, shq_stats* out) {
  if (!h || !out) return SHQ_ERR_ARG;
  cudaSetDevice(h->device);
  resolve_phases(h);
  *out = h->stats;
  return SHQ_OK;
}
int shq_reset_stats(shq_t h) {
  if (!h) return SHQ_ERR_ARG;
  cudaSetDevice(h->device);
  resolve_phases(h);
  memset(&h->stats, 0, sizeof h->stats);
  return SHQ_OK;
}
void* shq_stream(shq_t h) { return h ? (void*)h->stream : nullptr; }

}  // extern "C"

// ---- unitig entry points: implemented in shq_unitig.cuh (included below) ----
#include "shq_unitig.cuh"
