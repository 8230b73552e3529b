"""sh-assembly_b200: thin Python binding (ctypes) over libshq.so, the B200-native C-ABI library
that replaces SH-assembly's counting / deNoise / unitig hot path (include/shq.h).

The directory name carries a hyphen (it mirrors the reference's name), so import it with
    shq = importlib.import_module("sh-assembly_b200")
This module is plumbing for tests and bench.py; the product is the C ABI.  There is no CPU
fallback: if libshq.so is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libshq.so")

SEED = 2038074761  # src/CQF-deNoise.cpp:83

STATUS = {0: "SHQ_OK", 1: "SHQ_ERR_ARG", 2: "SHQ_ERR_CUDA", 3: "SHQ_ERR_CAPACITY", 4: "SHQ_ERR_MALFORMED",
          5: "SHQ_ERR_IO", 6: "SHQ_ERR_FORMAT", 7: "SHQ_ERR_COUNT_OVERFLOW", 8: "SHQ_ERR_STATE"}


class ShqError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        super().__init__("%s: %s" % (STATUS.get(status, status), detail))


class shq_stats(C.Structure):
    _fields_ = [("kernels_launched", C.c_uint64), ("batches", C.c_uint64), ("keys_hashed", C.c_uint64),
                ("ms_hash_count", C.c_double), ("ms_scatter", C.c_double), ("ms_aggregate", C.c_double),
                ("ms_merge", C.c_double), ("ms_layout", C.c_double), ("ms_denoise", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("bin_fallbacks", C.c_uint64),
                ("scatter_launches", C.c_uint64), ("scatter_text_bytes", C.c_uint64), ("scatter_words", C.c_uint64),
                ("agg_launches", C.c_uint64), ("agg_words", C.c_uint64), ("agg_distinct", C.c_uint64),
                ("merge_launches", C.c_uint64), ("merge_in", C.c_uint64), ("merge_out", C.c_uint64),
                ("denoise_launches", C.c_uint64), ("denoise_entries", C.c_uint64),
                ("layout_launches", C.c_uint64), ("layout_entries", C.c_uint64), ("host_syncs", C.c_uint64),
                ("spec_rejected", C.c_uint64), ("ms_split", C.c_double), ("split_launches", C.c_uint64), ("split_words", C.c_uint64),
                ("wide_batches", C.c_uint64), ("wide_fallbacks", C.c_uint64)]


class shq_unitig_opts(C.Structure):
    _fields_ = [("K", C.c_int), ("abundance_min", C.c_int), ("solid_abundance_min", C.c_int),
                ("solid_abundance_max", C.c_int)]


class shq_params(C.Structure):
    _fields_ = [("num_true_kmers", C.c_uint64), ("num_false_kmers", C.c_uint64), ("num_deNoise", C.c_int),
                ("lower_bound", C.c_int), ("upper_bound", C.c_int), ("qb", C.c_uint64), ("hb", C.c_uint64),
                ("trigger", C.c_uint64), ("fr", C.c_double)]


def build(verbose=False):
    """Compile libshq.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libshq.so failed")
    return LIB_PATH


_lib = None

# every symbol include/shq.h declares
EXPORTS = [
    "shq_strerror", "shq_last_error", "shq_create", "shq_create_shard", "shq_destroy", "shq_reset", "shq_set_denoise_schedule", "shq_set_progress_log",
    "shq_insert_fastq_chunk", "shq_insert_fastq_pinned", "shq_insert_fastq_pinned_chunks", "shq_insert_fastq_device", "shq_insert_keys", "shq_insert_keys_device", "shq_finish",
    "shq_sync", "shq_counts", "shq_rounds_done", "shq_denoise", "shq_image_size", "shq_export_cqf", "shq_save",
    "shq_load", "shq_load_image", "shq_geometry", "shq_iterate", "shq_query_set_traveled", "shq_query",
    "shq_hash_fastq_chunk", "shq_unitigs_begin", "shq_unitigs_set_seed_rule", "shq_unitigs_add_reads", "shq_unitigs_build", "shq_unitigs_count",
    "shq_unitigs_write_fasta", "shq_unitigs_fasta", "shq_unitigs_free", "shq_get_stats", "shq_reset_stats", "shq_stream",
    "shq_reader_open", "shq_reader_open_memory", "shq_reader_next", "shq_reader_pin", "shq_reader_close", "shq_sync_copies", "shq_build_kmer_spectrum",
    "shq_materialize", "shq_param_block", "shq_cdfpoi_positive", "shq_cdfpoi",
    "shq_mg_setup", "shq_mg_open_peers", "shq_mg_bin_local", "shq_mg_insert_pull",
    "shq_mg_image_ipc", "shq_mg_open_image_peers", "shq_mg_close_image_peers", "shq_unitigs_begin_sharded",
    "shq_mg_stage_pull", "shq_mg_commit", "shq_mg_discard", "shq_mg_denoise_positions", "shq_mg_denoise_walk", "shq_mg_denoise_finalize", "shq_mg_denoise_apply",
    "shq_mg_bin_fastq_device", "shq_mg_insert_binned", "shq_mg_layout_aggregate", "shq_mg_layout_tail", "shq_mg_materialize_shard", "shq_header",
    "shq_hash_fastq_device", "shq_partition_keys_device", "shq_state_view", "shq_adopt_state", "shq_set_counts", "shq_copy_device",
]


def lib():
    """Load libshq.so.  Raises (loudly) if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libshq.so is not built (%s); run __graft_entry__.build() -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
        L.shq_strerror.restype = C.c_char_p
        L.shq_strerror.argtypes = [C.c_int]
        L.shq_last_error.restype = C.c_char_p
        L.shq_last_error.argtypes = [vp]
        L.shq_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(vp)]
        L.shq_create_shard.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.shq_destroy.argtypes = [vp]
        L.shq_destroy.restype = None
        L.shq_reset.argtypes = [vp]
        L.shq_set_denoise_schedule.argtypes = [vp, u64, C.c_uint32, C.c_int]
        L.shq_set_progress_log.argtypes = [vp, C.c_int, u64]
        L.shq_insert_fastq_chunk.argtypes = [vp, vp, sz, C.c_int]
        L.shq_insert_fastq_pinned.argtypes = [vp, vp, sz, C.c_int]
        L.shq_insert_fastq_device.argtypes = [vp, vp, sz, vp, sz, C.c_int]
        L.shq_insert_fastq_pinned_chunks.argtypes = [vp, vp, sz, vp, sz, C.c_int]
        L.shq_insert_keys.argtypes = [vp, vp, sz]
        L.shq_insert_keys_device.argtypes = [vp, vp, sz]
        L.shq_finish.argtypes = [vp]
        L.shq_sync.argtypes = [vp]
        L.shq_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.shq_rounds_done.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.shq_denoise.argtypes = [vp, C.POINTER(u64)]
        L.shq_image_size.argtypes = [vp, C.POINTER(sz)]
        L.shq_export_cqf.argtypes = [vp, vp, sz, C.POINTER(sz)]
        L.shq_save.argtypes = [vp, C.c_char_p]
        L.shq_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.shq_load_image.argtypes = [vp, sz, C.c_int, C.POINTER(vp)]
        L.shq_geometry.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.shq_iterate.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
        L.shq_query_set_traveled.argtypes = [vp, vp, sz, vp, vp]
        L.shq_query.argtypes = [vp, vp, sz, vp]
        L.shq_hash_fastq_chunk.argtypes = [vp, vp, sz, C.c_int, vp, sz, C.POINTER(sz)]
        L.shq_unitigs_begin.argtypes = [vp, C.POINTER(shq_unitig_opts), C.POINTER(vp)]
        L.shq_unitigs_set_seed_rule.argtypes = [vp, C.c_int]
        L.shq_unitigs_add_reads.argtypes = [vp, vp, sz]
        L.shq_unitigs_build.argtypes = [vp]
        L.shq_unitigs_count.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
        L.shq_unitigs_write_fasta.argtypes = [vp, C.c_char_p]
        L.shq_unitigs_fasta.argtypes = [vp, vp, sz, C.POINTER(sz)]
        L.shq_unitigs_free.argtypes = [vp]
        L.shq_unitigs_free.restype = None
        L.shq_get_stats.argtypes = [vp, C.POINTER(shq_stats)]
        L.shq_reset_stats.argtypes = [vp]
        L.shq_stream.restype = vp
        L.shq_stream.argtypes = [vp]
        L.shq_hash_fastq_device.argtypes = [vp, vp, sz, C.c_int, vp, sz, C.POINTER(sz)]
        L.shq_partition_keys_device.argtypes = [vp, vp, sz, C.c_int, vp, C.POINTER(u64)]
        L.shq_mg_bin_fastq_device.argtypes = [vp, vp, sz, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]
        L.shq_mg_insert_binned.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, sz]
        L.shq_mg_setup.argtypes = [vp, sz, vp]
        L.shq_mg_open_peers.argtypes = [vp, vp, C.c_int]
        L.shq_mg_bin_local.argtypes = [vp, vp, sz, C.c_int, C.c_int, vp, C.POINTER(sz)]
        L.shq_mg_insert_pull.argtypes = [vp, vp, C.c_int]
        L.shq_mg_image_ipc.argtypes = [vp, vp]
        L.shq_mg_open_image_peers.argtypes = [vp, vp, C.c_int]
        L.shq_mg_close_image_peers.argtypes = [vp]
        L.shq_unitigs_begin_sharded.argtypes = [vp, C.POINTER(shq_unitig_opts), u64, C.POINTER(vp)]
        L.shq_mg_stage_pull.argtypes = [vp, vp, C.c_int, C.POINTER(u64), C.POINTER(u64)]
        L.shq_mg_commit.argtypes = [vp]
        L.shq_mg_discard.argtypes = [vp]
        L.shq_mg_denoise_walk.argtypes = [vp, C.POINTER(C.c_int64), vp]
        L.shq_mg_denoise_positions.argtypes = [vp, C.POINTER(C.c_int64)]
        L.shq_mg_denoise_finalize.argtypes = [vp]
        L.shq_mg_denoise_apply.argtypes = [vp, C.c_int, C.POINTER(u64)]
        L.shq_mg_layout_aggregate.argtypes = [vp, C.POINTER(C.c_int64)]
        L.shq_mg_layout_tail.argtypes = [vp, C.POINTER(C.c_int64), sz, vp, vp, vp, C.POINTER(sz)]
        L.shq_mg_materialize_shard.argtypes = [vp, vp, vp, vp, sz, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
        L.shq_header.argtypes = [vp, u64, u64, vp]
        L.shq_state_view.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]
        L.shq_adopt_state.argtypes = [vp, vp, vp, sz, u64, u64]
        L.shq_set_counts.argtypes = [vp, u64, u64]
        L.shq_copy_device.argtypes = [vp, vp, vp, sz]
        L.shq_reader_open.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char, C.POINTER(vp)]
        L.shq_reader_open_memory.argtypes = [C.POINTER(vp), C.POINTER(sz), C.c_int, u64, C.POINTER(vp)]
        L.shq_reader_next.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(C.c_int)]
        L.shq_reader_pin.argtypes = [vp]
        L.shq_sync_copies.argtypes = [vp]
        L.shq_reader_close.argtypes = [vp]
        L.shq_reader_close.restype = None
        L.shq_build_kmer_spectrum.argtypes = [vp, vp, C.c_int]
        L.shq_materialize.argtypes = [vp]
        L.shq_param_block.argtypes = [C.c_int, u64, u64, C.c_double, C.c_char_p, C.c_int, C.c_double, C.POINTER(shq_params)]
        L.shq_cdfpoi_positive.restype = C.c_double
        L.shq_cdfpoi_positive.argtypes = [C.c_double, C.c_double]
        L.shq_cdfpoi.restype = C.c_double
        L.shq_cdfpoi.argtypes = [C.c_double, C.c_double]
        _lib = L
    return _lib


def param_block(K, n_true, N_total, alpha=-1.0, error_profile="", deNoise=-1, fr=0.0):
    """src/CQF-deNoise.cpp:83-161 through the C ABI."""
    out = shq_params()
    rc = lib().shq_param_block(K, n_true, N_total, alpha, error_profile.encode(), deNoise, fr, C.byref(out))
    if rc:
        raise ShqError(rc, "shq_param_block")
    return out


def build_unitigs(cqf, chunks, K, s_min=2, x_min=2, x_max=1000000, master_rule=False, global_ndistinct=None):
    """Contiger device path: returns the unitig FASTA text (BCALM/Minia style headers).
    global_ndistinct: cqf is one shard of a sharded filter whose peers' images are open (multi-GPU)."""
    L = lib()
    opts = shq_unitig_opts(K, s_min, x_min, x_max)
    u = C.c_void_p()
    if global_ndistinct is None:
        cqf._check(L.shq_unitigs_begin(cqf.h, C.byref(opts), C.byref(u)))
    else:
        cqf._check(L.shq_unitigs_begin_sharded(cqf.h, C.byref(opts), int(global_ndistinct), C.byref(u)))
    try:
        if master_rule:
            cqf._check(L.shq_unitigs_set_seed_rule(u, 1))
        for ch in chunks:
            p, n, keep = _buf(ch)
            cqf._check(L.shq_unitigs_add_reads(u, p, n))
        cqf._check(L.shq_unitigs_build(u))
        w = C.c_size_t()
        cqf._check(L.shq_unitigs_fasta(u, None, 0, C.byref(w)))
        buf = C.create_string_buffer(max(1, w.value))
        cqf._check(L.shq_unitigs_fasta(u, buf, w.value, C.byref(w)))
        return buf.raw[:w.value].decode()
    finally:
        L.shq_unitigs_free(u)


class Reader:
    """seqFile_batch (cqf/CQF_mt.h:334-412): chunks in the reference's t=1 order."""

    def __init__(self, paths=None, fmode="f", buffers=None, part_size=0):
        self.L = lib()
        h = C.c_void_p()
        if buffers is not None:
            self._keep = [np.frombuffer(b, np.uint8) for b in buffers]
            ptrs = (C.c_void_p * len(buffers))(*[a.ctypes.data for a in self._keep])
            lens = (C.c_size_t * len(buffers))(*[a.nbytes for a in self._keep])
            rc = self.L.shq_reader_open_memory(ptrs, lens, len(buffers), part_size, C.byref(h))
        else:
            arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
            rc = self.L.shq_reader_open(arr, len(paths), fmode.encode(), C.byref(h))
        if rc:
            raise ShqError(rc, "shq_reader_open")
        self.h = h

    def __iter__(self):
        return self

    def next_raw(self):
        """(address, size, file_index) of the next chunk, or None; valid until the next call."""
        p, n, fi = C.c_void_p(), C.c_size_t(), C.c_int()
        rc = self.L.shq_reader_next(self.h, C.byref(p), C.byref(n), C.byref(fi))
        if rc:
            raise ShqError(rc, "shq_reader_next")
        if not p.value:
            return None
        return p.value, n.value, fi.value

    def __next__(self):
        r = self.next_raw()
        if r is None:
            raise StopIteration
        return r[2], C.string_at(r[0], r[1])

    def close(self):
        if getattr(self, "h", None):
            self.L.shq_reader_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _buf(b):
    """(address, length, keepalive) of a bytes / bytearray / numpy buffer without copying."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b)
        return a.ctypes.data, a.nbytes, a
    a = np.frombuffer(b, np.uint8)
    return a.ctypes.data, a.nbytes, a


class CQF:
    """Host-side mirror of the reference's class CQF_mt (cqf/CQF_mt.h:414-558) over the C ABI."""

    def __init__(self, qb=None, hb=None, seed=SEED, device=0, shard=0, nshards=1, handle=None):
        self.L = lib()
        if handle is not None:
            self.h = handle
        else:
            h = C.c_void_p()
            self._check(self.L.shq_create_shard(qb, hb if hb is not None else qb + 8, seed, device, shard, nshards, C.byref(h)), None)
            self.h = h
        qb_, hb_, ns, xn, nb = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.shq_geometry(self.h, C.byref(qb_), C.byref(hb_), C.byref(ns), C.byref(xn), C.byref(nb))
        self.qb, self.hb, self.nslots, self.xnslots, self.nblocks = qb_.value, hb_.value, ns.value, xn.value, nb.value

    def _check(self, rc, h="self"):
        if rc:
            detail = ""
            hh = self.h if h == "self" else h
            if hh:
                detail = (self.L.shq_last_error(hh) or b"").decode()
            raise ShqError(rc, detail)

    def close(self):
        if getattr(self, "h", None):
            self.L.shq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(self.L.shq_reset(self.h))

    # --- counting
    def set_denoise_schedule(self, trigger, rounds, end_denoise):
        self._check(self.L.shq_set_denoise_schedule(self.h, trigger, rounds, 1 if end_denoise else 0))

    def insert_fastq_chunk(self, text, K):
        p, n, keep = _buf(text)
        self._check(self.L.shq_insert_fastq_chunk(self.h, p, n, K))

    def insert_fastq_device(self, d_ptr, n, chunk_ends, K):
        ce = np.ascontiguousarray(chunk_ends, np.uint64)
        self._check(self.L.shq_insert_fastq_device(self.h, d_ptr, n, ce.ctypes.data, len(ce), K))

    def insert_keys(self, keys):
        k = np.ascontiguousarray(keys, np.uint64)
        self._check(self.L.shq_insert_keys(self.h, k.ctypes.data, len(k)))

    def insert_keys_device(self, d_ptr, n):
        self._check(self.L.shq_insert_keys_device(self.h, d_ptr, n))

    def finish(self):
        self._check(self.L.shq_finish(self.h))

    def materialize(self):
        self._check(self.L.shq_materialize(self.h))

    def build_kmer_spectrum(self, reader, K):
        self._check(self.L.shq_build_kmer_spectrum(self.h, reader.h, K))

    def insert_fastq_pinned_raw(self, addr, n, K):
        self._check(self.L.shq_insert_fastq_pinned(self.h, addr, n, K))

    def insert_fastq_pinned_chunks(self, addr, n, chunk_ends, K):
        ce = np.ascontiguousarray(chunk_ends, np.uint64)
        self._check(self.L.shq_insert_fastq_pinned_chunks(self.h, addr, n, ce.ctypes.data, len(ce), K))

    def insert_fastq_chunk_raw(self, addr, n, K):
        self._check(self.L.shq_insert_fastq_chunk(self.h, addr, n, K))

    def sync(self):
        self._check(self.L.shq_sync(self.h))

    def counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.L.shq_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def rounds_done(self):
        r = C.c_uint32()
        self._check(self.L.shq_rounds_done(self.h, C.byref(r)))
        return r.value

    def denoise(self):
        r = C.c_uint64()
        self._check(self.L.shq_denoise(self.h, C.byref(r)))
        return r.value

    # --- image
    def export(self) -> bytes:
        n = C.c_size_t()
        self._check(self.L.shq_image_size(self.h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        w = C.c_size_t()
        self._check(self.L.shq_export_cqf(self.h, buf.ctypes.data, n.value, C.byref(w)))
        return buf[:w.value].tobytes()

    def save(self, path):
        self._check(self.L.shq_save(self.h, path.encode()))

    @classmethod
    def load(cls, path, device=0):
        L = lib()
        h = C.c_void_p()
        rc = L.shq_load(path.encode(), device, C.byref(h))
        if rc:
            raise ShqError(rc, "shq_load(%s)" % path)
        return cls(handle=h)

    @classmethod
    def from_bytes(cls, image, device=0):
        L = lib()
        h = C.c_void_p()
        p, n, keep = _buf(image)
        rc = L.shq_load_image(p, n, device, C.byref(h))
        if rc:
            raise ShqError(rc, "shq_load_image")
        return cls(handle=h)

    def iterate(self):
        n = C.c_size_t()
        self._check(self.L.shq_iterate(self.h, None, None, 0, C.byref(n)))
        k = np.zeros(n.value, np.uint64)
        c = np.zeros(n.value, np.uint64)
        if n.value:
            self._check(self.L.shq_iterate(self.h, k.ctypes.data, c.ctypes.data, n.value, C.byref(n)))
        return k, c

    # --- queries
    def query_set_traveled(self, keys):
        k = np.ascontiguousarray(keys, np.uint64)
        cnt = np.zeros(len(k), np.uint64)
        was = np.zeros(len(k), np.uint8)
        self._check(self.L.shq_query_set_traveled(self.h, k.ctypes.data, len(k), cnt.ctypes.data, was.ctypes.data))
        return cnt, was

    def query(self, keys):
        k = np.ascontiguousarray(keys, np.uint64)
        cnt = np.zeros(len(k), np.uint64)
        self._check(self.L.shq_query(self.h, k.ctypes.data, len(k), cnt.ctypes.data))
        return cnt

    def hash_fastq_chunk(self, text, K):
        p, n, keep = _buf(text)
        nk = C.c_size_t()
        cap = max(16, n)
        out = np.zeros(cap, np.uint64)
        self._check(self.L.shq_hash_fastq_chunk(self.h, p, n, K, out.ctypes.data, cap, C.byref(nk)))
        return out[:nk.value]

    # --- sharded building blocks
    def hash_fastq_device(self, d_text, n, K, d_keys, cap):
        nk = C.c_size_t()
        self._check(self.L.shq_hash_fastq_device(self.h, d_text, n, K, d_keys, cap, C.byref(nk)))
        return nk.value

    def partition_keys_device(self, d_in, n, nparts, d_out):
        counts = (C.c_uint64 * nparts)()
        self._check(self.L.shq_partition_keys_device(self.h, d_in, n, nparts, d_out, counts))
        return list(counts)

    def mg_bin_fastq_device(self, d_text, n, K, coarse_bits):
        """(device pointer to the binned words, device pointer to the 2^coarse_bits + 1 offsets, number of words)"""
        w, o, nw = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._check(self.L.shq_mg_bin_fastq_device(self.h, d_text, n, K, coarse_bits, C.byref(w), C.byref(o), C.byref(nw)))
        return w.value, o.value, nw.value

    def mg_insert_binned(self, d_recv, seg_start, d_counts, world, coarse_bits, fine_bits, n_words):
        ss = np.ascontiguousarray(seg_start, np.uint64)
        self._check(self.L.shq_mg_insert_binned(self.h, d_recv, ss.ctypes.data, d_counts, world, coarse_bits, fine_bits, n_words))

    def mg_setup(self, max_words) -> bytes:
        """Allocate the two exchange buffers; returns their CUDA IPC handles (2 x 64 bytes)."""
        buf = np.zeros(128, np.uint8)
        self._check(self.L.shq_mg_setup(self.h, max_words, buf.ctypes.data))
        return buf.tobytes()

    def mg_open_peers(self, all_handles: bytes, world):
        a = np.frombuffer(all_handles, np.uint8)
        self._check(self.L.shq_mg_open_peers(self.h, a.ctypes.data, world))

    def mg_bin_local(self, d_text, n, K, parity, d_fine_hist):
        nw = C.c_size_t()
        self._check(self.L.shq_mg_bin_local(self.h, d_text, n, K, parity, d_fine_hist, C.byref(nw)))
        return nw.value

    def mg_insert_pull(self, d_all_hist, parity):
        self._check(self.L.shq_mg_insert_pull(self.h, d_all_hist, parity))

    def mg_image_ipc(self) -> bytes:
        buf = np.zeros(64, np.uint8)
        self._check(self.L.shq_mg_image_ipc(self.h, buf.ctypes.data))
        return buf.tobytes()

    def mg_open_image_peers(self, all_handles: bytes, world):
        a = np.frombuffer(all_handles, np.uint8)
        self._check(self.L.shq_mg_open_image_peers(self.h, a.ctypes.data, world))

    def mg_close_image_peers(self):
        self._check(self.L.shq_mg_close_image_peers(self.h))

    def mg_stage_pull(self, d_all_hist, parity):
        """(new keys, occurrences) of this shard's share of the batch; commit or discard afterwards"""
        nn, no = C.c_uint64(), C.c_uint64()
        self._check(self.L.shq_mg_stage_pull(self.h, d_all_hist, parity, C.byref(nn), C.byref(no)))
        return nn.value, no.value

    def mg_commit(self):
        self._check(self.L.shq_mg_commit(self.h))

    def mg_discard(self):
        self._check(self.L.shq_mg_discard(self.h))

    def mg_denoise_positions(self, carry):
        g = (C.c_int64 * 2)(int(carry[0]), int(carry[1]))
        self._check(self.L.shq_mg_denoise_positions(self.h, g))

    def mg_denoise_walk(self, carry, state):
        """state: numpy uint64[16], updated in place"""
        g = (C.c_int64 * 2)(int(carry[0]), int(carry[1]))
        self._check(self.L.shq_mg_denoise_walk(self.h, g, state.ctypes.data))

    def mg_denoise_apply(self, protect_last):
        r = C.c_uint64()
        self._check(self.L.shq_mg_denoise_apply(self.h, 1 if protect_last else 0, C.byref(r)))
        return r.value

    def mg_layout_aggregate(self):
        g = (C.c_int64 * 2)()
        self._check(self.L.shq_mg_layout_aggregate(self.h, g))
        return int(g[0]), int(g[1])

    def mg_layout_tail(self, carry, cap, d_keys, d_counts, d_pos):
        g = (C.c_int64 * 2)(int(carry[0]), int(carry[1]))
        nt = C.c_size_t()
        self._check(self.L.shq_mg_layout_tail(self.h, g, cap, d_keys, d_counts, d_pos, C.byref(nt)))
        return nt.value

    def mg_materialize_shard(self, d_keys, d_counts, d_pos, n_spill):
        img, fb, nb = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self.L.shq_mg_materialize_shard(self.h, d_keys, d_counts, d_pos, n_spill, C.byref(img), C.byref(fb), C.byref(nb)))
        return img.value, fb.value, nb.value

    def header(self, ndistinct, nelts) -> bytes:
        buf = np.zeros(128, np.uint8)
        self._check(self.L.shq_header(self.h, ndistinct, nelts, buf.ctypes.data))
        return buf.tobytes()

    def state_view(self):
        k, c, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._check(self.L.shq_state_view(self.h, C.byref(k), C.byref(c), C.byref(n)))
        return k.value, c.value, n.value

    def adopt_state(self, d_keys, d_counts, n, ndistinct, nelts):
        self._check(self.L.shq_adopt_state(self.h, d_keys, d_counts, n, ndistinct, nelts))

    def copy_device(self, dst, src, nbytes):
        self._check(self.L.shq_copy_device(self.h, dst, src, nbytes))

    def set_counts(self, ndistinct, nelts):
        self._check(self.L.shq_set_counts(self.h, ndistinct, nelts))

    def stats(self):
        s = shq_stats()
        self._check(self.L.shq_get_stats(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def reset_stats(self):
        self._check(self.L.shq_reset_stats(self.h))

    def stream(self):
        return self.L.shq_stream(self.h)
