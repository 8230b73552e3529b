"""ctypes access to the TEST-ONLY checkers (oracle/liboracle.so, oracle/_ref/libshq_ref.so).

Test infrastructure: import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never from the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = C.c_void_p


def build(ref=True):
    """Compile liboracle.so (always) and, when /root/reference exists, the checkers made of UNMODIFIED reference sources:
    _ref/libshq_ref.so (gqf.c + nthash.hpp), _ref/Contiger_ref and _ref/CQF-deNoise_ref (the two command lines, compiled
    against the header stand-ins of oracle/shim)."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref", "contiger_ref", "cqf_denoise_ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


class orc_params(C.Structure):
    _fields_ = [("num_true_kmers", C.c_uint64), ("num_false_kmers", C.c_uint64),
                ("num_deNoise", C.c_int), ("lower_bound", C.c_int), ("upper_bound", C.c_int),
                ("qb", C.c_uint64), ("hb", C.c_uint64), ("trigger", C.c_uint64), ("fr", C.c_double)]


_orc = None
_ref = None


def oracle():
    global _orc
    if _orc is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.orc_reads_to_keys.restype = C.c_uint64
        L.orc_reads_to_keys.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, _u64p, C.c_uint64]
        L.orc_chunker_new.restype = C.c_void_p
        L.orc_chunker_new.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint64]
        L.orc_chunker_next.restype = C.c_int64
        L.orc_chunker_next.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        L.orc_chunker_free.argtypes = [C.c_void_p]
        L.orc_cqf_new.restype = C.c_void_p
        L.orc_cqf_new.argtypes = [C.c_int, C.c_int, C.c_uint32]
        L.orc_cqf_free.argtypes = [C.c_void_p]
        L.orc_cqf_insert_keys.restype = C.c_uint64
        L.orc_cqf_insert_keys.argtypes = [C.c_void_p, _u64p, C.c_uint64]
        L.orc_cqf_insert_counted.restype = C.c_uint64
        L.orc_cqf_insert_counted.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64]
        L.orc_cqf_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_cqf_set_counts.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        for f in ("orc_cqf_denoise", "orc_cqf_denoise_ideal", "orc_cqf_image_size"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_cqf_iterate.restype = C.c_uint64
        L.orc_cqf_iterate.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64]
        L.orc_cqf_export.restype = C.c_int
        L.orc_cqf_export.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_cqf_import.restype = C.c_void_p
        L.orc_cqf_import.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_cqf_query_set_traveled.restype = C.c_int
        L.orc_cqf_query_set_traveled.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_cqf_query_batch.argtypes = [C.c_void_p, _u64p, C.c_uint64, _u64p, C.c_void_p]
        L.orc_count_run.restype = C.c_void_p
        L.orc_count_run.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint64,
                                    C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int,
                                    C.POINTER(C.c_int)]
        L.orc_param_block.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_int,
                                      C.c_double, C.POINTER(orc_params)]
        L.orc_cdfpoi.restype = C.c_double
        L.orc_cdfpoi.argtypes = [C.c_double, C.c_double]
        L.orc_cdfpoi_positive.restype = C.c_double
        L.orc_cdfpoi_positive.argtypes = [C.c_double, C.c_double]
        L.orc_mean_CDF2deNoise.restype = C.c_int
        L.orc_mean_CDF2deNoise.argtypes = [C.c_double, C.c_double]
        L.orc_true2falseKmer_DP.restype = C.c_double
        L.orc_true2falseKmer_DP.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_ntpc64_base.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_ntpc64_roll.argtypes = [C.c_ubyte, C.c_ubyte, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_contiger_run.restype = C.c_void_p
        L.orc_contiger_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint64,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_contiger_fasta_size.restype = C.c_uint64
        L.orc_contiger_fasta_size.argtypes = [C.c_void_p]
        L.orc_contiger_fasta.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_contiger_queries.restype = C.c_uint64
        L.orc_contiger_queries.argtypes = [C.c_void_p]
        L.orc_contiger_free.argtypes = [C.c_void_p]
        _orc = L
    return _orc


def ref_path():
    return os.path.join(HERE, "_ref", "libshq_ref.so")


def have_ref():
    return os.path.exists(ref_path())


def ref():
    """The unmodified reference core behind ref_shim.cpp (None if it was never built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            return None
        L = C.CDLL(ref_path())
        L.ref_ntpc64_base.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_ntpc64_roll.argtypes = [C.c_ubyte, C.c_ubyte, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_qf_new.restype = C.c_void_p
        L.ref_qf_new.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.ref_qf_free.argtypes = [C.c_void_p]
        L.ref_qf_load.restype = C.c_void_p
        L.ref_qf_load.argtypes = [C.c_char_p]
        L.ref_qf_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_qf_meta.restype = C.c_uint64
        L.ref_qf_meta.argtypes = [C.c_void_p, C.c_int]
        L.ref_qf_set_counts.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.ref_qf_metadata_ptr.restype = C.c_void_p
        L.ref_qf_metadata_ptr.argtypes = [C.c_void_p]
        L.ref_qf_blocks_ptr.restype = C.c_void_p
        L.ref_qf_blocks_ptr.argtypes = [C.c_void_p]
        L.ref_qf_insert_keys.restype = C.c_uint64
        L.ref_qf_insert_keys.argtypes = [C.c_void_p, _u64p, C.c_uint64]
        L.ref_qf_insert_counted.restype = C.c_uint64
        L.ref_qf_insert_counted.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64]
        L.ref_qf_count.restype = C.c_uint64
        L.ref_qf_count.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_qf_count_set_traveled.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ref_qf_query_batch.argtypes = [C.c_void_p, _u64p, C.c_uint64, _u64p, C.c_void_p]
        L.ref_qf_iterate.restype = C.c_uint64
        L.ref_qf_iterate.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64]
        L.ref_qf_denoise.restype = C.c_uint64
        L.ref_qf_denoise.argtypes = [C.c_void_p]
        L.ref_qf_denoise_all_clusters.restype = C.c_uint64
        L.ref_qf_denoise_all_clusters.argtypes = [C.c_void_p]
        L.ref_qf_check.argtypes = [C.c_void_p]
        L.ref_count_mt.restype = C.c_double
        L.ref_count_mt.argtypes = [C.c_void_p, C.c_void_p, _u64p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int,
                                   C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        _ref = L
    return _ref


# ------------------------------------------------------------------ helpers
def _p(a):
    return a.ctypes.data


def reads_to_keys(text: bytes, K: int, hb: int) -> np.ndarray:
    L = oracle()
    n = L.orc_reads_to_keys(text, len(text), K, hb, None, 0)
    if n == 2**64 - 1:
        raise ValueError("malformed FASTQ chunk")
    out = np.zeros(n, np.uint64)
    L.orc_reads_to_keys(text, len(text), K, hb, _p(out), n)
    return out


def chunks(files, part_size=1 << 23):
    """Yield (file_index, chunk_bytes) in the reference's t=1 order."""
    L = oracle()
    bufs = (C.c_char_p * len(files))(*files)
    lens = (C.c_size_t * len(files))(*[len(f) for f in files])
    h = L.orc_chunker_new(len(files), bufs, lens, part_size)
    cap = part_size + 65536
    buf = C.create_string_buffer(cap)
    idx = C.c_int(0)
    try:
        while True:
            n = L.orc_chunker_next(h, buf, cap, C.byref(idx))
            if n == -1:
                return
            if n < 0:
                raise ValueError("chunker error %d" % n)
            yield idx.value, buf.raw[:n]
    finally:
        L.orc_chunker_free(h)


def chunk_ends(addr, nbytes, part_size=1 << 23):
    """Cumulative end offsets of the reference's chunks (fastq_read_parts, cqf/CQF_mt.h:735-816) over ONE in-memory
    file at `addr`; only the sizes are kept, so this works on tens of GB."""
    L = oracle()
    bufs = (C.c_char_p * 1)(C.cast(C.c_void_p(addr), C.c_char_p))
    lens = (C.c_size_t * 1)(nbytes)
    h = L.orc_chunker_new(1, bufs, lens, part_size)
    cap = part_size + 65536
    buf = C.create_string_buffer(cap)
    idx = C.c_int(0)
    ends, off = [], 0
    try:
        while True:
            n = L.orc_chunker_next(h, buf, cap, C.byref(idx))
            if n == -1:
                break
            if n < 0:
                raise ValueError("chunker error %d" % n)
            off += n
            ends.append(off)
    finally:
        L.orc_chunker_free(h)
    return np.array(ends, np.uint64)


class OracleCqf:
    def __init__(self, qb=None, hb=None, seed=2038074761, handle=None):
        self.L = oracle()
        self.h = handle if handle is not None else self.L.orc_cqf_new(qb, hb, seed)

    def close(self):
        if self.h:
            self.L.orc_cqf_free(self.h)
            self.h = None

    __del__ = close

    def insert_keys(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        return self.L.orc_cqf_insert_keys(self.h, _p(keys), len(keys))

    def insert_counted(self, keys, counts):
        keys = np.ascontiguousarray(keys, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint64)
        return self.L.orc_cqf_insert_counted(self.h, _p(keys), _p(counts), len(keys))

    def counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.orc_cqf_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def denoise(self, ideal=False):
        return (self.L.orc_cqf_denoise_ideal if ideal else self.L.orc_cqf_denoise)(self.h)

    def iterate(self):
        n = self.L.orc_cqf_iterate(self.h, None, None, 0)
        k = np.zeros(n, np.uint64)
        c = np.zeros(n, np.uint64)
        self.L.orc_cqf_iterate(self.h, _p(k), _p(c), n)
        return k, c

    def export(self) -> bytes:
        n = self.L.orc_cqf_image_size(self.h)
        buf = np.zeros(n, np.uint8)
        rc = self.L.orc_cqf_export(self.h, _p(buf), n)
        if rc:
            raise OverflowError("oracle export rc=%d" % rc)
        return buf.tobytes()

    @classmethod
    def from_bytes(cls, data: bytes):
        L = oracle()
        arr = np.frombuffer(data, np.uint8)
        h = L.orc_cqf_import(_p(arr), len(arr))
        if not h:
            raise ValueError("bad .cqf image")
        return cls(handle=h)

    def query_batch(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        cnt = np.zeros(len(keys), np.uint64)
        was = np.zeros(len(keys), np.uint8)
        self.L.orc_cqf_query_batch(self.h, _p(keys), len(keys), _p(cnt), _p(was))
        return cnt, was


def count_run(files, K, qb, hb, trigger, num_denoise, end_denoise, part_size=1 << 23):
    L = oracle()
    bufs = (C.c_char_p * len(files))(*files)
    lens = (C.c_size_t * len(files))(*[len(f) for f in files])
    rounds = C.c_int(0)
    h = L.orc_count_run(len(files), bufs, lens, part_size, K, qb, hb, trigger, num_denoise,
                        1 if end_denoise else 0, C.byref(rounds))
    return OracleCqf(handle=h), rounds.value


def param_block(K, n_true, N_total, alpha=-1.0, true2false=0.0, deNoise=-1, fr=0.0):
    out = orc_params()
    oracle().orc_param_block(K, n_true, N_total, alpha, true2false, deNoise, fr, C.byref(out))
    return out


class RefCqf:
    """The real gqf.c behind the shim (container only / prebuilt .so on the GPU box)."""

    def __init__(self, qb=None, hb=None, seed=2038074761, path=None):
        self.L = ref()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libshq_ref.so not built")
        self.h = self.L.ref_qf_load(path.encode()) if path else self.L.ref_qf_new(qb, hb, seed)

    def close(self):
        if self.h:
            self.L.ref_qf_free(self.h)
            self.h = None

    __del__ = close

    def insert_keys(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        return self.L.ref_qf_insert_keys(self.h, _p(keys), len(keys))

    def insert_counted(self, keys, counts):
        keys = np.ascontiguousarray(keys, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint64)
        return self.L.ref_qf_insert_counted(self.h, _p(keys), _p(counts), len(keys))

    def set_counts(self, ndistinct, nelts):
        self.L.ref_qf_set_counts(self.h, nelts, ndistinct)

    def denoise(self, ideal=False):
        return (self.L.ref_qf_denoise_all_clusters if ideal else self.L.ref_qf_denoise)(self.h)

    def iterate(self):
        cap = self.L.ref_qf_iterate(self.h, None, None, 0)
        k = np.zeros(cap, np.uint64)
        c = np.zeros(cap, np.uint64)
        self.L.ref_qf_iterate(self.h, _p(k), _p(c), cap)
        return k, c

    def export(self) -> bytes:
        size = self.L.ref_qf_meta(self.h, 0)
        md = C.string_at(self.L.ref_qf_metadata_ptr(self.h), 128)
        bl = C.string_at(self.L.ref_qf_blocks_ptr(self.h), size)
        return md + bl

    def save(self, path):
        self.L.ref_qf_save(self.h, path.encode())

    def query_batch(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        cnt = np.zeros(len(keys), np.uint64)
        was = np.zeros(len(keys), np.uint8)
        self.L.ref_qf_query_batch(self.h, _p(keys), len(keys), _p(cnt), _p(was))
        return cnt, was

    def check(self):
        return self.L.ref_qf_check(self.h)

    def count_mt(self, text_addr, chunk_ends, K, nthreads, trigger=0, rounds=0, end_denoise=False, max_seconds=0.0):
        """Multi-threaded reference counting over text at text_addr; returns (seconds, kmers, chunks, rounds)."""
        ce = np.ascontiguousarray(chunk_ends, np.uint64)
        km, ch, rd = C.c_uint64(), C.c_uint64(), C.c_uint32()
        secs = self.L.ref_count_mt(self.h, text_addr, _p(ce), len(ce), K, nthreads, trigger, rounds, 1 if end_denoise else 0,
                                   max_seconds, C.byref(km), C.byref(ch), C.byref(rd))
        return secs, km.value, ch.value, rd.value


# ------------------------------------------------------- synthetic inputs
def synth_genome(G, seed=42):
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, G)]


def synth_fastq(G=100000, cov=20, L=150, err=0.002, seed=42, n_rate=0.0, genome=None):
    """iid genome, uniform positions, strand flip p=0.5, substitutions with
    p=err, 4-line FASTQ with '\\n' ends and a constant quality line."""
    rng = np.random.default_rng(seed + 1)
    if genome is None:
        genome = synth_genome(G, seed)
    G = len(genome)
    nreads = int(G * cov / L)
    pos = rng.integers(0, G - L + 1, nreads)
    idx = pos[:, None] + np.arange(L)[None, :]
    reads = genome[idx]
    comp = np.zeros(256, np.uint8)
    comp[ord("A")], comp[ord("C")], comp[ord("G")], comp[ord("T")] = ord("T"), ord("G"), ord("C"), ord("A")
    flip = rng.random(nreads) < 0.5
    reads[flip] = comp[reads[flip][:, ::-1]]
    if err > 0:
        mask = rng.random(reads.shape) < err
        code = np.zeros(256, np.uint8)
        code[ord("A")], code[ord("C")], code[ord("G")], code[ord("T")] = 0, 1, 2, 3
        sub = (code[reads[mask]] + rng.integers(1, 4, mask.sum())) % 4
        reads[mask] = np.frombuffer(b"ACGT", np.uint8)[sub]
    if n_rate > 0:
        reads[rng.random(reads.shape) < n_rate] = ord("N")
    hdr_w = len(str(nreads))
    rec = 1 + 1 + hdr_w + 1 + L + 1 + 2 + L + 1  # "@r" + id + \n + seq + \n + "+\n" + qual + \n
    out = np.empty((nreads, rec), np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("r")
    ids = np.arange(nreads)
    for d in range(hdr_w):
        out[:, 2 + hdr_w - 1 - d] = ord("0") + (ids // 10**d) % 10
    o = 2 + hdr_w
    out[:, o] = ord("\n")
    out[:, o + 1:o + 1 + L] = reads
    out[:, o + 1 + L] = ord("\n")
    out[:, o + 2 + L] = ord("+")
    out[:, o + 3 + L] = ord("\n")
    out[:, o + 4 + L:o + 4 + 2 * L] = ord("I")
    out[:, o + 4 + 2 * L] = ord("\n")
    return out.tobytes()


# ------------------------------------------------------------------ Contiger
def contiger_run(cqf: "OracleCqf", files, K, s_min=2, x_min=2, x_max=1000000, master_rule=False, lifo=False, part_size=1 << 23):
    """Sequential Contiger restatement; returns (fasta_text, n_queries).  Mutates cqf's traveled bits."""
    L = oracle()
    bufs = (C.c_char_p * len(files))(*files)
    lens = (C.c_size_t * len(files))(*[len(f) for f in files])
    hb = np.frombuffer(cqf.export()[32:40], np.uint64)[0]
    r = L.orc_contiger_run(cqf.h, int(hb), len(files), bufs, lens, part_size, K, s_min, x_min, x_max,
                           (1 if master_rule else 0) | (2 if lifo else 0))
    n = L.orc_contiger_fasta_size(r)
    buf = C.create_string_buffer(n)
    L.orc_contiger_fasta(r, buf)
    nq = L.orc_contiger_queries(r)
    L.orc_contiger_free(r)
    return buf.raw[:n].decode(), nq


def parse_unitigs(fasta: str):
    """[(header_fields, seq)] from BCALM-style unitig FASTA."""
    out = []
    lines = fasta.split("\n")
    for i in range(0, len(lines) - 1, 2):
        if lines[i].startswith(">"):
            out.append((lines[i][1:].split(" "), lines[i + 1]))
    return out


_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: str) -> str:
    return s.encode().translate(_COMP)[::-1].decode()


def canonical_unitig_set(fasta: str):
    """Sorted list of min(seq, RC(seq)) -- the orientation/numbering-free view used for parity."""
    return sorted(min(s, revcomp(s)) for _, s in parse_unitigs(fasta))


def canonical_links(fasta: str, stale=None):
    """Link structure up to renumbering/orientation: set of ((canon_seq_a, end_a), (canon_seq_b, end_b)) where
    end is 'L' or 'R' of the canonical orientation.  Links whose target id is out of range are skipped and, if `stale`
    is a list, appended to it: the reference's end-k-mer map keeps entries of contigs that lost a race and were cleared
    (contig_assembly.cpp:935-1010 never removes them), so its FASTA can name ids that no unitig has."""
    us = parse_unitigs(fasta)
    canon = []
    for _, s in us:
        r = revcomp(s)
        canon.append((min(s, r), s <= r))      # (canonical seq, stored orientation is canonical?)
    links = set()
    for i, (hdr, s) in enumerate(us):
        for f in hdr:
            if f.startswith("L:"):
                _, side, j, orient = f.split(":")
                j = int(j)
                if j >= len(us):
                    if stale is not None:
                        stale.append((i, f))
                    continue
                # this unitig's end: '+' side = its right end in stored orientation
                a_end_right = (side == "+")
                a = (canon[i][0], "R" if a_end_right == canon[i][1] else "L")
                # neighbour entered at its left end if orient '+', right end if '-'
                b_left = (orient == "+")
                b = (canon[j][0], "L" if b_left == canon[j][1] else "R")
                links.add((a, b))
    return links
