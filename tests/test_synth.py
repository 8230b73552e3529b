"""The synthetic read generator (C twin; the CUDA twin is compared with it in tests/test_gpu_fullsize.py)."""
import numpy as np

import synth


def test_records_are_wellformed_and_deterministic():
    sp = synth.make(50000, 20, 150, 0.01, 3)
    assert sp.n_reads == 50000 * 20 // 150 and sp.id_width == len(str(sp.n_reads))
    assert sp.rec_bytes == 2 + sp.id_width + 1 + 150 + 1 + 2 + 150 + 1
    a = synth.cpu_fill(sp, 0, sp.n_reads, threads=1)
    b = synth.cpu_fill(sp, 0, sp.n_reads, threads=4)
    assert a.tobytes() == b.tobytes()
    # any slice equals the same records of the whole set
    c = synth.cpu_fill(sp, 1000, 37)
    assert c.tobytes() == a[1000 * sp.rec_bytes:1037 * sp.rec_bytes].tobytes()
    lines = a.tobytes().split(b"\n")
    assert lines[-1] == b"" and (len(lines) - 1) == 4 * sp.n_reads
    assert lines[0] == b"@r" + b"0" * sp.id_width and lines[2] == b"+" and lines[3] == b"I" * 150
    assert set(lines[1]) <= set(b"ACGT")
    assert lines[4 * 1234] == b"@r%0*d" % (sp.id_width, 1234)


def test_reads_come_from_the_genome_with_the_asked_error_rate():
    sp = synth.make(200000, 30, 100, 0.02, 9)
    g = synth.genome(sp).tobytes()
    t = synth.cpu_fill(sp, 0, 3000).reshape(3000, sp.rec_bytes)
    seqs = t[:, 2 + sp.id_width + 1: 2 + sp.id_width + 1 + 100]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    mism = tot = flips = found = 0
    for s in seqs[:400]:
        s = s.tobytes()
        for strand, q in ((0, s), (1, s.translate(comp)[::-1])):
            # an error-free 24-mer somewhere in the read locates it
            hit = -1
            for off in (0, 25, 50, 75):
                j = g.find(q[off:off + 24])
                if j >= off:
                    hit = j - off
                    break
            if hit >= 0 and hit + 100 <= len(g):
                ref = g[hit:hit + 100]
                mism += sum(x != y for x, y in zip(q, ref))
                tot += 100
                flips += strand
                found += 1
                break
    assert found > 380
    assert 0.3 < flips / found < 0.7
    assert 0.012 < mism / tot < 0.028


def test_rank_slices_partition_the_set():
    sp = synth.make(10000, 10, 150, 0.0, 1)
    got = []
    for r in range(4):
        first, n = synth.rank_slice(sp, r, 4)
        got.append(synth.cpu_fill(sp, first, n))
    assert np.concatenate(got).tobytes() == synth.cpu_fill(sp, 0, sp.n_reads).tobytes()
