"""CPU suite: the sequential Contiger restatement on hand-checkable graphs, and its invariance under
visiting order (the reference is racy; only the canonical unitig set is stable -- SURVEY.md section 7, H2)."""
import numpy as np


def fastq_of(reads):
    return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))


def build(po, reads, K, qb=22):
    text = fastq_of(reads)
    o, _ = po.count_run([text], K, qb, qb + 8, 0, 0, False)
    return text, o


def tile(genome, L, step):
    return [genome[i:i + L] for i in range(0, len(genome) - L + 1, step)] * 2     # every k-mer seen >= 2x


def test_linear_genome_is_one_unitig(po):
    K = 21
    g = po.synth_genome(3000, seed=3).tobytes()
    text, o = build(po, tile(g, 100, 7) + [g[-100:]] * 2, K)
    fa, _ = po.contiger_run(o, [text], K)
    us = po.canonical_unitig_set(fa)
    assert us == [min(g.decode(), po.revcomp(g.decode()))]
    hdr, seq = po.parse_unitigs(fa)[0]
    assert hdr[1] == "LN:i:3000" and not any(f.startswith("L:") for f in hdr)


def test_repeat_splits_unitigs_and_links(po):
    K = 21
    a, b, c, r = (po.synth_genome(n, seed=s).tobytes().decode() for n, s in ((400, 1), (400, 2), (400, 3), (60, 4)))
    genome = (a + r + b + r + c).encode()          # the repeat r (>= K) joins two contexts
    text, o = build(po, tile(genome, 120, 5) + [genome[-120:]] * 2, K)
    fa, _ = po.contiger_run(o, [text], K)
    us = po.canonical_unitig_set(fa)
    assert len(us) == 4                              # a-side, shared repeat core, b (between the copies), c-side
    total_kmers = sum(len(u) - K + 1 for u in us)
    assert total_kmers == len({min(genome[i:i + K].decode(), po.revcomp(genome[i:i + K].decode())) for i in range(len(genome) - K + 1)})
    links = po.canonical_links(fa)
    assert len(links) >= 8 and all((y, x) in links for (x, y) in links)           # every link is seen from both ends


def test_visiting_order_invariance(po):
    K = 25
    text = po.synth_fastq(G=30000, cov=40, err=0.004, L=100, seed=77)
    res = []
    for master in (False, True):
        for lifo in (False, True):
            o, _ = po.count_run([text], K, 19, 27, 0, 0, True)
            fa, _ = po.contiger_run(o, [text], K, master_rule=master, lifo=lifo)
            res.append((po.canonical_unitig_set(fa), po.canonical_links(fa)))
    assert res[0] == res[1]                          # FIFO vs LIFO (worker rule)
    assert res[2] == res[3]                          # FIFO vs LIFO (master rule)


def test_seed_thresholds(po):
    K = 21
    g = po.synth_genome(2000, seed=9).tobytes()
    reads = tile(g, 100, 9) + [g[-100:]] * 2
    text, o = build(po, reads, K)
    fa, _ = po.contiger_run(o, [text], K, s_min=2, x_min=2, x_max=3)      # counts are far above X=3: no seed accepted
    assert po.parse_unitigs(fa) == []
    o2 = build(po, reads, K)[1]
    fa2, _ = po.contiger_run(o2, [text], K, s_min=2, x_min=2, x_max=100000)
    assert len(po.parse_unitigs(fa2)) == 1
