"""Parity of the Contiger device path against the sequential oracle restatement.  -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_both(po, shq, files, K, qb, s=2, x=2, X=1000000, part=1 << 23):
    o, _ = po.count_run(files, K, qb, qb + 8, 0, 0, True, part_size=part)
    img = o.export()
    fa_o, nq = po.contiger_run(o, files, K, s, x, X, part_size=part)
    g = shq.CQF.from_bytes(img)
    chunks = [c for _, c in shq.Reader(buffers=files, part_size=part)]
    fa_g = shq.build_unitigs(g, chunks, K, s, x, X)
    g.close()
    return fa_o, fa_g


@pytest.mark.parametrize("K,G,cov,err,L,seed", [(21, 20000, 40, 0.003, 100, 5), (31, 60000, 30, 0.004, 150, 7), (47, 50000, 40, 0.002, 150, 9),
                                               (25, 30000, 60, 0.01, 100, 11), (33, 40000, 30, 0.003, 120, 13)])
def test_unitigs_match_oracle(po, shq, K, G, cov, err, L, seed):
    files = [po.synth_fastq(G=G, cov=cov, err=err, L=L, seed=seed)]
    fa_o, fa_g = run_both(po, shq, files, K, 19)
    so, sg = po.canonical_unitig_set(fa_o), po.canonical_unitig_set(fa_g)
    assert len(sg) == len(so), (len(sg), len(so))
    assert sg == so
    assert po.canonical_links(fa_g) == po.canonical_links(fa_o)


def test_unitigs_with_repeats_and_two_files(po, shq):
    rng = np.random.default_rng(3)
    g = po.synth_genome(30000, seed=17)
    rep = g[1000:1400].copy()
    genome = np.concatenate([g[:8000], rep, g[8000:16000], rep[::-1].copy(), g[16000:], rep])   # planted repeats
    files = [po.synth_fastq(cov=35, err=0.003, L=100, seed=19 + i, genome=genome) for i in range(2)]
    fa_o, fa_g = run_both(po, shq, files, 27, 19, part=1 << 20)
    assert po.canonical_unitig_set(fa_g) == po.canonical_unitig_set(fa_o)
    assert po.canonical_links(fa_g) == po.canonical_links(fa_o)


def test_unitigs_thresholds(po, shq):
    files = [po.synth_fastq(G=30000, cov=30, err=0.005, L=100, seed=23)]
    fa_o, fa_g = run_both(po, shq, files, 23, 19, s=3, x=4, X=60)
    assert po.canonical_unitig_set(fa_g) == po.canonical_unitig_set(fa_o)


def test_unitig_invariants(po, shq):
    K = 31
    files = [po.synth_fastq(G=50000, cov=30, err=0.003, seed=29)]
    fa_o, fa_g = run_both(po, shq, files, K, 19)
    us = po.parse_unitigs(fa_g)
    seqs = [s for _, s in us]
    assert len(set(min(s, po.revcomp(s)) for s in seqs)) == len(seqs)          # no duplicate unitig
    kmers = set()
    for s in seqs:
        for i in range(len(s) - K + 1):
            k = s[i:i + K]
            c = min(k, po.revcomp(k))
            assert c not in kmers                                                # every k-mer in exactly one unitig
            kmers.add(c)
    for hdr, s in us:
        ln = int([f for f in hdr if f.startswith("LN:i:")][0][5:])
        assert ln == len(s)


def test_unitigs_irregular_chunk_host_cursor(po, shq):
    """A seed with 'N' whose quality line starts with '@' derails the reference's line cursor for one record
    (contig_assembly.cpp:1868-1871 `continue`s without skipLines); the device path detects it and replays the
    literal host cursor for that chunk."""
    K = 25
    text = po.synth_fastq(G=30000, cov=30, err=0.003, L=100, seed=41)
    recs = text.split(b"\n")
    out = []
    for i in range(0, len(recs) - 1, 4):
        h, s, p, q = recs[i:i + 4]
        if (i // 4) % 37 == 5:
            s = s[:50] + b"N" + s[51:]
            q = b"@" + q[1:]
        out += [h, s, p, q]
    text2 = b"\n".join(out) + b"\n"
    fa_o, fa_g = run_both(po, shq, [text2], K, 19)
    assert po.canonical_unitig_set(fa_g) == po.canonical_unitig_set(fa_o)
    assert po.canonical_links(fa_g) == po.canonical_links(fa_o)


# ---- against OUTPUT OF THE UNMODIFIED REFERENCE CONTIGER (tests/golden/contiger_ref.json, see tests/test_contiger_ref.py) ----
import json
import os

from tests import contiger_cases as cc

_GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contiger_ref.json")))


@pytest.mark.parametrize("name", sorted(n for n in _GOLD if n != "repeats_two_files"))
def test_unitigs_match_reference_binary_output(po, shq, name):
    """one read chunk: the reference applies ONE seed rule to all reads (its master thread's, len/2, in almost every run);
    the device path under that rule gives the reference's canonical unitig set and link structure"""
    case = cc.cases(po)[name]
    o, _ = po.count_run(case["files"], case["K"], case["qb"], case["qb"] + 8, 0, 0, True, part_size=case["part"])
    img = o.export()
    chunks = [c for _, c in shq.Reader(buffers=case["files"], part_size=case["part"])]
    digests = []
    for master in (True, False):
        g = shq.CQF.from_bytes(img)
        fa = shq.build_unitigs(g, chunks, case["K"], case["s"], case["x"], case["X"], master_rule=master)
        g.close()
        digests.append(cc.unitig_digest(po, fa))
    outcomes = _GOLD[name]["reference_outcomes"]
    assert outcomes and all(d in digests for d in outcomes)
    assert digests[0] in outcomes or digests[1] in outcomes


def test_unitigs_two_chunks_between_the_seed_rules(po, shq):
    """several chunks: master and worker race for them in the reference, its output varies between the two rules' results;
    the device path reproduces both ends of that range (links are the same under both)"""
    name = "repeats_two_files"
    case = cc.cases(po)[name]
    o, _ = po.count_run(case["files"], case["K"], case["qb"], case["qb"] + 8, 0, 0, True, part_size=case["part"])
    img = o.export()
    chunks = [c for _, c in shq.Reader(buffers=case["files"], part_size=case["part"])]
    ds = []
    for master in (True, False):
        g = shq.CQF.from_bytes(img)
        fa = shq.build_unitigs(g, chunks, case["K"], case["s"], case["x"], case["X"], master_rule=master)
        g.close()
        ds.append(cc.unitig_digest(po, fa))
    for d in _GOLD[name]["reference_outcomes"]:
        assert min(x["n_unitigs"] for x in ds) - 2 <= d["n_unitigs"] <= max(x["n_unitigs"] for x in ds) + 2
        assert d["links_sha256"] in [x["links_sha256"] for x in ds]
