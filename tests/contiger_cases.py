"""The Contiger parity cases, shared by the golden-fixture generator (tests/golden/make_contiger_golden.py: the UNMODIFIED
reference Contiger, oracle/_ref/Contiger_ref, at -t 1), the CPU suite (oracle vs that output) and the GPU suite."""
import hashlib

import numpy as np


def cases(po):
    """name -> dict(files, K, qb, s, x, X, part)"""
    out = {}
    for K, G, cov, err, L, seed in [(21, 20000, 40, 0.003, 100, 5), (31, 60000, 30, 0.004, 150, 7), (47, 50000, 40, 0.002, 150, 9),
                                    (25, 30000, 60, 0.01, 100, 11), (33, 40000, 30, 0.003, 120, 13)]:
        out["k%d_g%d_s%d" % (K, G, seed)] = dict(files=[po.synth_fastq(G=G, cov=cov, err=err, L=L, seed=seed)], K=K, qb=19, s=2, x=2, X=1000000, part=1 << 23)
    g = po.synth_genome(30000, seed=17)
    rep = g[1000:1400].copy()
    genome = np.concatenate([g[:8000], rep, g[8000:16000], rep[::-1].copy(), g[16000:], rep])   # planted repeats
    out["repeats_two_files"] = dict(files=[po.synth_fastq(cov=35, err=0.003, L=100, seed=19 + i, genome=genome) for i in range(2)],
                                    K=27, qb=19, s=2, x=2, X=1000000, part=1 << 20)
    out["thresholds"] = dict(files=[po.synth_fastq(G=30000, cov=30, err=0.005, L=100, seed=23)], K=23, qb=19, s=3, x=4, X=60, part=1 << 23)
    out["invariants_k31"] = dict(files=[po.synth_fastq(G=50000, cov=30, err=0.003, seed=29)], K=31, qb=19, s=2, x=2, X=1000000, part=1 << 23)
    # many false-positive neighbours: a small filter (hb = 25, about 40 % full) under a 30 kb genome
    out["collisions_qb17"] = dict(files=[po.synth_fastq(G=30000, cov=25, err=0.002, L=120, seed=31)], K=29, qb=17, s=2, x=2, X=1000000, part=1 << 23)
    return out


def unitig_digest(po, fasta):
    """(number of unitigs, sha256 of the sorted canonical sequences, sha256 of the sorted canonical links)"""
    us = po.canonical_unitig_set(fasta)
    links = sorted(po.canonical_links(fasta))
    return {"n_unitigs": len(us), "total_bases": sum(len(u) for u in us),
            "set_sha256": hashlib.sha256("\n".join(us).encode()).hexdigest(),
            "links_sha256": hashlib.sha256(repr(links).encode()).hexdigest(), "n_links": len(links)}
