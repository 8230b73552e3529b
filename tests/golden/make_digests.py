#!/usr/bin/env python
"""Full-size digests of the BASELINE configurations, produced by the UNMODIFIED reference core
(oracle/_ref: cqf/gqf.c + base/nthash.hpp) at t = 1 on the synthetic read sets of synth/workloads.py.

    python tests/golden/make_digests.py ecoli0 ecoli celegans20 celegans

Writes / updates tests/golden/digests.json: per workload the parameter block, the runtime counters, the number of
deNoise sweeps and sha256 of the serialised .cqf (128-byte header + blocks) and of the FASTQ text.  bench.py and the
full-size `-m gpu` tests compare the CUDA path with these.  Runs only where /root/reference exists (this container);
configs[2] takes about an hour of one core."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from synth import workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "digests.json")


def digest(name, world=1):
    w = wl.scaled(name, world)
    sp = synth.make(w["G"], w["cov"], w["L"], w["err"], wl.SEED)
    N = wl.n_kmers(w)
    p = po.param_block(w["K"], w["n_true"], N, alpha=w["err"], deNoise=w["deNoise"])
    t0 = time.time()
    if world == 1:
        text = synth.cpu_fill(sp, 0, sp.n_reads)
        ends = po.chunk_ends(text.ctypes.data, text.nbytes)
    else:
        # `world` read files, one per rank; at t = 1 the reference takes their chunks round robin
        # (seqFile_batch::getDataChunk, cqf/CQF_mt.h:364-390): the same job as the chunk-interleaved text below
        import numpy as np
        files, cuts = [], []
        for r in range(world):
            first, n = synth.rank_slice(sp, r, world)
            f = synth.cpu_fill(sp, first, n)
            e = po.chunk_ends(f.ctypes.data, f.nbytes)
            files.append(f)
            cuts.append([0] + [int(x) for x in e])
        pieces = []
        for j in range(max(len(c) - 1 for c in cuts)):
            for r in range(world):
                if j + 1 < len(cuts[r]):
                    pieces.append(files[r][cuts[r][j]:cuts[r][j + 1]])
        text = np.concatenate(pieces)
        ends = np.cumsum([len(x) for x in pieces]).astype(np.uint64)
        del files, pieces
    print("%s: %d reads, %.2f GB of text, %d chunks, qb=%d rounds=%d trigger=%d (%.0f s)" % (
        name, sp.n_reads, text.nbytes / 1e9, len(ends), p.qb, p.num_deNoise, p.trigger, time.time() - t0), flush=True)
    th = hashlib.sha256(memoryview(text)).hexdigest()
    r = po.RefCqf(int(p.qb), int(p.hb))
    secs, kmers, chunks, rounds = r.count_mt(text.ctypes.data, ends, w["K"], 1, trigger=int(p.trigger), rounds=int(p.num_deNoise),
                                             end_denoise=w["end"])
    assert chunks == len(ends) and kmers == N, (chunks, len(ends), kmers, N)
    nd, ne = r.L.ref_qf_meta(r.h, 10), r.L.ref_qf_meta(r.h, 9)
    img = r.export()
    r.close()
    d = {"workload": name, "world": world, "seed": wl.SEED, "genome_bp": w["G"], "coverage": w["cov"], "read_len": w["L"], "error_rate": w["err"],
         "K": w["K"], "n_true": w["n_true"], "n_kmers": N, "deNoise": w["deNoise"], "endDeNoise": w["end"],
         "qb": int(p.qb), "hb": int(p.hb), "num_deNoise": int(p.num_deNoise), "trigger": int(p.trigger),
         "n_chunks": int(len(ends)), "text_bytes": int(text.nbytes), "text_sha256": th,
         "ndistinct": int(nd), "nelts": int(ne), "denoise_rounds": int(rounds), "image_bytes": len(img),
         "image_sha256": hashlib.sha256(img).hexdigest(),
         "made_by": "oracle/_ref (unmodified gqf.c + nthash.hpp), t=1, %.0f s" % secs}
    print(json.dumps(d), flush=True)
    return d


def main():
    names = sys.argv[1:] or ["ecoli0", "ecoli", "celegans20"]
    for nm in names:
        name, _, world = nm.partition("@")
        d = digest(name, int(world or 1))
        cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
        cur[nm] = d
        json.dump(cur, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
