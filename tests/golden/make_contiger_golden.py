#!/usr/bin/env python
"""Golden fixtures for the Contiger path, produced by the UNMODIFIED reference Contiger
(/root/reference/src/contig_assembly.cpp compiled against oracle/shim into oracle/_ref/Contiger_ref, `make -C oracle
contiger_ref`) at -t 1 on the cases of tests/contiger_cases.py.  The .cqf it loads is written by the unmodified
reference core (oracle/_ref/libshq_ref.so, qf_serialize).

    python tests/golden/make_contiger_golden.py            # -> tests/golden/contiger_ref.json

Each case is run several times: at -t 1 the reference still has two consumers of read chunks (the master, seed rule
len/2, and one worker, seed rule len/2 - K/2; contig_assembly.cpp:2067-2071 vs 1859-1863) racing for the chunks, so
its output can differ between runs at the level of small islands.  The fixture records every distinct outcome."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
from tests import contiger_cases as cc  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "Contiger_ref")
OUT = os.path.join(ROOT, "tests", "golden", "contiger_ref.json")


def run_reference(case, tmp, runs=3):
    files, K = case["files"], case["K"]
    r = po.RefCqf(case["qb"], case["qb"] + 8)
    for f in files:
        r.insert_keys(po.reads_to_keys(f, K, case["qb"] + 8))
    # CQF-deNoise --endDeNoise on every case (as tests/test_gpu_unitigs.py does): count-1 keys are swept
    removed = r.denoise()
    k, c = r.iterate()
    r.set_counts(len(k), int(c.sum()))
    cqf = os.path.join(tmp, "t.cqf")
    r.save(cqf)
    image = open(cqf, "rb").read()
    r.close()
    paths = []
    for i, f in enumerate(files):
        p = os.path.join(tmp, "reads%d.fq" % i)
        open(p, "wb").write(f)
        paths.append(p)
    lst = os.path.join(tmp, "list.txt")
    open(lst, "w").write("\n".join(paths) + "\n")
    outs = []
    for i in range(runs):
        fa = os.path.join(tmp, "ref%d.fa" % i)
        subprocess.run([BIN, "-k", str(K), "-i", lst, "-c", cqf, "-s", str(case["s"]), "-x", str(case["x"]), "-X", str(case["X"]), "-t", "1", "-o", fa],
                       check=True, capture_output=True, timeout=300)
        outs.append(open(fa).read())
    return image, outs


def main():
    gold = {}
    for name, case in cc.cases(po).items():
        with tempfile.TemporaryDirectory() as tmp:
            image, outs = run_reference(case, tmp)
        digests = []
        for fa in outs:
            d = cc.unitig_digest(po, fa)
            if d not in digests:
                digests.append(d)
        gold[name] = {"K": case["K"], "qb": case["qb"], "s": case["s"], "x": case["x"], "X": case["X"], "reference_outcomes": digests}
        print(name, digests, flush=True)
    json.dump(gold, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
