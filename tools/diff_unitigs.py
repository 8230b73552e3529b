"""Debug helper (GPU box): run oracle and device Contiger on a synthetic case and print the differences."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
shq = importlib.import_module("sh-assembly_b200")

def run(K, G, cov, err, L, seed, qb=19):
    files = [po.synth_fastq(G=G, cov=cov, err=err, L=L, seed=seed)]
    o, _ = po.count_run(files, K, qb, qb + 8, 0, 0, True)
    img = o.export()
    fa_o, nq = po.contiger_run(o, files, K)
    g = shq.CQF.from_bytes(img)
    chunks = [c for _, c in shq.Reader(buffers=files)]
    fa_g = shq.build_unitigs(g, chunks, K)
    so, sg = set(po.canonical_unitig_set(fa_o)), set(po.canonical_unitig_set(fa_g))
    print("case", K, G, cov, err, L, seed, "oracle", len(so), "gpu", len(sg), "only_oracle", len(so - sg), "only_gpu", len(sg - so))
    o2 = po.OracleCqf.from_bytes(img)
    def cnt(k):
        import numpy as np
        keys = po.reads_to_keys(b"@x\n" + k.encode() + b"\n+\n" + b"I" * len(k) + b"\n", K, qb + 8)
        return [int(x) for x in o2.query_batch(keys)[0]]
    for name, d in (("only_oracle", so - sg), ("only_gpu", sg - so)):
        for s in sorted(d)[:6]:
            print(" ", name, len(s), s[:80], "counts", cnt(s)[:8], "..")
    lo, lg = po.canonical_links(fa_o), po.canonical_links(fa_g)
    print("  links: oracle", len(lo), "gpu", len(lg), "only_oracle", len(lo - lg), "only_gpu", len(lg - lo))

if __name__ == "__main__":
    for case in [(21, 20000, 40, 0.003, 100, 5), (31, 60000, 30, 0.004, 150, 7), (47, 50000, 40, 0.002, 150, 9), (25, 30000, 60, 0.01, 100, 11), (33, 40000, 30, 0.003, 120, 13)]:
        run(*case)
