"""Debug helper (GPU box): E. coli-scale Contiger parity, device vs sequential oracle."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
shq = importlib.import_module("sh-assembly_b200")
import bench

def main(G=4600000, cov=100, K=31, err=0.002):
    text, _, nreads = bench.gen_fastq(G, cov, 150, err, 42)
    n_total = nreads * (150 - K + 1)
    p = shq.param_block(K, G, n_total, alpha=err, deNoise=2)
    q = shq.CQF(int(p.qb), int(p.hb))
    q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), True)
    chunks = []
    for _, ch in shq.Reader(buffers=[text]):
        chunks.append(ch)
        q.insert_fastq_chunk(ch, K)
    q.finish()
    img = q.export()
    t0 = time.time(); fa_g = shq.build_unitigs(q, chunks, K); tg = time.time() - t0
    o = po.OracleCqf.from_bytes(img)
    t0 = time.time(); fa_o, nq = po.contiger_run(o, [text.tobytes()], K); to = time.time() - t0
    so, sg = set(po.canonical_unitig_set(fa_o)), set(po.canonical_unitig_set(fa_g))
    print("oracle", len(so), "gpu", len(sg), "only_oracle", len(so - sg), "only_gpu", len(sg - so), "t_gpu %.2f t_cpu %.2f" % (tg, to))
    o2 = po.OracleCqf.from_bytes(img)
    hb = int(p.hb)
    def cnt(k):
        keys = po.reads_to_keys(b"@x\n" + k.encode() + b"\n+\n" + b"I" * len(k) + b"\n", K, hb)
        return [int(x) for x in o2.query_batch(keys)[0]]
    tx = text.tobytes()
    for name, d in (("only_oracle", so - sg), ("only_gpu", sg - so)):
        for s in sorted(d)[:8]:
            print(" ", name, len(s), s[:100], "counts", cnt(s)[:6])
            for nm, k in (("first", s[:K]), ("last", s[-K:])):
                for kk, tag in ((k, "fwd"), (po.revcomp(k), "rc")):
                    succ = [(x, cnt(kk[1:] + x)[0]) for x in "ACGT"]
                    print("     ", nm, tag, "in reads:", tx.count(kk.encode()) + tx.count(po.revcomp(kk).encode()), "succ", succ)
    lo, lg = po.canonical_links(fa_o), po.canonical_links(fa_g)
    print("links: oracle", len(lo), "gpu", len(lg), "only_oracle", len(lo - lg), "only_gpu", len(lg - lo))

if __name__ == "__main__":
    main()
