"""The BASELINE.json configurations as synthetic workloads; shared by bench.py, the full-size GPU
tests and tests/golden/make_digests.py so that all three talk about the same bytes."""

# name -> genome length, coverage, read length, substitution rate, K, -n, --deNoise (-1 = automatic), --endDeNoise
WORKLOADS = {
    # BASELINE configs[0]: E. coli-size, --deNoise=0
    "ecoli0": dict(G=4_600_000, cov=100, L=150, err=0.002, K=31, n_true=4_600_000, deNoise=0, end=False,
                   title="BASELINE configs[0]"),
    # BASELINE configs[1]: same reads, --deNoise=2 --endDeNoise (effective: 1 mid-stream round + the end round)
    "ecoli": dict(G=4_600_000, cov=100, L=150, err=0.002, K=31, n_true=4_600_000, deNoise=2, end=True,
                  title="BASELINE configs[1]"),
    # BASELINE configs[2]: 100 Mb C. elegans-scale genome, K=47, -n 1.2e8 -e 0.00234, automatic rounds
    "celegans": dict(G=100_000_000, cov=100, L=150, err=0.00234, K=47, n_true=120_000_000, deNoise=-1, end=False,
                     title="BASELINE configs[2]"),
    # 20 Mb twin of configs[2] (full-size parity test that the t=1 oracle finishes in minutes)
    "celegans20": dict(G=20_000_000, cov=100, L=150, err=0.00234, K=47, n_true=24_000_000, deNoise=-1, end=False,
                       title="20 Mb twin of BASELINE configs[2]"),
    # 1/8 of BASELINE configs[3] per GPU: human-scale 30x, K=31, --deNoise=0 (the sharded run multiplies G by N)
    "human8": dict(G=375_000_000, cov=30, L=150, err=0.002, K=31, n_true=375_000_000, deNoise=0, end=False,
                   title="1/8 of BASELINE configs[3] per GPU"),
}
SEED = 42


def scaled(name, world=1):
    """Workload `name` grown for `world` shards (weak scaling: genome and -n grow with the number of GPUs)."""
    w = dict(WORKLOADS[name])
    w["G"] *= world
    w["n_true"] *= world
    w["name"] = name
    return w


def n_kmers(w):
    nreads = int(w["G"] * w["cov"] / w["L"])
    return nreads * (w["L"] - w["K"] + 1)


def describe(w, p):
    return ("%s: synthetic %.3g Mb genome, %dx %d bp reads, %.3g%% subst, k=%d, -n %d -N %d -e %g --deNoise=%d%s "
            "(effective %d mid-stream round(s)%s), qb=%d hb=%d"
            % (w["title"], w["G"] / 1e6, w["cov"], w["L"], 100 * w["err"], w["K"], w["n_true"], n_kmers(w), w["err"], w["deNoise"],
               " --endDeNoise" if w["end"] else "", p.num_deNoise, " + end round" if w["end"] else "", p.qb, p.hb))
