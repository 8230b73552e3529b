"""CPU suite: the oracle restatement against the golden vectors generated from the reference
(tests/golden/golden.json) and, when it is available, against the reference core itself."""
import base64
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from tests.golden import gen

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_nthash_golden(po):
    O = po.oracle()
    for case in GOLD["nthash"]:
        b = base64.b64decode(case["seq_b64"])
        K = case["K"]
        fh, rh = C.c_uint64(), C.c_uint64()
        O.orc_ntpc64_base(b, K, C.byref(fh), C.byref(rh))
        assert [fh.value, rh.value] == case["fh_rh"][0]
        for j, i in enumerate(range(K, K + len(case["fh_rh"]) - 1)):
            O.orc_ntpc64_roll(b[i - K], b[i], K, C.byref(fh), C.byref(rh))
            assert [fh.value, rh.value] == case["fh_rh"][j + 1]


def test_nthash_survey_kats(po):
    """The vectors quoted in SURVEY.md 8(c)."""
    O = po.oracle()
    s = b"ACGTTGCATGCAAGCTTAGCCGATCGATTACGGATCCGTAACGTTAGCAAGGCTTAACC"
    fh, rh = C.c_uint64(), C.c_uint64()
    O.orc_ntpc64_base(s, 31, C.byref(fh), C.byref(rh))
    assert (fh.value, rh.value) == (0xC0F982981807387F, 0x4544D7CBA3257150)
    O.orc_ntpc64_roll(s[0], s[31], 31, C.byref(fh), C.byref(rh))
    assert (fh.value, rh.value) == (0x6B2239DAAC1CAE02, 0x6EA0C5CC78076AE2)
    assert min(fh.value, rh.value) & ((1 << 34) - 1) == 0x2AC1CAE02
    O.orc_ntpc64_base(s, 47, C.byref(fh), C.byref(rh))
    assert (fh.value, rh.value) == (0x39117C2C72AE163B, 0x6B97016E556105B0)
    O.orc_ntpc64_base(b"ACGTNACGTACGTACGTACGTACGTACGTAC", 31, C.byref(fh), C.byref(rh))
    assert (fh.value, rh.value) == (0xF340E528A35A362D, 0x4611D2EA27A27DBF)


def test_cqf_hand_case(po):
    case = GOLD["cqf"][0]
    o = po.OracleCqf(case["qb"], case["hb"])
    for k, c in case["inserts"]:
        o.insert_counted(np.array([k], np.uint64), np.array([c], np.uint64))
    img = o.export()
    assert img == base64.b64decode(case["image_b64"])
    # the layout quoted in SURVEY.md 8(c)
    blk0 = img[128:128 + 89]
    assert blk0[0] == 0
    assert int.from_bytes(blk0[1:9], "little") == 0x38 and int.from_bytes(blk0[9:17], "little") == 0xA200
    assert blk0[25 + 3:25 + 16] == bytes.fromhex("000002" "05" "1001" "f0" "80008101" "0101")
    k, c = o.iterate()
    assert k.tolist() == case["iter_keys"] and c.tolist() == case["iter_counts"]
    qc, qw = o.query_batch(np.array(case["query_keys"], np.uint64))
    assert qc.tolist() == case["query_counts"] and qw.tolist() == case["query_was"]
    assert o.export() == base64.b64decode(case["image_after_query_b64"])
    o = po.OracleCqf(case["qb"], case["hb"])     # deNoise happens while all traveled bits are 0
    for k, c in case["inserts"]:
        o.insert_counted(np.array([k], np.uint64), np.array([c], np.uint64))
    assert o.denoise(ideal=True) == case["removed_ideal"]
    assert o.export() == base64.b64decode(case["image_after_denoise_b64"])


@pytest.mark.parametrize("case", GOLD["cqf"][1:], ids=lambda c: c["name"])
def test_cqf_seeded_golden(po, case):
    keys = gen.multiset(case["seed"], case["hb"], case["n_base"], case["n_single"])
    assert len(keys) == case["n_keys"]
    o = po.OracleCqf(case["qb"], case["hb"])
    assert o.insert_keys(keys) == case["ndistinct"]
    img = o.export()
    assert hashlib.sha256(img).hexdigest() == case["image_sha256"]
    if "image_b64" in case:
        assert img == base64.b64decode(case["image_b64"])
    # import(export) is the identity on the multiset
    k1, c1 = o.iterate()
    k2, c2 = po.OracleCqf.from_bytes(img).iterate()
    assert (k1 == k2).all() and (c1 == c2).all()
    assert int(c1.sum()) == case["n_keys"] and len(k1) == case["ndistinct"]
    for rnd in case["denoise_rounds"]:
        assert o.denoise() == rnd["removed"]
        assert hashlib.sha256(o.export()).hexdigest() == rnd["image_sha256"]
        k, c = o.iterate()
        assert k[c == 1].tolist()[:64] == rnd["surviving_singletons"]
    qk = np.concatenate([keys[:200], gen.splitmix64(5, 200) & np.uint64((1 << case["hb"]) - 1)])
    qc, qw = o.query_batch(qk)
    qc2, qw2 = o.query_batch(qk)
    assert hashlib.sha256(qc.tobytes() + qw.tobytes() + qc2.tobytes() + qw2.tobytes()).hexdigest() == case["query_sha256"]
    assert hashlib.sha256(o.export()).hexdigest() == case["image_after_query_sha256"]


def test_range_end_quirk_is_pinned():
    """At least one golden case exercises the reference's surviving one-slot cluster."""
    assert any(r["surviving_singletons"] for c in GOLD["cqf"][1:] for r in c["denoise_rounds"])


def test_order_independence_and_denoise_equals_rebuild(po):
    keys = gen.multiset(77, 21, 3000, 4000)
    a, b = po.OracleCqf(13, 21), po.OracleCqf(13, 21)
    a.insert_keys(keys)
    b.insert_keys(keys[::-1].copy())
    assert a.export() == b.export()
    a.denoise(ideal=True)
    k, c = b.iterate()
    keep = c > 1
    r = po.OracleCqf(13, 21)
    r.insert_counted(k[keep], c[keep])
    nd, ne = a.counts()
    r.L.orc_cqf_set_counts(r.h, nd, ne)
    assert a.export() == r.export()


def test_reads_to_keys_rules(po):
    K, hb = 5, 20
    mask = (1 << hb) - 1
    O = po.oracle()

    def kmer(s):
        fh, rh = C.c_uint64(), C.c_uint64()
        O.orc_ntpc64_base(s, K, C.byref(fh), C.byref(rh))
        return min(fh.value, rh.value) & mask

    def rec(s):
        return b"@x\n" + s + b"\n+\n" + b"I" * len(s) + b"\n"

    # no N test on the first K bases; N at index >= K restarts after it; short tails are dropped
    s = b"ANGTACGTNACGTTNACG"
    want = [kmer(s[i:i + K]) for i in range(0, 4)] + [kmer(s[9:14])]
    got = po.reads_to_keys(rec(s), K, hb)
    assert got.tolist() == want
    # lower-case n does not restart (only upper-case 'N' is tested, CQF_mt.h:680)
    s2 = b"ACGTAnCGTACG"
    assert len(po.reads_to_keys(rec(s2), K, hb)) == len(s2) - K + 1
    # reads shorter than K and empty reads yield nothing; several records
    assert len(po.reads_to_keys(rec(b"ACG") + rec(b"") + rec(b"ACGTA"), K, hb)) == 1
    with pytest.raises(ValueError):
        po.reads_to_keys(b"@x\nACGTACGT\n+\n", K, hb)


def test_chunker_rules(po):
    text = po.synth_fastq(G=50000, cov=40, seed=3)           # ~4.4 MB
    part = 1 << 20
    chunks = [c for _, c in po.chunks([text], part_size=part)]
    assert b"".join(chunks) == text
    assert len(chunks) >= 4
    for c in chunks[:-1]:
        assert c[:1] == b"@" and c[-1:] == b"\n" and c.count(b"\n") % 4 == 0
        assert part - 32767 - 400 <= len(c) <= part + 65535
    # exact multiple of part_size: one extra (tail-only) chunk, as feof() only turns true on a short read
    t2 = text[: 2 * part]
    t2 = t2[: t2.rfind(b"\n@") + 1]
    pad = 2 * part - len(t2)
    t2 = t2 + b"@p\n" + b"A" * (pad - 3 - 1 - 3 - (pad - 8) // 2 - 1) + b"\n"  # not a valid record; only sizes matter here
    # round-robin over files
    files = [po.synth_fastq(G=20000, cov=30, seed=s) for s in (1, 2, 3)]
    order = [i for i, _ in po.chunks(files, part_size=1 << 19)]
    assert order[:6] == [0, 1, 2, 0, 1, 2]
    # keys are the same whichever way the stream is cut
    a = np.sort(np.concatenate([po.reads_to_keys(c, 31, 30) for _, c in po.chunks([files[0]], part_size=1 << 19)]))
    b = np.sort(po.reads_to_keys(files[0], 31, 30))
    assert (a == b).all()


def test_param_block_matches_survey_table(po):
    # SURVEY.md 8(a) / BASELINE.md section 2
    p = po.param_block(31, 4600000, 367999920, alpha=0.002, deNoise=0)
    assert (p.qb, p.hb, p.num_deNoise) == (26, 34, 0)
    p = po.param_block(31, 4600000, 367999920, alpha=0.002, deNoise=2)
    assert (p.qb, p.hb, p.num_deNoise) == (25, 33, 1)
    assert abs(p.trigger - 1.567e7) < 1e5
    p = po.param_block(47, 120000000, 16506371070, alpha=0.00234, deNoise=-1)
    assert (p.qb, p.num_deNoise) == (29, 8)
    p = po.param_block(47, 120000000, 6930000000, alpha=0.00234, deNoise=-1)
    assert (p.qb, p.num_deNoise) == (29, 3)
    p = po.param_block(31, 3000000000, 72000000000, alpha=0.002, deNoise=0)
    assert (p.qb, p.hb) == (34, 42)


def test_poisson_cdf_against_scipy(po):
    from scipy import stats
    O = po.oracle()
    for mean in (0.5, 3.0, 79.0, 123.0, 400.0):
        for x in (0, 1, 2, 5, 10, 50, 100, 200, 500):
            assert abs(O.orc_cdfpoi(x, mean) - stats.poisson.cdf(x, mean)) <= 1e-12 + 1e-9 * stats.poisson.cdf(x, mean)


def test_true2false_dp(po):
    errs = np.full(150, 0.002)
    r = po.oracle().orc_true2falseKmer_DP(errs.ctypes.data, 150, 31)
    p_true = (1 - 0.002) ** 31
    assert abs(r - p_true / (1 - p_true)) < 1e-9


def test_count_run_schedule(po):
    """State machine at t=1: rounds fire after the first chunk that reaches the trigger."""
    K, qb = 25, 21
    files = [po.synth_fastq(G=150000, cov=25, err=0.01, seed=31 + i, L=100) for i in range(2)]
    part = 1 << 20
    keys_per_chunk = [po.reads_to_keys(c, K, qb + 8) for _, c in po.chunks(files, part_size=part)]
    total = len(np.unique(np.concatenate(keys_per_chunk)))
    trigger = total // 3
    q, rounds = po.count_run(files, K, qb, qb + 8, trigger, 2, True, part_size=part)
    # replay by hand with the CQF model
    m = po.OracleCqf(qb, qb + 8)
    left, done = 2, 0
    for ks in keys_per_chunk:
        m.insert_keys(ks)
        if left and m.counts()[0] >= trigger:
            left -= 1
            m.denoise()
            done += 1
    if (left and m.counts()[0] >= trigger) or True:
        m.denoise()
        done += 1
    assert rounds == done
    assert q.counts() == m.counts()
    assert q.export() == m.export()


@pytest.mark.skipif(not os.path.exists("/root/reference/cqf/gqf.c"), reason="reference sources only exist in the build container")
def test_oracle_against_reference_core(po):
    rng_seed = 123
    for qb, n_base, n_single in [(9, 150, 120), (13, 2500, 2000), (16, 20000, 25000), (21, 400000, 900000)]:
        hb = qb + 8
        keys = gen.multiset(rng_seed + qb, hb, n_base, n_single)
        o, r = po.OracleCqf(qb, hb), po.RefCqf(qb, hb)
        assert o.insert_keys(keys) == r.insert_keys(keys)
        r.set_counts(*o.counts())
        assert o.export() == r.export()
        for _ in range(2):
            assert o.denoise() == r.denoise()
            r.set_counts(*o.counts())
            assert o.export() == r.export()
            assert r.check() == 0
        qk = np.concatenate([keys[:500], gen.splitmix64(9, 500) & np.uint64((1 << hb) - 1)])
        a, b = o.query_batch(qk), r.query_batch(qk)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        assert o.export() == r.export()
    # hashing: oracle key stream == reference NTPC64 stream on a read with N
    text = po.synth_fastq(G=3000, cov=3, n_rate=0.01, seed=4)
    R = po.ref()
    want = []
    for line in text.split(b"\n")[1::4]:
        s0 = 0
        while len(line) - s0 >= 31:
            fh, rh = C.c_uint64(), C.c_uint64()
            R.ref_ntpc64_base(line[s0:], 31, C.byref(fh), C.byref(rh))
            want.append(min(fh.value, rh.value) % (1 << 30))
            restart = False
            for i in range(s0 + 31, len(line)):
                if line[i:i + 1] == b"N":
                    s0, restart = i + 1, True
                    break
                R.ref_ntpc64_roll(line[i - 31], line[i], 31, C.byref(fh), C.byref(rh))
                want.append(min(fh.value, rh.value) % (1 << 30))
            if not restart:
                break
    assert po.reads_to_keys(text, 31, 30).tolist() == want
