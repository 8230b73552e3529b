"""Generates tests/golden/golden.json from the UNMODIFIED reference core.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
It loads oracle/_ref/libshq_ref.so (gqf.c + nthash.hpp compiled from /root/reference by
oracle/Makefile) -- never the restatement -- so the fixtures pin the oracle to the reference.
"""
import base64
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402
from tests.golden import gen  # noqa: E402


def nthash_cases(R):
    out = []
    seqs = ["ACGTTGCATGCAAGCTTAGCCGATCGATTACGGATCCGTAACGTTAGCAAGGCTTAACC",
            "acgttgcatgcaagcttagccgatcgattacggatccgtaacgttagcaaggcttaacc",
            "ACGTNACGTACGTACGTACGTACGTACGTACNNACGTRYKMSWBDHVacgtnACGTAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA",
            gen.dna(7, 400).decode(), gen.dna(8, 300, b"ACGTNacgtnRY\x01\x03\x04\x07").decode("latin1")]
    for s in seqs:
        b = s.encode("latin1")
        for K in (4, 21, 31, 47, 64, 65, 100):
            if len(b) < K + 3:
                continue
            fh, rh = C.c_uint64(), C.c_uint64()
            R.ref_ntpc64_base(b, K, C.byref(fh), C.byref(rh))
            vals = [[fh.value, rh.value]]
            for i in range(K, min(len(b), K + 40)):
                R.ref_ntpc64_roll(b[i - K], b[i], K, C.byref(fh), C.byref(rh))
                vals.append([fh.value, rh.value])
            out.append({"seq_b64": base64.b64encode(b).decode(), "K": K, "fh_rh": vals})
    return out


def cqf_cases():
    cases = []
    # the hand case of SURVEY.md 8(c)
    r = po.RefCqf(6, 14)
    ins = [(0x310, 1), (0x310, 1), (0x305, 1), (0x300, 3), (0x480, 130), (0x3F0, 1), (0x501, 2)]
    for k, c in ins:
        if c == 1:
            r.insert_keys(np.array([k], np.uint64))
        else:
            r.insert_counted(np.array([k], np.uint64), np.array([c], np.uint64))
    r.set_counts(6, 139)
    img0 = r.export()
    k, c = r.iterate()
    q1 = r.query_batch(np.array([0x310, 0x310, 0x311, 0x480], np.uint64))
    img1 = r.export()
    # deNoise runs during counting, when every traveled bit is still 0: use a fresh filter
    r = po.RefCqf(6, 14)
    for k_, c_ in ins:
        r.insert_counted(np.array([k_], np.uint64), np.array([c_], np.uint64))
    removed = r.denoise(ideal=True)
    r.set_counts(6 - removed, 139 - removed)
    cases.append({"name": "survey_hand_case", "qb": 6, "hb": 14, "inserts": ins,
                  "image_b64": base64.b64encode(img0).decode(), "iter_keys": k.tolist(), "iter_counts": c.tolist(),
                  "query_keys": [0x310, 0x310, 0x311, 0x480], "query_counts": q1[0].tolist(), "query_was": q1[1].tolist(),
                  "image_after_query_b64": base64.b64encode(img1).decode(), "removed_ideal": int(removed),
                  "image_after_denoise_b64": base64.b64encode(r.export()).decode()})
    # seeded multisets: full image for the small ones, sha256 for the large ones
    for name, qb, n_base, n_single, seed in [("small_q8", 8, 60, 70, 1068), ("mid_q12", 12, 900, 1200, 1912),
                                             ("dense_q12", 12, 1500, 1400, 2512), ("big_q21", 21, 500000, 700000, 501021),
                                             ("quirk_q21", 21, 200000, 500000, 2011), ("quirk_q22", 22, 300000, 1000000, 2000)]:
        hb = qb + 8
        keys = gen.multiset(seed, hb, n_base, n_single)
        r = po.RefCqf(qb, hb)
        fresh = r.insert_keys(keys)
        r.set_counts(fresh, len(keys))
        img = r.export()
        case = {"name": name, "qb": qb, "hb": hb, "seed": seed, "n_base": n_base, "n_single": n_single,
                "n_keys": int(len(keys)), "ndistinct": int(fresh), "image_sha256": hashlib.sha256(img).hexdigest(), "check": r.check()}
        if qb <= 12:
            case["image_b64"] = base64.b64encode(img).decode()
        rounds = []
        nd, ne = fresh, len(keys)
        for _ in range(2):
            removed = r.denoise()
            nd -= removed
            ne -= removed
            r.set_counts(nd, ne)
            k, c = r.iterate()
            rounds.append({"removed": int(removed), "image_sha256": hashlib.sha256(r.export()).hexdigest(),
                           "surviving_singletons": k[c == 1].tolist()[:64], "check": r.check()})
        case["denoise_rounds"] = rounds
        # queries with traveled bits
        qk = np.concatenate([keys[:200], gen.splitmix64(5, 200) & np.uint64((1 << hb) - 1)])
        qc, qw = r.query_batch(qk)
        qc2, qw2 = r.query_batch(qk)
        case["query_sha256"] = hashlib.sha256(qc.tobytes() + qw.tobytes() + qc2.tobytes() + qw2.tobytes()).hexdigest()
        case["image_after_query_sha256"] = hashlib.sha256(r.export()).hexdigest()
        cases.append(case)
    return cases


def main():
    po.build(ref=True)
    R = po.ref()
    assert R is not None, "needs /root/reference"
    gold = {"generated_from": "oracle/_ref/libshq_ref.so (unmodified /root/reference/cqf/gqf.c + base/nthash.hpp)",
            "nthash": nthash_cases(R), "cqf": cqf_cases()}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote golden.json:", len(gold["nthash"]), "ntHash cases,", len(gold["cqf"]), "CQF cases")


if __name__ == "__main__":
    main()
