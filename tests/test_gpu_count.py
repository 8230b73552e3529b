"""Parity of the CUDA counting path (through the C ABI) against the oracle.  -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fastq(po, **kw):
    return po.synth_fastq(**kw)


def test_hash_matches_oracle(po, shq):
    q = shq.CQF(20, 28)
    for K, kw in [(31, dict(G=20000, cov=5)), (47, dict(G=20000, cov=4, seed=3)), (21, dict(G=5000, cov=6, L=63, seed=5)),
                  (31, dict(G=20000, cov=4, n_rate=0.01, seed=9)), (5, dict(G=3000, cov=3, L=40, n_rate=0.05, seed=11)),
                  (200, dict(G=30000, cov=5, L=700, n_rate=0.002, seed=13))]:
        text = fastq(po, **kw)
        want = np.sort(po.reads_to_keys(text, K, 28))
        got = np.sort(q.hash_fastq_chunk(text, K))
        assert len(got) == len(want), (K, len(got), len(want))
        assert (got == want).all(), K
    q.close()


def test_hash_odd_bytes(po, shq):
    # lower case, IUPAC letters, raw low bytes, reads shorter than K, empty sequence lines
    recs = [b"ACGTNNNNacgtRYKMSWacgtacgtacgtacgtACGTACGTACGTACGTAAAACCCCGGGGTTTT", b"ACG", b"",
            b"N" * 40, b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTNACGTACGTACGTACGTACGTACGTACGTACGTACGTAN",
            bytes([1, 3, 4, 7]) * 12 + b"ACGTACGTAC", b"ACGT" * 100 + b"N" + b"TTGA" * 50,
            # odd bytes deep inside long reads: the 8-window fast step must hand over to the byte-exact step and back
            b"GATTACA" * 20 + b"R" + b"CATG" * 40 + b"n" + b"TTGCA" * 30, b"ACCGT" * 25 + bytes([3]) + b"GGTCA" * 31 + b"Y",
            b"acgtt" * 40 + b"NN" + b"ccgta" * 9 + b"N" + b"G" * 70, b"T" * 33 + b"-" + b"ACGGT" * 17 + b"N"]
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(recs))
    q = shq.CQF(16, 24)
    for K in (4, 11, 31, 32, 47):
        want = np.sort(po.reads_to_keys(text, K, 24))
        got = np.sort(q.hash_fastq_chunk(text, K))
        assert len(got) == len(want) and (got == want).all(), K
    q.close()


def test_malformed_chunk_is_an_error(po, shq):
    q = shq.CQF(16, 24)
    with pytest.raises(shq.ShqError) as e:
        q.insert_fastq_chunk(b"@r\nACGTACGTACGT\n+\n", 5)
        q.sync()
    assert e.value.status == 4
    q.close()


@pytest.mark.parametrize("qb,nk", [(10, 600), (14, 9000), (18, 150000), (22, 3000000)])
def test_insert_keys_image_matches_oracle(po, shq, qb, nk):
    hb = qb + 8
    rng = np.random.default_rng(qb)
    base = rng.integers(0, 1 << hb, size=nk // 3, dtype=np.uint64)
    keys = np.concatenate([base, base[: nk // 6], base[: nk // 12], rng.integers(0, 1 << hb, size=nk // 2, dtype=np.uint64),
                           np.repeat(base[:5], 300), np.repeat(base[5:6], 20000),
                           np.array([0, 0, 0, 255, 256, (1 << hb) - 1, (1 << hb) - 1], np.uint64)])
    rng.shuffle(keys)
    o = po.OracleCqf(qb, hb)
    o.insert_keys(keys)
    g = shq.CQF(qb, hb)
    half = len(keys) // 2
    g.insert_keys(keys[:half])
    g.insert_keys(keys[half:])
    assert g.counts() == o.counts()
    gk, gc = g.iterate()
    ok, oc = o.iterate()
    assert len(gk) == len(ok) and (gk == ok).all() and (gc == oc).all()
    assert g.export() == o.export()
    # deNoise twice (second is a no-op apart from the range-end quirk)
    for _ in range(2):
        assert g.denoise() == o.denoise()
        assert g.counts() == o.counts()
        assert g.export() == o.export()
    g.close()


def test_capacity_error(shq):
    g = shq.CQF(8, 16)
    keys = np.arange(0, 600, dtype=np.uint64) * 97 % (1 << 16)
    g.insert_keys(keys)
    with pytest.raises(shq.ShqError) as e:
        g.export()
    assert e.value.status == 3
    g.close()


def test_fastq_counting_matches_oracle(po, shq):
    K, qb = 31, 20
    text = fastq(po, G=60000, cov=30, err=0.004, n_rate=0.0005, seed=21)
    files = [text]
    o, rounds = po.count_run(files, K, qb, qb + 8, 0, 0, False)
    g = shq.CQF(qb, qb + 8)
    for _, ch in po.chunks(files):
        g.insert_fastq_chunk(ch, K)
    g.finish()
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    g.close()


@pytest.mark.parametrize("rounds,end", [(1, True), (3, False), (2, True)])
def test_denoise_schedule_matches_oracle(po, shq, rounds, end):
    """Mid-stream rounds pinned to reference chunk boundaries (t=1), several files round-robin."""
    K, qb = 25, 22
    files = [fastq(po, G=150000, cov=25, err=0.01, seed=31 + i, L=100) for i in range(3)]
    part = 1 << 20          # small parts so that the run has ~30 chunks
    total_distinct = len(np.unique(np.concatenate([po.reads_to_keys(f, K, qb + 8) for f in files])))
    trigger = total_distinct // (rounds + 1)
    o, o_rounds = po.count_run(files, K, qb, qb + 8, trigger, rounds, end, part_size=part)
    g = shq.CQF(qb, qb + 8)
    g.set_denoise_schedule(trigger, rounds, end)
    for _, ch in po.chunks(files, part_size=part):
        g.insert_fastq_chunk(ch, K)
    g.finish()
    assert g.rounds_done() == o_rounds
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    g.close()


def test_load_query_roundtrip(po, shq, tmp_path):
    qb, hb = 16, 24
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << hb, size=30000, dtype=np.uint64)
    keys = np.concatenate([keys, keys[:12000], np.repeat(keys[:3], 1000)])
    o = po.OracleCqf(qb, hb)
    o.insert_keys(keys)
    img = o.export()
    path = str(tmp_path / "t.cqf")
    open(path, "wb").write(img)
    g = shq.CQF.load(path)
    assert g.export() == img
    q = np.concatenate([keys[:3000], rng.integers(0, 1 << hb, size=3000, dtype=np.uint64)])
    q = np.unique(q)                      # distinct keys: duplicate order inside a batch is unspecified
    oc, ow = o.query_batch(q)
    gc, gw = g.query_set_traveled(q)
    assert (gc == oc).all() and (gw == ow).all()
    oc, ow = o.query_batch(q)             # second time: traveled bits now set
    gc, gw = g.query_set_traveled(q)
    assert (gc == oc).all() and (gw == ow).all()
    assert g.export() == o.export()       # traveled words included
    gk, gcnt = g.iterate()
    ok, ocnt = o.iterate()
    assert (gk == ok).all() and (gcnt == ocnt).all()
    g.save(path)
    assert open(path, "rb").read() == o.export()
    g.close()


@pytest.mark.parametrize("batch_mb", ["1", "4"])
def test_pinned_pipeline_matches_oracle(po, shq, batch_mb, monkeypatch):
    """Zero-copy pinned path: several small device batches, copies overlapped with compute, schedule armed."""
    import torch
    monkeypatch.setenv("SHQ_MAX_BATCH_MB", batch_mb)
    K, qb = 27, 22
    files = [fastq(po, G=120000, cov=30, err=0.008, seed=61 + i, L=120) for i in range(2)]
    part = 1 << 19
    total_distinct = len(np.unique(np.concatenate([po.reads_to_keys(f, K, qb + 8) for f in files])))
    trigger = total_distinct // 3
    o, o_rounds = po.count_run(files, K, qb, qb + 8, trigger, 2, True, part_size=part)
    g = shq.CQF(qb, qb + 8)
    g.set_denoise_schedule(trigger, 2, True)
    keep = []
    for _, ch in po.chunks(files, part_size=part):
        t = torch.frombuffer(bytearray(ch), dtype=torch.uint8).pin_memory()
        keep.append(t)
        g.insert_fastq_pinned_raw(t.data_ptr(), t.numel(), K)
    g.finish()
    assert g.rounds_done() == o_rounds
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    g.close()


@pytest.mark.parametrize("wide,path", [("2", "device"), ("2", "pinned"), ("1", "device")])
def test_wide_batches_match_oracle(po, shq, wide, path, monkeypatch):
    """Batches past the armed chunk cap (binned without chunk ids, merged speculatively): a wide batch that reaches the
    deNoise trigger is dropped and redone in halves.  SHQ_WIDE=2 forces every pending-round batch wide, so the halving runs
    down to single chunks around each of the three boundaries; SHQ_WIDE=1 is the production rule (by the new-key rate)."""
    import torch
    monkeypatch.setenv("SHQ_MAX_BATCH_MB", "6")
    monkeypatch.setenv("SHQ_ARMED_CHUNKS", "2")
    monkeypatch.setenv("SHQ_WIDE", wide)
    K, qb = 27, 22
    files = [fastq(po, G=150000, cov=30, err=0.008, seed=91 + i, L=120) for i in range(2)]
    part = 1 << 19
    total_distinct = len(np.unique(np.concatenate([po.reads_to_keys(f, K, qb + 8) for f in files])))
    trigger = total_distinct // 4
    o, o_rounds = po.count_run(files, K, qb, qb + 8, trigger, 3, False, part_size=part)
    g = shq.CQF(qb, qb + 8)
    g.set_denoise_schedule(trigger, 3, False)
    chunks = [ch for _, ch in po.chunks(files, part_size=part)]
    if path == "pinned":
        keep = []
        for ch in chunks:
            t = torch.frombuffer(bytearray(ch), dtype=torch.uint8).pin_memory()
            keep.append(t)
            g.insert_fastq_pinned_raw(t.data_ptr(), t.numel(), K)
    else:
        text = b"".join(chunks)
        ends = np.cumsum([len(ch) for ch in chunks]).astype(np.uint64)
        d = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
        d[:len(text)].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
        torch.cuda.synchronize()
        g.insert_fastq_device(d.data_ptr(), len(text), ends, K)
    g.finish()
    st = g.stats()
    assert g.rounds_done() == o_rounds == 3
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    assert st["wide_batches"] > 0
    if wide == "2":
        assert st["wide_fallbacks"] >= 3
    g.close()


def test_shard_handles_of_large_filters(shq):
    """configs[2] scaled to 4 and 8 ranks is a qb = 31 / 32 filter: the shard geometry must not go through 32-bit shifts."""
    for qb, n in [(31, 4), (32, 8), (34, 8)]:
        g = shq.CQF(qb, qb + 8, shard=n - 1, nshards=n)
        assert g.nblocks > (1 << (qb - 6))
        g.close()


def _mg_text(po):
    return fastq(po, G=80000, cov=30, err=0.004, n_rate=0.0003, seed=77)


def test_two_level_binned_insert_world1(po, shq):
    """The multi-GPU exchange path on one rank: coarse binning over the whole key space, owner-side split into fine
    bins, aggregate + merge; image and counters must equal the oracle's."""
    import importlib
    import torch
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    K, qb = 31, 20
    text = _mg_text(po)
    o, _ = po.count_run([text], K, qb, qb + 8, 0, 0, False)
    t = np.frombuffer(text, np.uint8)
    d = torch.empty(len(t) + 64, dtype=torch.uint8, device="cuda")
    d[:len(t)].copy_(torch.from_numpy(t.copy()))
    torch.cuda.synchronize()
    for cb, fb in [(8, 4), (11, 1), (6, 5)]:
        g = shq.CQF(qb, qb + 8)
        wp, op, nw = g.mg_bin_fastq_device(d.data_ptr(), len(t), K, cb)
        off = mg.dev_view(torch, op, (1 << cb) + 1, 0)
        counts = (off[1:] - off[:-1]).contiguous()
        torch.cuda.synchronize()
        g.mg_insert_binned(wp, np.zeros(1, np.uint64), counts.data_ptr(), 1, cb, fb, nw)
        assert g.counts() == o.counts(), (cb, fb)
        assert g.export() == o.export(), (cb, fb)
        g.close()


def test_pull_exchange_and_shard_image_world1(po, shq):
    """shq_mg_setup / open_peers / bin_local / insert_pull and the carry-based shard image with a single rank."""
    import torch
    K, qb = 31, 20
    text = _mg_text(po)
    o, _ = po.count_run([text], K, qb, qb + 8, 0, 0, False)
    t = np.frombuffer(text, np.uint8)
    d = torch.empty(len(t) + 64, dtype=torch.uint8, device="cuda")
    d[:len(t)].copy_(torch.from_numpy(t.copy()))
    torch.cuda.synchronize()
    g = shq.CQF(qb, qb + 8)
    handles = g.mg_setup(len(t))
    g.mg_open_peers(handles, 1)
    hist = torch.zeros(4096, dtype=torch.int32, device="cuda")
    half = (len(t) // 2)
    cut = text.rfind(b"\n@", 0, half) + 1          # a record boundary ('@' cannot start a quality line here: all 'I')
    for parity, (a, b) in enumerate([(0, cut), (cut, len(t))]):
        nw = g.mg_bin_local(d.data_ptr() + a, b - a, K, parity, hist.data_ptr())
        # published layout: begin / end of each of the 2048 global coarse bins inside the exchange buffer
        assert nw > 0 and int((hist[2048:4096] - hist[:2048]).sum().item()) == nw
        g.mg_insert_pull(hist.data_ptr(), parity)
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    assert g.stats()["bin_fallbacks"] == 0          # one hashing pass, fixed-capacity bins on both sides
    # shard image (world 1: no carry, no spill) == the blocks of the exported file
    agg = g.mg_layout_aggregate()
    NEG = -(1 << 60)
    keys = torch.zeros(16, dtype=torch.int64, device="cuda")
    nt = g.mg_layout_tail((0, NEG), 16, keys.data_ptr(), keys.data_ptr(), keys.data_ptr())
    assert nt == 0 and agg[0] > 0
    img, fb_, nb = g.mg_materialize_shard(0, 0, 0, 0)
    assert fb_ == 0 and nb == g.nblocks
    out = torch.empty(nb * 89, dtype=torch.uint8, device="cuda")
    g.copy_device(out.data_ptr(), img, nb * 89)
    assert g.header(*g.counts()) + out.cpu().numpy().tobytes() == o.export()
    g.close()


def test_pull_exchange_skewed_bins_fall_back(po, shq):
    """The exchange bins in one hashing pass into fixed-capacity bins on the source and on the owner; a read set dominated
    by one k-mer overflows both and each side redoes its step with exact offsets (counting pass) -- same image."""
    import torch
    K, qb = 31, 20
    base = fastq(po, G=30000, cov=10, err=0.003, seed=93)
    rep = b"".join(b"@p%d\n%s\n+\n%s\n" % (i, b"A" * 150, b"I" * 150) for i in range(6000))
    text = base + rep
    o, _ = po.count_run([text], K, qb, qb + 8, 0, 0, False)
    t = np.frombuffer(text, np.uint8)
    d = torch.empty(len(t) + 64, dtype=torch.uint8, device="cuda")
    d[:len(t)].copy_(torch.from_numpy(t.copy()))
    torch.cuda.synchronize()
    g = shq.CQF(qb, qb + 8)
    g.mg_open_peers(g.mg_setup(2 * len(t)), 1)
    hist = torch.zeros(4096, dtype=torch.int32, device="cuda")
    nw = g.mg_bin_local(d.data_ptr(), len(t), K, 0, hist.data_ptr())
    assert int((hist[2048:4096] - hist[:2048]).sum().item()) == nw
    g.mg_insert_pull(hist.data_ptr(), 0)
    assert g.stats()["bin_fallbacks"] >= 2
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    g.close()


def test_skewed_bins_fall_back_to_exact_binning(po, shq):
    """One-pass binning gives every bin the mean + 6 %; a read set dominated by one k-mer overflows its bin and the
    batch must be redone with exact offsets -- same image as the oracle either way."""
    K, qb = 31, 20
    base = fastq(po, G=30000, cov=10, err=0.003, seed=91)
    rep = b"".join(b"@p%d\n%s\n+\n%s\n" % (i, b"A" * 150, b"I" * 150) for i in range(6000))
    text = base + rep
    o, _ = po.count_run([text], K, qb, qb + 8, 0, 0, False)
    g = shq.CQF(qb, qb + 8)
    for _, ch in po.chunks([text]):
        g.insert_fastq_chunk(ch, K)
    g.finish()
    assert g.stats()["bin_fallbacks"] >= 1
    assert g.counts() == o.counts()
    assert g.export() == o.export()
    g.close()
