"""CPU suite, world_size 2 over gloo: the N>1 host logic of sh-assembly_b200/multigpu.py
(owner rule, all-to-all plan and exchange, state gather, key-range slicing) against the oracle."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    from oracle import pyoracle as po
    K, qb = 21, 16
    hb = qb + 8
    text = po.synth_fastq(G=8000, cov=10, L=80, seed=100 + rank)       # every rank hashes its own reads
    keys = po.reads_to_keys(text, K, hb)
    own = mg.owner_of(keys, hb, world)
    order = np.argsort(own, kind="stable")
    grouped = torch.from_numpy(keys[order].astype(np.int64))
    counts = np.bincount(own, minlength=world).tolist()
    recv, recv_counts = mg.exchange(dist, grouped, counts, world)
    recv = recv.numpy().astype(np.uint64)
    lo, hi = (1 << hb) // world * rank, (1 << hb) // world * (rank + 1)
    assert ((recv >= lo) & (recv < hi)).all()
    # local shard as a sorted (key,count) state; gather; compare with the oracle over all reads
    uk, uc = np.unique(recv, return_counts=True)
    allk, allc, sizes = mg.gather_states(dist, torch.from_numpy(uk.astype(np.int64)), torch.from_numpy(uc.astype(np.int32)), world)
    texts = [po.synth_fastq(G=8000, cov=10, L=80, seed=100 + r) for r in range(world)]
    o = po.OracleCqf(qb, hb)
    for t in texts:
        o.insert_keys(po.reads_to_keys(t, K, hb))
    ok, oc = o.iterate()
    assert (allk.numpy().astype(np.uint64) == ok).all() and (allc.numpy().astype(np.uint64) == oc).all()
    b, e = mg.slice_range(allk.numpy().astype(np.uint64), lo, hi)
    assert e - b == len(uk) and sum(sizes[:rank]) == b
    # the plan arithmetic
    all_counts = [None] * world
    dist.all_gather_object(all_counts, counts)
    rc, so, ro = mg.plan_for_rank(rank, all_counts)
    assert rc == recv_counts and so[0] == 0 and ro[0] == 0
    # two-level exchange: coarse bins over the whole key space, 32-bit words, one slice per owner
    cb, cps, fb = mg.exchange_geometry(world, coarse_bits=6)
    lbc = hb - cb
    bins = (keys >> np.uint64(lbc)).astype(np.int64)
    order = np.argsort(bins, kind="stable")
    words = (keys[order] & np.uint64((1 << lbc) - 1)).astype(np.int32)
    cnt = torch.from_numpy(np.bincount(bins, minlength=1 << cb).astype(np.int32))
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    allc = np.stack([t.numpy() for t in allc])
    sc, rcnt, seg_start, my_counts = mg.binned_plan(allc, rank, world)
    assert sum(sc) == len(words) and seg_start[0] == 0
    rbuf = torch.empty(sum(rcnt), dtype=torch.int32)
    dist.all_to_all_single(rbuf, torch.from_numpy(words), output_split_sizes=rcnt, input_split_sizes=sc)
    got = []
    rb = rbuf.numpy().astype(np.uint64)
    for s in range(world):
        o = int(seg_start[s])
        for c in range(cps):
            n = int(my_counts[s][c])
            got.append(((np.uint64(rank * cps + c) << np.uint64(lbc)) | rb[o:o + n]))
            o += n
    got = np.sort(np.concatenate(got))
    assert (got == np.sort(recv)).all()          # same multiset as the 64-bit key exchange above
    ret[rank] = len(recv)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_exchange_and_gather():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and sum(ret.values()) > 0


def test_binned_plan_arithmetic():
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    assert mg.exchange_geometry(2) == (11, 1024, 2) and mg.exchange_geometry(8) == (11, 256, 4)
    rng = np.random.default_rng(3)
    world, nb = 4, 32
    allc = rng.integers(0, 50, size=(world, nb))
    tot_recv = 0
    for r in range(world):
        sc, rc, ss, mc = mg.binned_plan(allc, r, world)
        assert sc == [int(allc[r, d * 8:(d + 1) * 8].sum()) for d in range(world)]
        assert rc == [int(allc[s, r * 8:(r + 1) * 8].sum()) for s in range(world)]
        assert (mc == allc[:, r * 8:(r + 1) * 8]).all() and ss.tolist() == np.concatenate([[0], np.cumsum(rc)[:-1]]).tolist()
        tot_recv += sum(rc)
    assert tot_recv == allc.sum()


def test_maxplus_carry_matches_sequential_layout():
    """The per-shard composed maps + carry_before reproduce the sequential rule start = max(q, previous end + 1)."""
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    rng = np.random.default_rng(11)
    nslots, world = 1 << 12, 4
    q = np.sort(rng.integers(0, nslots, size=3000))
    ln = rng.integers(1, 4, size=len(q))
    free, first_free_at = 0, {}
    aggs = [(0, mg.NEG_INF)] * world
    prev_q = -1
    for qi, li in zip(q.tolist(), ln.tolist()):
        shard = qi // (nslots // world)
        first_free_at.setdefault(shard, free)               # first free slot when the shard's first entry arrives
        m = (li, qi + li if qi != prev_q else mg.NEG_INF)
        aggs[shard] = mg.maxplus_compose(aggs[shard], m)
        free = max(free + li, qi + li) if qi != prev_q else free + li
        prev_q = qi
    for r in range(world):
        c = mg.carry_before(aggs, r)
        assert max(c[0], c[1], 0) == first_free_at[r] or (r == 0 and first_free_at[r] == 0)
    tot = mg.carry_before(aggs, world)
    assert max(tot[0], tot[1]) == free


def test_owner_rule():
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    hb = 30
    keys = np.array([0, (1 << 28) - 1, 1 << 28, (1 << 30) - 1], np.uint64)
    assert mg.owner_of(keys, hb, 4).tolist() == [0, 0, 1, 3]
    assert mg.owner_of(keys, hb, 1).tolist() == [0, 0, 0, 0]


def test_local_chunk_range_partitions_the_round_robin():
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    world, n_local = 4, [5, 5, 4, 3]
    for c0, c1 in [(0, 20), (3, 9), (7, 8), (0, 1), (13, 40), (6, 6)]:
        got = []
        for r in range(world):
            lo, hi = mg.local_chunk_range(c0, c1, r, world, n_local[r])
            got += [j * world + r for j in range(lo, hi)]
        want = [c for c in range(c0, c1) if (c // world) < n_local[c % world]]
        assert sorted(got) == want


def test_trigger_search_reproduces_the_t1_schedule():
    """TriggerSearch over a fake key stream: rounds fire after exactly the chunks the sequential t = 1 rule names."""
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    rng = np.random.default_rng(5)
    for trial in range(6):
        nchunks = int(rng.integers(20, 120))
        chunks = [rng.integers(0, 3000 + 400 * c, size=int(rng.integers(50, 400))) for c in range(nchunks)]   # keys drift: new keys keep coming
        trigger, rounds = int(rng.integers(800, 3000)), int(rng.integers(1, 4))
        end = bool(trial % 2)
        # sequential reference: count -> fire after the first chunk that leaves ndistinct >= trigger; a sweep drops count-1 keys
        state, left, fired_ref = {}, rounds, []
        for c, ks in enumerate(chunks):
            for k in ks.tolist():
                state[k] = state.get(k, 0) + 1
            if left and len(state) >= trigger:
                left -= 1
                state = {k: v for k, v in state.items() if v != 1}
                fired_ref.append(c)
        if (left and len(state) >= trigger) or end:
            state = {k: v for k, v in state.items() if v != 1}
            fired_ref.append("eof")
        # the search: staged intervals against a committed state
        st = {"state": {}, "pending": None, "fired": [], "last": -1}

        def stage(c0, c1):
            add = {}
            for c in range(c0, c1):
                for k in chunks[c].tolist():
                    add[k] = add.get(k, 0) + 1
            st["pending"] = (add, c1 - 1)
            return sum(1 for k in add if k not in st["state"]), sum(add.values())

        def commit():
            add, last = st["pending"]
            for k, v in add.items():
                st["state"][k] = st["state"].get(k, 0) + v
            st["last"] = last
            st["pending"] = None

        def discard():
            st["pending"] = None

        def fire():
            before = len(st["state"])
            st["state"] = {k: v for k, v in st["state"].items() if v != 1}
            st["fired"].append(st["last"])
            return before - len(st["state"])

        ts = mg.TriggerSearch(trigger, rounds, stage, commit, discard, fire)
        for b0 in range(0, nchunks, 32):
            ts.feed(b0, min(nchunks, b0 + 32))
        n_mid = len(st["fired"])
        ts.finish(end)
        fired = st["fired"][:n_mid] + (["eof"] if len(st["fired"]) > n_mid else [])
        assert fired == fired_ref, (trial, fired, fired_ref)
        assert st["state"] == state and ts.ndistinct == len(state) and ts.nelts == sum(state.values())


def test_denoise_walk_finalize_closes_the_open_range():
    """shq_mg_denoise_finalize (host only): the range that is open when the entries run out ends on the very last entry,
    which is protected iff it is a one-slot cluster with count 1 sitting exactly on the range end."""
    import ctypes as C
    shq = importlib.import_module("sh-assembly_b200")
    L = shq.lib()

    def run(phase, e0, cur_end, pos, end, count, gap, owner=3):
        st = np.zeros(16, np.uint64)
        st[0], st[2], st[4] = phase, e0, cur_end
        st[5], st[6], st[7], st[8], st[9], st[10] = 1, pos, end, count, gap, owner
        assert L.shq_mg_denoise_finalize(st.ctypes.data) == 0
        return int(st[11])

    assert run(1, 1000, 0, 999, 999, 1, 1) == 1 << 3          # range ends at e0 - 1 = 999: one-slot cluster there, count 1
    assert run(1, 1000, 0, 999, 999, 2, 1) == 0               # count 2 is never a candidate
    assert run(1, 1000, 0, 999, 999, 1, 0) == 0               # adjacent to its predecessor: part of a longer cluster
    assert run(1, 1000, 0, 998, 998, 1, 1) == 0               # not on the range end
    assert run(1, 1000, 0, 1000, 1000, 1, 1) == 1 << 3        # covers slot e0: its cluster (itself) ends the range
    assert run(2, 1000, 1203, 1203, 1203, 1, 1) == 1 << 3     # cluster walk reached the end of all entries
    assert run(2, 1000, 1203, 1200, 1203, 1, 0) == 0
    assert run(0, 0, 0, 5, 5, 1, 1) == 0                      # no range open
