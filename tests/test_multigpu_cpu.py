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


def test_owner_rule():
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    hb = 30
    keys = np.array([0, (1 << 28) - 1, 1 << 28, (1 << 30) - 1], np.uint64)
    assert mg.owner_of(keys, hb, 4).tolist() == [0, 0, 1, 3]
    assert mg.owner_of(keys, hb, 1).tolist() == [0, 0, 0, 0]
