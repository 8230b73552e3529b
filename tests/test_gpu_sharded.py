"""The sharded filter with the deNoise schedule, bit-exact against the t = 1 oracle over N read files.

Two processes share the one GPU of the test box (gloo for the control plane; the k-mer exchange is the same pull over
CUDA-IPC mapped buffers that runs over NVLink between real GPUs): staged exchange batches, the search for the deNoise
boundary chunk, the sharded sweep with its rank-to-rank range walk, the sharded image with carry and boundary stitch."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, case, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if "rounds_env" in case:
        os.environ["SHQ_MG_ROUNDS"] = str(case["rounds_env"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from oracle import pyoracle as po
    shq = importlib.import_module("sh-assembly_b200")
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    K, qb, part = case["K"], case["qb"], case["part"]
    texts = [po.synth_fastq(G=case["G"], cov=case["cov"], err=case["err"], n_rate=0.0003, seed=case["seed"] + r, L=case.get("L", 150)) for r in range(world)]
    if case.get("short_last"):
        texts[-1] = texts[-1][:len(texts[-1]) // 3]
        texts[-1] = texts[-1][:texts[-1].rfind(b"\n@") + 1]
    text = np.frombuffer(texts[rank], np.uint8)
    ends, off = [], 0
    for _, ch in shq.Reader(buffers=[text], part_size=part):
        off += len(ch)
        ends.append(off)
    d = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    d[:len(text)].copy_(torch.from_numpy(text.copy()))
    torch.cuda.synchronize()
    sc = mg.ShardedCounter(qb, qb + 8, rank, world, 0, dist)
    assert sc.mode == "pull"
    nd, ne = sc.count(d.data_ptr(), len(text), K, ends, trigger=case["trigger"], rounds=case["rounds"], end_denoise=case["end"])
    img = sc.export_sharded()
    fa = None
    if case.get("unitigs"):
        # Contiger over the sharded image: rank 0 reads the other shard's blocks through the CUDA-IPC mapping
        all_chunks = [c for _, c in shq.Reader(buffers=texts, part_size=part)] if rank == 0 else None
        fa = sc.build_unitigs(all_chunks, K)
    if rank == 0:
        o, rounds = po.count_run(texts, K, qb, qb + 8, case["trigger"], case["rounds"], case["end"], part_size=part)
        ret["counts"] = ((nd, ne), o.counts())
        ret["rounds"] = (sc.sweeps, rounds)
        ret["image_equal"] = img == o.export()
        ret["search"] = (sc.search.attempts, sc.search.rejected)
        if fa is not None:
            fo, _ = po.contiger_run(o, texts, K, part_size=part)
            ret["unitigs_equal"] = po.canonical_unitig_set(fa) == po.canonical_unitig_set(fo) and po.canonical_links(fa) == po.canonical_links(fo)
            ret["n_unitigs"] = len(po.parse_unitigs(fa))
    dist.barrier()
    dist.destroy_process_group()


CASES = {
    # trigger low enough for two mid-stream rounds, plus the end round; small batches so that boundaries fall inside batches
    "two_rounds_end": dict(K=31, qb=21, part=1 << 18, G=60000, cov=40, err=0.004, seed=50, trigger=90000, rounds=2, end=True, rounds_env=6, unitigs=True),
    # --deNoise=0: one batch after the other, no sweep
    "plain": dict(K=27, qb=20, part=1 << 18, G=50000, cov=30, err=0.003, seed=60, trigger=0, rounds=0, end=False),
    # ranks with different numbers of chunks, K=47, rounds left over at EOF
    "ragged": dict(K=47, qb=21, part=1 << 17, G=40000, cov=30, err=0.004, seed=70, trigger=70000, rounds=5, end=False, rounds_env=5, short_last=True),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_two_shards_match_the_oracle(name):
    import torch.multiprocessing as mp
    world = 2
    port = 29700 + (os.getpid() % 1500)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, CASES[name], ret), nprocs=world, join=True)
    assert ret["counts"][0] == ret["counts"][1], ret["counts"]
    assert ret["rounds"][0] == ret["rounds"][1], ret["rounds"]
    assert ret["image_equal"]
    if CASES[name]["rounds"]:
        assert ret["search"][0] > 0
    if CASES[name].get("unitigs"):
        assert ret["unitigs_equal"] and ret["n_unitigs"] > 10
