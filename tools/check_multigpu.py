"""torchrun --nproc-per-node N tools/check_multigpu.py : the sharded filter on N real GPUs against the t = 1 oracle over N
read files -- counters, deNoise sweeps, the sharded image (bit-exact .cqf) and the unitigs -- with a deNoise schedule
whose boundaries fall inside batches."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from oracle import pyoracle as po

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    shq = importlib.import_module("sh-assembly_b200")
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    os.environ.setdefault("SHQ_MG_ROUNDS", "6")
    K, qb, part = 31, 20 + (world.bit_length() - 1), 1 << 18      # the read set grows with the world size, so does the filter
    texts = [po.synth_fastq(G=100000, cov=30, err=0.004, n_rate=0.0003, seed=50 + r) for r in range(world)]
    text = np.frombuffer(texts[rank], np.uint8)
    ends, off = [], 0
    for _, ch in shq.Reader(buffers=[text], part_size=part):
        off += len(ch); ends.append(off)
    d = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    d[:len(text)].copy_(torch.from_numpy(text.copy()))
    torch.cuda.synchronize()
    for trigger, rounds, end in ((0, 0, False), (130000 * world, 2, True), (110000 * world, 4, False)):
        sc = mg.ShardedCounter(qb, qb + 8, rank, world, lr, dist)
        nd, ne = sc.count(d.data_ptr(), len(text), K, ends, trigger=trigger, rounds=rounds, end_denoise=end)
        img = sc.export_sharded()
        if rank == 0:
            o, orounds = po.count_run(texts, K, qb, qb + 8, trigger, rounds, end, part_size=part)
            ok = img == o.export()
            print("world", world, "schedule", (trigger, rounds, end), "sweeps", sc.sweeps, orounds, "counts", (nd, ne), o.counts(),
                  "sharded image equal:", ok, "search", (sc.search.attempts, sc.search.rejected), flush=True)
            assert ok and (nd, ne) == o.counts() and sc.sweeps == orounds
        if end:
            chunks = [c for _, c in shq.Reader(buffers=texts, part_size=part)] if rank == 0 else None
            fa = sc.build_unitigs(chunks, K)
            if rank == 0:
                fa_o, _ = po.contiger_run(o, texts, K, 2, 2, 1000000, part_size=part)
                same = po.canonical_unitig_set(fa) == po.canonical_unitig_set(fa_o) and po.canonical_links(fa) == po.canonical_links(fa_o)
                print("world", world, "unitigs", len(po.parse_unitigs(fa)), "equal to the oracle's:", same, flush=True)
                assert same
        dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
