"""torchrun --nproc-per-node N tools/check_multigpu.py : sharded counting vs the oracle (bit-exact .cqf)."""
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
    mg = importlib.import_module("sh-assembly_b200.multigpu")
    K, qb = 31, 20 + (world.bit_length() - 1)      # the read set grows with the world size, so does the filter
    texts = [po.synth_fastq(G=100000, cov=30, err=0.004, n_rate=0.0003, seed=50 + r) for r in range(world)]
    text = np.frombuffer(texts[rank], np.uint8)
    d = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    d[:len(text)].copy_(torch.from_numpy(text.copy()))
    torch.cuda.synchronize()
    for end in (False, True):
        sc = mg.ShardedCounter(qb, qb + 8, rank, world, lr, dist)
        sc.insert_fastq_device(d.data_ptr(), len(text), K)
        removed = sc.end_denoise() if end else 0
        img = sc.export()
        img2 = sc.export_sharded()
        nd, ne = sc.global_counts()
        if rank == 0:
            o, _ = po.count_run(texts, K, qb, qb + 8, 0, 0, end)
            ok = img == o.export()
            ok2 = img2 == o.export()
            print("world", world, "endDeNoise", end, "removed", removed, "counts", (nd, ne), o.counts(), "image equal:", ok,
                  "sharded image equal:", ok2, flush=True)
            assert ok and ok2 and (nd, ne) == o.counts()
        if end:
            fa = sc.build_unitigs(texts if rank == 0 else None, K)
            if rank == 0:
                fa_o, _ = po.contiger_run(o, texts, K, 2, 2, 1000000)
                same = po.canonical_unitig_set(fa) == po.canonical_unitig_set(fa_o) and po.canonical_links(fa) == po.canonical_links(fa_o)
                print("world", world, "unitigs", len(po.parse_unitigs(fa)), "equal to the oracle's:", same, flush=True)
                assert same
        dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
