"""Multi-GPU layer: the filter sharded by quotient range, one process per GPU.

Rank r owns keys [r, r+1) * 2^hb / world.  Per batch every rank hashes its own reads (libshq kernels),
groups the keys by owner (libshq), exchanges them with ONE all-to-all over NCCL (NVLink/NVSwitch) and inserts
what it receives into its shard (libshq).  There is no other data-path collective.  Exact deNoise sweeps and
the .cqf export need the reference's global slot layout (range hand-out of 2^20 slots, spill across shard
boundaries), so they gather the sorted shard states onto every rank / rank 0 and run the single-GPU code there.

torch / torch.distributed are plumbing only: device buffers, the process group, the collective.
The index arithmetic (`split_plan`, `owner_of`, `slice_range`) is pure Python/NumPy so that the N>1 logic is
covered by world_size-2 gloo tests on CPU (tests/test_multigpu_cpu.py).
"""
import ctypes as C
import importlib
import json
import os

import numpy as np


def owner_of(keys: np.ndarray, hb: int, world: int) -> np.ndarray:
    """Shard that owns each key: the high log2(world) bits of the hb-bit key."""
    bits = int(world).bit_length() - 1
    return (keys.astype(np.uint64) >> np.uint64(hb - bits)).astype(np.int64) if bits else np.zeros(len(keys), np.int64)


def split_plan(send_counts, all_send_counts):
    """all-to-all plan.  send_counts[j] = #keys this rank sends to j; all_send_counts[i][j] = what rank i sends to j.
    Returns (recv_counts, send_offsets, recv_offsets) for this rank (rank index inferred from the matching row)."""
    send_counts = [int(x) for x in send_counts]
    rank = None
    for i, row in enumerate(all_send_counts):
        if [int(x) for x in row] == send_counts:
            rank = i if rank is None else rank
    raise_if = rank is None
    if raise_if:
        raise ValueError("send_counts is not a row of all_send_counts")
    return plan_for_rank(rank, all_send_counts)


def plan_for_rank(rank, all_send_counts):
    recv_counts = [int(all_send_counts[i][rank]) for i in range(len(all_send_counts))]
    send_counts = [int(x) for x in all_send_counts[rank]]
    send_offsets = np.concatenate([[0], np.cumsum(send_counts)[:-1]]).astype(np.int64).tolist()
    recv_offsets = np.concatenate([[0], np.cumsum(recv_counts)[:-1]]).astype(np.int64).tolist()
    return recv_counts, send_offsets, recv_offsets


def slice_range(sorted_keys: np.ndarray, lo: int, hi: int):
    """[begin, end) of the entries with lo <= key < hi in a sorted key array."""
    return int(np.searchsorted(sorted_keys, np.uint64(lo), "left")), int(np.searchsorted(sorted_keys, np.uint64(hi), "left"))


def exchange(dist, send: "torch.Tensor", send_counts, world):
    """One all-to-all of int64 keys grouped by destination.  Works on CUDA (nccl) and CPU (gloo) tensors."""
    import torch
    sc = torch.tensor([int(x) for x in send_counts], dtype=torch.int64, device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(recv_counts), dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv, send[:sum(int(x) for x in send_counts)], output_split_sizes=recv_counts,
                           input_split_sizes=[int(x) for x in send_counts])
    return recv, recv_counts


def gather_states(dist, keys: "torch.Tensor", counts: "torch.Tensor", world):
    """All ranks receive the concatenation (in rank = key order) of every shard's sorted (key,count) state."""
    import torch
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=keys.device)
    sizes = [torch.empty(1, dtype=torch.int64, device=keys.device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes + [1])
    kp = torch.zeros(mx, dtype=torch.int64, device=keys.device)
    cp = torch.zeros(mx, dtype=torch.int32, device=keys.device)
    kp[:keys.numel()] = keys
    cp[:counts.numel()] = counts
    ks = [torch.empty(mx, dtype=torch.int64, device=keys.device) for _ in range(world)]
    cs = [torch.empty(mx, dtype=torch.int32, device=keys.device) for _ in range(world)]
    dist.all_gather(ks, kp)
    dist.all_gather(cs, cp)
    allk = torch.cat([ks[i][:sizes[i]] for i in range(world)])
    allc = torch.cat([cs[i][:sizes[i]] for i in range(world)])
    return allk, allc, sizes


class ShardedCounter:
    """CQF-deNoise over `world` GPUs: mid-stream counting with --deNoise=0, optional exact end sweep, exact export."""

    def __init__(self, qb, hb, rank, world, device, dist):
        import torch
        self.torch = torch
        self.shq = importlib.import_module("sh-assembly_b200")
        self.qb, self.hb, self.rank, self.world, self.device, self.dist = qb, hb, rank, world, device, dist
        self.shard = self.shq.CQF(qb, hb, device=device, shard=rank, nshards=world)
        self.key_lo = (1 << hb) // world * rank
        self.key_hi = (1 << hb) // world * (rank + 1)
        self.nelts_local = 0          # occurrences inserted into this shard
        self.full = None              # lazily created unsharded handle for sweeps / export
        self._bufs = {}

    def _buf(self, name, n, dtype):
        t = self._bufs.get(name)
        if t is None or t.numel() < n:
            t = self.torch.empty(int(n * 1.1) + 1024, dtype=dtype, device=self.torch.device("cuda", self.device))
            self._bufs[name] = t
        return t

    def reset(self):
        self.shard.reset()
        self.nelts_local = 0

    def insert_fastq_device(self, d_text_ptr, n_bytes, K):
        """Hash this rank's reads, exchange keys by owner, insert the received keys into the local shard."""
        torch = self.torch
        cap = n_bytes // 2 + 1024
        keys = self._buf("keys", cap, torch.int64)
        nk = self.shard.hash_fastq_device(d_text_ptr, n_bytes, K, keys.data_ptr(), cap)
        grouped = self._buf("grouped", max(nk, 1), torch.int64)
        counts = self.shard.partition_keys_device(keys.data_ptr(), nk, self.world, grouped.data_ptr())
        self.shard.sync()
        torch.cuda.current_stream().synchronize()
        recv, recv_counts = exchange(self.dist, grouped, counts, self.world)
        torch.cuda.synchronize()      # the collective ran on torch's streams; libshq works on its own stream
        self.shard.insert_keys_device(recv.data_ptr(), recv.numel())
        self.nelts_local += recv.numel()
        return nk

    def _state_tensors(self):
        torch = self.torch
        kp, cp, n = self.shard.state_view()
        dev = torch.device("cuda", self.device)
        k = torch.empty(n, dtype=torch.int64, device=dev)
        c = torch.empty(n, dtype=torch.int32, device=dev)
        if n:
            self.shard.copy_device(k.data_ptr(), kp, n * 8)
            self.shard.copy_device(c.data_ptr(), cp, n * 4)
        return k, c

    def _full_handle(self):
        if self.full is None:
            self.full = self.shq.CQF(self.qb, self.hb, device=self.device)
        return self.full

    def global_counts(self):
        torch = self.torch
        nd, _ = self.shard.counts()
        t = torch.tensor([nd, self.nelts_local], dtype=torch.int64, device=torch.device("cuda", self.device))
        self.dist.all_reduce(t)
        return int(t[0].item()), int(t[1].item())

    def end_denoise(self):
        """One exact sweep (reference range hand-out incl. the range-end rule) on the gathered global state;
        every rank keeps the survivors of its own key range."""
        torch = self.torch
        k, c = self._state_tensors()
        allk, allc, sizes = gather_states(self.dist, k, c, self.world)
        nd, ne = self.global_counts()
        full = self._full_handle()
        torch.cuda.synchronize()
        full.adopt_state(allk.data_ptr(), allc.data_ptr(), allk.numel(), nd, ne)
        removed = full.denoise()
        kp, cp, n = full.state_view()
        dev = torch.device("cuda", self.device)
        gk = torch.empty(n, dtype=torch.int64, device=dev)
        if n:
            full.copy_device(gk.data_ptr(), kp, n * 8)
        lo = int(torch.searchsorted(gk, torch.tensor([self.key_lo], dtype=torch.int64, device=dev)).item())
        hi = int(torch.searchsorted(gk, torch.tensor([self.key_hi], dtype=torch.int64, device=dev)).item()) if self.key_hi < (1 << 63) else n
        # counts of the removed keys: each removed key had count 1
        before = k.numel()
        self.shard.adopt_state(kp + lo * 8, cp + lo * 4, hi - lo, hi - lo, 0)
        self.nelts_local -= before - (hi - lo)
        return removed

    def export(self):
        """Rank 0 returns the reference-format .cqf bytes of the whole filter; other ranks return None."""
        torch = self.torch
        k, c = self._state_tensors()
        allk, allc, sizes = gather_states(self.dist, k, c, self.world)
        nd, ne = self.global_counts()
        if self.rank != 0:
            return None
        full = self._full_handle()
        torch.cuda.synchronize()
        full.adopt_state(allk.data_ptr(), allc.data_ptr(), allk.numel(), nd, ne)
        return full.export()

    def materialize_global(self):
        """Build the global image on rank 0 in device memory (benchmark: 'table complete')."""
        torch = self.torch
        k, c = self._state_tensors()
        allk, allc, sizes = gather_states(self.dist, k, c, self.world)
        nd, ne = self.global_counts()
        if self.rank == 0:
            full = self._full_handle()
            torch.cuda.synchronize()
            full.adopt_state(allk.data_ptr(), allc.data_ptr(), allk.numel(), nd, ne)
            full.materialize()
        return nd, ne


def bench_sharded(args, rank, world, local_rank, text, ends, p, n_local, n_total, config, ClockSampler):
    """bench.py for N>1: weak scaling, genome and filter grow with N, --deNoise=0 mid-stream + the end sweep."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    d_text = torch.empty(len(text) + 64, dtype=torch.uint8, device=dev)
    d_text[:len(text)].copy_(torch.from_numpy(text))
    torch.cuda.synchronize()
    sc = ShardedCounter(int(p.qb), int(p.hb), rank, world, local_rank, dist)

    def step():
        sc.reset()
        sc.insert_fastq_device(d_text.data_ptr(), len(text), args.k)
        if args.end_denoise:
            sc.end_denoise()
        return sc.materialize_global()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        nd, ne = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    st = sc.shard.stats()
    launches = torch.tensor([st["kernels_launched"]], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)
    if rank == 0:
        ms_step = float(ms.item()) / args.steps
        value = n_total / (ms_step / 1e3)
        cfg = dict(config)
        cfg["workload"] += " | N>1: --deNoise=0 mid-stream (sharded mid-stream rounds are not built yet), exact end sweep and image on the gathered state"
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        a_ins = args.read_len / (args.read_len - args.k + 1) + 32
        line = {"metric": "canonical k-mers inserted/sec", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": None, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "end-to-end through host buffers is measured at N=1 only"},
                "gpu_launches": int(launches.item()),
                "roofline": {"bound": "hbm", "kernel": "whole sharded insert path", "achieved": value * a_ins / 1e9 / world, "peak": peak, "unit": "GB/s",
                             "frac": value * a_ins / 1e9 / world / peak, "traffic": None, "note": "per GPU, against SURVEY 8(d) A_ins"},
                "result": {"ndistinct": int(nd), "nelts": int(ne), "kmers_per_step": int(n_total)}}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
