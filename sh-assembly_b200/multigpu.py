"""Multi-GPU layer: the filter sharded by quotient range, one process per GPU.

Rank r owns keys [r, r+1) * 2^hb / world.  Per batch every rank hashes its own reads (libshq kernels),
groups the keys by owner (libshq), exchanges them with ONE all-to-all over NCCL (NVLink/NVSwitch) and inserts
what it receives into its shard (libshq).  There is no other data-path collective.  Exact deNoise sweeps and
the .cqf export need the reference's global slot layout (range hand-out of 2^20 slots, spill across shard
boundaries), so they gather the sorted shard states onto every rank / rank 0 and run the single-GPU code there.

torch / torch.distributed are plumbing only: device buffers, the process group, the collective.
The index arithmetic (`plan_for_rank`, `owner_of`, `slice_range`) is pure Python/NumPy so that the N>1 logic is
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


def plan_for_rank(rank, all_send_counts):
    recv_counts = [int(all_send_counts[i][rank]) for i in range(len(all_send_counts))]
    send_counts = [int(x) for x in all_send_counts[rank]]
    send_offsets = np.concatenate([[0], np.cumsum(send_counts)[:-1]]).astype(np.int64).tolist()
    recv_offsets = np.concatenate([[0], np.cumsum(recv_counts)[:-1]]).astype(np.int64).tolist()
    return recv_counts, send_offsets, recv_offsets


def slice_range(sorted_keys: np.ndarray, lo: int, hi: int):
    """[begin, end) of the entries with lo <= key < hi in a sorted key array."""
    return int(np.searchsorted(sorted_keys, np.uint64(lo), "left")), int(np.searchsorted(sorted_keys, np.uint64(hi), "left"))


NEG_INF = -(1 << 60)


def maxplus_compose(x, y):
    """f(e) = max(e + a, b) as the pair (a, b); apply x first, then y (csrc/shq_scan.cuh MaxPlusOp)."""
    return (x[0] + y[0], max(x[1] + y[0], y[1]))


def carry_before(aggs, rank):
    """Composed slot-layout map of all shards before `rank` (identity for rank 0)."""
    c = (0, NEG_INF)
    for r in range(rank):
        c = maxplus_compose(c, aggs[r])
    return c


def local_chunk_range(c0, c1, rank, world, n_local):
    """The global chunk sequence of `world` read files is the reference's round robin (seqFile_batch::getDataChunk,
    cqf/CQF_mt.h:364-390): global chunk c = round * world + rank.  Returns the local chunks [lo, hi) of `rank` that fall
    into the global interval [c0, c1), clipped to the n_local chunks this rank has."""
    lo = max(0, -(-(c0 - rank) // world))
    hi = max(0, -(-(c1 - rank) // world))
    return min(lo, n_local), min(hi, n_local)


class TriggerSearch:
    """Locates the deNoise boundary of the t = 1 rule (a round fires after the first chunk that leaves
    ndistinct >= trigger, cqf/CQF_mt.h:837) WITHOUT chunk ids in the exchanged words: chunk intervals are staged
    speculatively -- stage(c0, c1) returns the number of new keys of the interval and leaves it uncommitted --, an
    interval that stays below the trigger is committed, one that crosses it is dropped and retried shorter.  The new-key
    rate per chunk seen so far sizes the next attempt just short of the expected crossing, so the search ends after a
    handful of short attempts.  Every decision uses global (all-reduced) numbers only: all ranks run the same sequence.

    stage/commit/discard/fire are callables; fire() runs one sweep and returns the number of removed keys."""

    def __init__(self, trigger, rounds, stage, commit, discard, fire):
        self.trigger, self.rounds_left = int(trigger), int(rounds)
        self.stage, self.commit, self.discard, self.fire = stage, commit, discard, fire
        self.ndistinct = 0
        self.nelts = 0
        self.rate = None          # new keys per global chunk, from the last staged interval
        self.attempts = 0
        self.rejected = 0

    def _fire(self):
        removed = self.fire()
        self.ndistinct -= removed
        self.nelts -= removed
        self.rounds_left -= 1

    def feed(self, c0, c1):
        """global chunks [c0, c1), in order"""
        while c0 < c1:
            if self.rounds_left <= 0:
                new, occ = self.stage(c0, c1)
                self.commit()
                self.ndistinct += new
                self.nelts += occ
                return
            e = c1
            if self.ndistinct >= self.trigger:
                e = c0 + 1                     # still over the trigger after a sweep: the next chunk fires again
            elif self.rate is not None and self.rate > 0:
                reach = (self.trigger - self.ndistinct) / self.rate
                if reach < (c1 - c0):
                    e = c0 + max(1, int(reach * 0.92))
            new, occ = self.stage(c0, e)
            self.attempts += 1
            if self.ndistinct + new < self.trigger or e - c0 == 1:
                self.commit()
                self.ndistinct += new
                self.nelts += occ
                if e - c0 > 1 or self.rate is None:
                    self.rate = new / (e - c0)
                if self.ndistinct >= self.trigger:
                    self._fire()
                c0 = e
            else:
                self.discard()
                self.rejected += 1
                self.rate = new / (e - c0)

    def finish(self, end_denoise):
        """EOF rule of fastq_to_uint64kmers_prod (cqf/CQF_mt.h:858-866)"""
        if (self.rounds_left > 0 and self.ndistinct >= self.trigger) or end_denoise:
            removed = self.fire()
            self.ndistinct -= removed
            self.nelts -= removed
            if self.rounds_left > 0:
                self.rounds_left -= 1


def exchange_geometry(world, coarse_bits=11, fine_total_bits=12):
    """(coarse_bits, coarse bins per shard, fine_bits): 2^coarse_bits bins over the whole key space at the source,
    every coarse bin split into 2^fine_bits at the owner so that a shard ends up with 2^fine_total_bits bins."""
    shard_bits = int(world).bit_length() - 1
    cb = max(coarse_bits, shard_bits + 1)
    cps = (1 << cb) // world
    fb = max(0, fine_total_bits - (cb - shard_bits))
    return cb, cps, fb


def binned_plan(all_counts: np.ndarray, rank: int, world: int):
    """all_counts[s][b] = words source s holds in global coarse bin b (owner of b = b // cps).
    Returns (send_counts, recv_counts, seg_start, my_counts): words this rank sends to / receives from every rank,
    the start of every source's slice in the receive buffer, and my_counts[s][c] = words of MY coarse bin c from s."""
    all_counts = np.asarray(all_counts, np.int64)
    nb = all_counts.shape[1]
    cps = nb // world
    per_owner = all_counts.reshape(world, world, cps)          # [source][owner][c]
    send_counts = per_owner[rank].sum(axis=1)
    recv_counts = per_owner[:, rank, :].sum(axis=1)
    seg_start = np.concatenate([[0], np.cumsum(recv_counts)[:-1]]).astype(np.uint64)
    my_counts = np.ascontiguousarray(per_owner[:, rank, :]).astype(np.uint32)
    return [int(x) for x in send_counts], [int(x) for x in recv_counts], seg_start, my_counts


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


class _DevView:
    """CUDA array interface over a raw device pointer owned by libshq (no copy, no ownership)."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def dev_view(torch, ptr, n, device):
    """int32 tensor aliasing n words at device address ptr."""
    return torch.as_tensor(_DevView(ptr, n), device=torch.device("cuda", device))


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
        # k-mer exchange: "pull" = owners read their bins out of the sources' HBM over NVLink inside the split kernel
        # (CUDA IPC, one node); "nccl" = one all-to-all of the binned 32-bit words
        self.mode = os.environ.get("SHQ_MG_EXCHANGE", "pull")
        self.parity = 0
        self.max_batch_bytes = 2 << 30
        if self.mode == "pull":
            # every rank must end up in the same mode: if mapping a peer's buffer fails anywhere (no CUDA IPC / no peer
            # access in this environment), all ranks switch to the all-to-all exchange -- same results, more traffic
            ok = 1
            mine = b"\0" * 128
            try:
                mine = self.shard.mg_setup(self.max_batch_bytes // 2 + (1 << 20))
            except self.shq.ShqError:
                ok = 0
            handles = [None] * world
            dist.all_gather_object(handles, mine)
            if ok:
                try:
                    self.shard.mg_open_peers(b"".join(handles), world)
                except self.shq.ShqError:
                    ok = 0
            flag = torch.tensor([ok], dtype=torch.int64, device=torch.device("cuda", device))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.mode = "nccl"

    def _buf(self, name, n, dtype):
        t = self._bufs.get(name)
        if t is None or t.numel() < n:
            t = self.torch.empty(int(n * 1.1) + 1024, dtype=dtype, device=self.torch.device("cuda", self.device))
            self._bufs[name] = t
        return t

    def reset(self):
        self.shard.reset()
        self.nelts_local = 0

    def insert_fastq_device(self, d_text_ptr, n_bytes, K, chunk_ends=None, max_batch_bytes=None):
        """Device-resident text of this rank, cut into batches at chunk boundaries (every rank runs the same number
        of exchange rounds; a rank that has run out of text takes part with empty batches)."""
        torch = self.torch
        cuts = self._cuts(n_bytes, chunk_ends, max_batch_bytes)
        nb = torch.tensor([len(cuts)], dtype=torch.int64, device=torch.device("cuda", self.device))
        self.dist.all_reduce(nb, op=self.dist.ReduceOp.MAX)
        total = 0
        for i in range(int(nb.item())):
            a, b = cuts[i] if i < len(cuts) else (0, 0)
            total += (self._pull_batch if self.mode == "pull" else self._exchange_batch)(d_text_ptr + a, b - a, K)
        return total

    def _cuts(self, n_bytes, chunk_ends, max_batch_bytes=None):
        """Batches of whole reference chunks, at most max_batch_bytes each: [(begin, end), ...]."""
        max_batch_bytes = max_batch_bytes or self.max_batch_bytes
        ends = [int(e) for e in chunk_ends] if chunk_ends is not None and len(chunk_ends) else [int(n_bytes)]
        cuts, start, prev = [], 0, 0
        for e in ends:
            if e - start > max_batch_bytes and prev > start:
                cuts.append((start, prev)); start = prev
            prev = e
        if prev > start:
            cuts.append((start, prev))
        return cuts

    def insert_fastq_pinned(self, h_text, d_buf, K, chunk_ends=None):
        """End-to-end variant: h_text is this rank's FASTQ text in PINNED host memory, d_buf a device buffer of the same
        size.  All batches are copied on a side stream, one after the other; batch i is exchanged and inserted as soon
        as its copy has landed, while the later ones are still crossing PCIe."""
        torch = self.torch
        n = h_text.numel()
        cuts = self._cuts(n, chunk_ends)
        cs = getattr(self, "_copy_stream", None)
        if cs is None:
            cs = self._copy_stream = torch.cuda.Stream(device=torch.device("cuda", self.device))
        cs.wait_stream(torch.cuda.current_stream())
        events = []
        with torch.cuda.stream(cs):
            for a, b in cuts:
                d_buf[a:b].copy_(h_text[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                events.append(ev)
        nb = torch.tensor([len(cuts)], dtype=torch.int64, device=torch.device("cuda", self.device))
        self.dist.all_reduce(nb, op=self.dist.ReduceOp.MAX)
        total = 0
        for i in range(int(nb.item())):
            a, b = (0, 0)
            if i < len(cuts):
                events[i].synchronize()
                a, b = cuts[i]
            total += (self._pull_batch if self.mode == "pull" else self._exchange_batch)(d_buf.data_ptr() + a, b - a, K)
        return total

    def _tick(self, name, t0):
        import time
        if os.environ.get("SHQ_MG_TIMING"):
            self.torch.cuda.synchronize()
            T = self._timing = getattr(self, "_timing", {})
            T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    def _pull_batch(self, d_text_ptr, n_bytes, K):
        """Fused exchange: bin locally into an IPC-mapped buffer, all-gather the (tiny) fine-bin histograms, then every
        owner's split kernel loads its coarse bins straight from the peers' buffers."""
        import time
        torch = self.torch
        dev = torch.device("cuda", self.device)
        nfg = self.world * 4096
        t0 = time.perf_counter()
        hist = self._buf("fine_hist", nfg, torch.int32)[:nfg]
        nw = self.shard.mg_bin_local(d_text_ptr, n_bytes, K, self.parity, hist.data_ptr())
        t0 = self._tick("bin_local", t0)
        allh = self._buf("all_hist", self.world * nfg, torch.int32)[:self.world * nfg]
        self.dist.all_gather_into_tensor(allh, hist)
        torch.cuda.ExternalStream(self.shard.stream(), device=dev).wait_stream(torch.cuda.current_stream())
        t0 = self._tick("all_gather_hist", t0)
        self.shard.mg_insert_pull(allh.data_ptr(), self.parity)
        t0 = self._tick("insert_pull", t0)
        self.parity ^= 1
        return nw

    def _exchange_batch(self, d_text_ptr, n_bytes, K):
        """Two-level exchange: bin this rank's k-mers over the whole key space (coarse bins; 32-bit words), ship every
        owner its slice with ONE all-to-all, split the received coarse bins into fine bins and insert them."""
        torch = self.torch
        world, rank = self.world, self.rank
        dev = torch.device("cuda", self.device)
        import time
        T = self._timing = getattr(self, "_timing", {})
        def tick(name, t0):
            if os.environ.get("SHQ_MG_TIMING"):
                torch.cuda.synchronize()
                T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return time.perf_counter()
        t0 = time.perf_counter()
        cb, cps, fb = exchange_geometry(world)
        nb = 1 << cb
        wp, op, nw = self.shard.mg_bin_fastq_device(d_text_ptr, n_bytes, K, cb)
        t0 = tick("bin", t0)
        if nw:
            off = dev_view(torch, op, nb + 1, self.device)       # libshq's offsets, valid until its next call
            counts = (off[1:] - off[:-1]).contiguous()
        else:
            counts = torch.zeros(nb, dtype=torch.int32, device=dev)
        allc = self._buf("allc", world * nb, torch.int32)[:world * nb]
        self.dist.all_gather_into_tensor(allc, counts)
        send_counts, recv_counts, seg_start, my_counts = binned_plan(allc.view(world, nb).cpu().numpy(), rank, world)
        t0 = tick("plan", t0)
        send = dev_view(torch, wp, nw, self.device) if nw else torch.zeros(0, dtype=torch.int32, device=dev)
        total = sum(recv_counts)
        recv = self._buf("recv", max(total, 1), torch.int32)[:total]
        self.dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts)
        mc = torch.from_numpy(my_counts.view(np.int32)).to(dev)
        # the collective ran on torch's streams; libshq works on its own stream, which waits for it on the device
        torch.cuda.ExternalStream(self.shard.stream(), device=dev).wait_stream(torch.cuda.current_stream())
        t0 = tick("all_to_all", t0)
        self.shard.mg_insert_binned(recv.data_ptr(), seg_start, mc.data_ptr(), world, cb, fb, total)
        t0 = tick("insert", t0)
        return nw


    # ---------------------------------------------------------------- the same job at every N
    def _ctl_allreduce(self, vals):
        """sum of a few integers over the ranks (control plane)"""
        torch = self.torch
        if self.dist.get_backend() == "nccl":
            t = torch.tensor([int(v) for v in vals], dtype=torch.int64, device=torch.device("cuda", self.device))
            self.dist.all_reduce(t)
            return [int(x) for x in t.tolist()]
        t = torch.tensor([int(v) for v in vals], dtype=torch.int64)
        self.dist.all_reduce(t)
        return [int(x) for x in t.tolist()]

    def _ctl_allgather_hist(self, hist, allh):
        if self.dist.get_backend() == "nccl":
            self.dist.all_gather_into_tensor(allh, hist)
        else:   # gloo (two processes on one GPU in the tests): staged through the host
            parts = [self.torch.empty(hist.numel(), dtype=hist.dtype) for _ in range(self.world)]
            self.dist.all_gather(parts, hist.cpu())
            allh.copy_(self.torch.cat(parts))

    def _tick(self, key, t0):
        """host wall clock per call site (every one of them ends with a device or collective round trip)"""
        import time
        t1 = time.perf_counter()
        W = self.wall = getattr(self, "wall", {})
        W[key] = W.get(key, 0.0) + (t1 - t0) * 1e3
        return t1

    def _stage_interval(self, c0, c1):
        """stage global chunks [c0, c1): every rank bins its local share, owners pull, aggregate and count new keys"""
        import time
        t = time.perf_counter()
        torch = self.torch
        dev = torch.device("cuda", self.device)
        lo, hi = local_chunk_range(c0, c1, self.rank, self.world, len(self._ends))
        a = int(self._ends[lo - 1]) if lo > 0 else 0
        b = int(self._ends[hi - 1]) if hi > lo else a
        nfg = self.world * 4096
        hist = self._buf("fine_hist", nfg, torch.int32)[:nfg]
        evs = getattr(self, "h2d_events", None)
        if evs:      # end-to-end runs: the text of this interval must have landed
            for off, ev in evs:
                ev.synchronize()
                if off >= b:
                    break
        # a rank-local failure (malformed text, exchange buffer too small ...) must not leave the peers waiting in a
        # collective: the status rides in the last word of the layout buffer (the layout uses the first 4096) and in the
        # count all-reduce, and every rank raises together
        err = None
        try:
            self.shard.mg_bin_local(self._text_ptr + a, b - a, self._K, self.parity, hist.data_ptr())
        except self.shq.ShqError as e:
            err = e
        if self.world > 1:
            hist[nfg - 1] = 1 if err else 0
        t = self._tick("bin_local", t)
        allh = self._buf("all_hist", self.world * nfg, torch.int32)[:self.world * nfg]
        self._ctl_allgather_hist(hist, allh)
        torch.cuda.ExternalStream(self.shard.stream(), device=dev).wait_stream(torch.cuda.current_stream())
        failed = int(allh.view(self.world, nfg)[:, nfg - 1].sum().item()) if self.world > 1 else (1 if err else 0)
        t = self._tick("allgather_layouts(wait for peers)", t)
        if failed:
            raise err if err else RuntimeError("rank %d: a peer failed while binning chunks [%d, %d)" % (self.rank, c0, c1))
        nn = no = 0
        try:
            nn, no = self.shard.mg_stage_pull(allh.data_ptr(), self.parity)
        except self.shq.ShqError as e:
            err = e
        t = self._tick("stage_pull", t)
        self.parity ^= 1
        tot = self._ctl_allreduce([nn, no, 1 if err else 0])
        self._tick("allreduce_counts(wait for peers)", t)
        if tot[2]:
            raise err if err else RuntimeError("rank %d: a peer failed while inserting chunks [%d, %d)" % (self.rank, c0, c1))
        return tot[0], tot[1]

    def _commit(self):
        import time
        t = time.perf_counter()
        self.shard.mg_commit()
        self._tick("commit", t)

    def _discard(self):
        self.shard.mg_discard()

    def denoise(self):
        """One deNoise sweep over the sharded filter (exact: the reference's range hand-out incl. the range-end rule).
        Layout carry (max-plus) -> the range walk handed from rank to rank -> local compaction.  Returns the number of
        keys removed over all shards."""
        import time
        t = time.perf_counter()
        torch = self.torch
        a, b = self.shard.mg_layout_aggregate()
        aggs = self._ctl_allgather_pairs((a, b))
        carry = carry_before(aggs, self.rank)
        self.shard.mg_denoise_positions(carry)      # every rank lays its shard out now; the chain below is serial
        state = np.zeros(16, np.uint64)
        for r in range(self.world):
            if self.rank == r:
                self.shard.mg_denoise_walk(carry, state)
            state = self._ctl_bcast_state(state, r)
        self.shq.lib().shq_mg_denoise_finalize(state.ctypes.data)
        protect = (int(state[11]) >> self.rank) & 1
        removed = self.shard.mg_denoise_apply(protect)
        r = self._ctl_allreduce([removed])[0]
        self._tick("denoise_sweep", t)
        return r

    def _allgather_dev(self, allb, buf):
        """all_gather_into_tensor for device tensors on any backend (gloo: staged through the host)"""
        if self.dist.get_backend() == "nccl":
            self.dist.all_gather_into_tensor(allb, buf)
        else:
            parts = [self.torch.empty(buf.numel(), dtype=buf.dtype) for _ in range(self.world)]
            self.dist.all_gather(parts, buf.cpu())
            allb.copy_(self.torch.cat(parts))

    def _ctl_allgather_pairs(self, pair):
        torch = self.torch
        if self.dist.get_backend() == "nccl":
            dev = torch.device("cuda", self.device)
            mine = torch.tensor([int(pair[0]), int(pair[1])], dtype=torch.int64, device=dev)
            allg = torch.empty(2 * self.world, dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(allg, mine)
            v = allg.view(self.world, 2).tolist()
        else:
            parts = [torch.empty(2, dtype=torch.int64) for _ in range(self.world)]
            self.dist.all_gather(parts, torch.tensor([int(pair[0]), int(pair[1])], dtype=torch.int64))
            v = [p.tolist() for p in parts]
        return [(int(x[0]), int(x[1])) for x in v]

    def _ctl_bcast_state(self, state, src):
        torch = self.torch
        t = torch.from_numpy(state.view(np.int64).copy())
        if self.dist.get_backend() == "nccl":
            t = t.to(torch.device("cuda", self.device))
            self.dist.broadcast(t, src)
            t = t.cpu()
        else:
            self.dist.broadcast(t, src)
        return t.numpy().view(np.uint64).copy()

    def count(self, d_text_ptr, n_bytes, K, chunk_ends, trigger=0, rounds=0, end_denoise=False):
        """CQF_mt::build_KmerSpectrum over `world` read files, one per rank (this rank's text at d_text_ptr, cut into
        the reference's chunks by chunk_ends), with the t = 1 deNoise schedule on the global chunk order.
        Returns (global ndistinct, global nelts, sweeps done)."""
        self._text_ptr, self._K = int(d_text_ptr), int(K)
        self._ends = [int(e) for e in chunk_ends] if chunk_ends is not None and len(chunk_ends) else ([int(n_bytes)] if n_bytes else [])
        R = self._ctl_allreduce_max(len(self._ends))
        ts = TriggerSearch(trigger, rounds, self._stage_interval, self._commit, self._discard, self.denoise)
        self.search = ts
        j = 0
        while j < R:
            # rounds per batch: 32 (256 MiB of text per rank) while the shards are small, up to 256 (2 GiB) as they grow --
            # every batch pays one merge over the whole shard (same rule as batch_target_bytes on one GPU)
            per_shard = ts.ndistinct / self.world
            grow = 1.0 if per_shard < 2.5e7 else min(8.0, per_shard / 8.0e6)
            rb = int(32 * grow) if not os.environ.get("SHQ_MG_ROUNDS") else int(os.environ["SHQ_MG_ROUNDS"])
            j1 = min(R, j + rb)
            ts.feed(j * self.world, j1 * self.world)
            j = j1
        ts.finish(end_denoise)
        self.sweeps = self.shard.rounds_done()
        return ts.ndistinct, ts.nelts

    def _ctl_allreduce_max(self, v):
        torch = self.torch
        if self.dist.get_backend() == "nccl":
            t = torch.tensor([int(v)], dtype=torch.int64, device=torch.device("cuda", self.device))
        else:
            t = torch.tensor([int(v)], dtype=torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())

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
        nd, ne = self.shard.counts()
        t = self._ctl_allreduce([nd, ne])
        return t[0], t[1]

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
        _, ne_local = self.shard.counts()
        self.shard.adopt_state(kp + lo * 8, cp + lo * 4, hi - lo, hi - lo, ne_local - (before - (hi - lo)))
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

    SPILL_CAP = 1 << 14

    def materialize_sharded(self):
        """Every rank builds the blocks of its own quotient range in its HBM (no gather): the slot layout is a max-plus
        scan, so a shard needs the composed map of its predecessors (2 x int64, all-gathered) and the handful of
        entries of the previous shard whose slots reach across the boundary (boundary stitch).
        Returns (device pointer of the shard image, first block, number of blocks)."""
        import time
        t_mat = time.perf_counter()
        torch = self.torch
        dev = torch.device("cuda", self.device)
        a, b = self.shard.mg_layout_aggregate()
        aggs = self._ctl_allgather_pairs((a, b))
        carry = carry_before(aggs, self.rank)
        cap = self.SPILL_CAP
        # one buffer per rank: [n_tail | keys | pos | counts] as int64 words
        buf = self._buf("spill", 1 + 3 * cap, torch.int64)[:1 + 3 * cap]
        kp, pp, cp = buf[1:1 + cap], buf[1 + cap:1 + 2 * cap], buf[1 + 2 * cap:1 + 3 * cap]
        nt = self.shard.mg_layout_tail(carry, cap, kp.data_ptr(), cp.data_ptr(), pp.data_ptr())
        buf[0] = nt
        allb = self._buf("spill_all", self.world * (1 + 3 * cap), torch.int64)[:self.world * (1 + 3 * cap)]
        self._allgather_dev(allb, buf)
        n_spill = 0
        sk = sc = sp = 0
        if self.rank > 0:
            prev = allb.view(self.world, 1 + 3 * cap)[self.rank - 1]
            n_spill = int(prev[0].item())
            sk, sp, sc = prev[1:1 + cap].data_ptr(), prev[1 + cap:1 + 2 * cap].data_ptr(), prev[1 + 2 * cap:1 + 3 * cap].data_ptr()
        torch.cuda.synchronize()
        r = self.shard.mg_materialize_shard(sk, sc, sp, n_spill)
        self._tick("materialize_shard", t_mat)
        return r

    def export_sharded(self):
        """Rank 0 returns the reference-format .cqf bytes assembled from the shard images; other ranks return None."""
        torch = self.torch
        img, fb, nb = self.materialize_sharded()
        t = torch.empty(nb * 89, dtype=torch.uint8, device=torch.device("cuda", self.device))
        self.shard.copy_device(t.data_ptr(), img, nb * 89)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (fb, t.cpu().numpy().tobytes()))
        nd, ne = self.global_counts()
        if self.rank != 0:
            return None
        parts.sort(key=lambda x: x[0])
        return self.shard.header(nd, ne) + b"".join(p[1] for p in parts)

    def build_unitigs(self, chunks, K, s_min=2, x_min=2, x_max=1000000, master_rule=False):
        """Contiger over the SHARDED filter: every rank keeps the blocks of its own quotient range in its HBM
        (materialize_sharded); rank 0 maps the peers' block arrays through CUDA IPC and runs the device path with its
        neighbour queries (and traveled-bit atomics) going straight to the owners' memory -- no rank ever holds the whole
        image.  The node table and the chain contraction still live on rank 0 (owner-partitioned nodes with a stitch of chain
        ends are not built).  Rank 0 returns the FASTA text, other ranks None.  `chunks` = the read chunks of ALL ranks, on rank 0."""
        self.materialize_sharded()
        handles = [None] * self.world
        self.dist.all_gather_object(handles, self.shard.mg_image_ipc())
        nd, _ = self.global_counts()
        fa = None
        if self.rank == 0:
            self.shard.mg_open_image_peers(b"".join(handles), self.world)
            try:
                fa = self.shq.build_unitigs(self.shard, chunks, K, s_min, x_min, x_max, master_rule=master_rule, global_ndistinct=nd)
            finally:
                self.shard.mg_close_image_peers()
        self.dist.barrier()       # the peers' images stay alive until rank 0 is done with them
        return fa

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


def bench_sharded(args, rank, world, local_rank, B):
    """bench.py for N>1 (B = the bench module): weak scaling, genome and filter grow with N; every rank generates and
    hashes its slice of the global read set and owns 1/N of the quotient space."""
    import torch
    import torch.distributed as dist
    import synth
    from synth import workloads as wl
    shq = importlib.import_module("sh-assembly_b200")
    dev = torch.device("cuda", local_rank)
    name = args.workload
    w = wl.scaled(name, world)
    sp = synth.make(w["G"], w["cov"], w["L"], w["err"], wl.SEED)
    n_total = wl.n_kmers(w)
    p = shq.param_block(w["K"], w["n_true"], n_total, alpha=w["err"], deNoise=w["deNoise"])
    first, nloc = synth.rank_slice(sp, rank, world)
    n_bytes = nloc * sp.rec_bytes
    d_text = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)
    synth.gpu_fill(sp, first, nloc, d_text.data_ptr(), torch.cuda.current_stream().cuda_stream)
    try:
        h_text = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
        host_note = "page-locked host text"
    except RuntimeError:      # the node cannot page-lock world x this much: pageable host memory (slower copies, same bytes)
        h_text = torch.empty(n_bytes, dtype=torch.uint8)
        host_note = "pageable host text (page-locking %d bytes per rank failed)" % n_bytes
    h_text.copy_(d_text[:n_bytes])
    torch.cuda.synchronize()
    text = h_text.numpy()
    rd = shq.Reader(buffers=[text])
    ends, off = [], 0
    while True:
        r = rd.next_raw()
        if r is None:
            break
        off += r[1]
        ends.append(off)
    rd.close()
    ends = np.array(ends, np.uint64)
    config = B.config_dict(w, world, p, name)
    ClockSampler = B.ClockSampler
    args.k = w["K"]
    sc = ShardedCounter(int(p.qb), int(p.hb), rank, world, local_rank, dist)

    trigger, rounds, end = int(p.trigger), int(p.num_deNoise), bool(w["end"])

    def step():
        sc.reset()
        nd_, ne_ = sc.count(d_text.data_ptr(), n_bytes, w["K"], ends, trigger=trigger, rounds=rounds, end_denoise=end)
        sc.materialize_sharded()
        return nd_, ne_

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sc._timing = {}
    sc.wall = {}
    sc.shard.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
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
    wall = {k: v / args.steps for k, v in sc.wall.items()}
    launches = torch.tensor([st["kernels_launched"]], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)
    search = (sc.search.attempts, sc.search.rejected)
    sweeps = sc.sweeps

    # end to end: pinned host text -> H2D in 64 MiB slices on a side stream (a batch is binned as soon as its slices have
    # landed) -> exchange + insert -> shard image -> D2H
    import time
    _, _, nb_blocks = sc.materialize_sharded()
    h_img = torch.empty(nb_blocks * 89, dtype=torch.uint8).pin_memory()
    cs = torch.cuda.Stream(device=dev)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def step_e2e():
        sc.reset()
        cs.wait_stream(torch.cuda.current_stream())
        evs = []
        with torch.cuda.stream(cs):
            for a0 in range(0, n_bytes, 64 << 20):
                b0 = min(n_bytes, a0 + (64 << 20))
                d_text[a0:b0].copy_(h_text[a0:b0], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                evs.append((b0, ev))
        sc.h2d_events = evs
        sc.count(d_text.data_ptr(), n_bytes, w["K"], ends, trigger=trigger, rounds=rounds, end_denoise=end)
        sc.h2d_events = None
        img, _, nbk = sc.materialize_sharded()
        h_img[:nbk * 89].copy_(torch.as_tensor(_DevView(img, nbk * 89, "|u1"), device=dev))
        torch.cuda.synchronize()

    step_e2e()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    dist.barrier()
    ms_e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    io = torch.tensor([n_bytes, nb_blocks * 89], dtype=torch.int64, device=dev)
    dist.all_reduce(io)
    # parity: the sharded image against the digest of the unmodified reference core run over the same N read files
    gold = B.golden_digest(name, world)
    parity = {"image_sha256": None, "expected": gold["image_sha256"] if gold else None, "match": False,
              "checker": gold["made_by"] if gold else "no committed digest for %s@%d (tests/golden/digests.json)" % (name, world)}
    if gold:
        import hashlib
        parts = [None] * world
        dist.all_gather_object(parts, (sc.rank, bytes(h_img[:nb_blocks * 89].numpy())))
        if rank == 0:
            parts.sort(key=lambda x: x[0])
            hsh = hashlib.sha256(sc.shard.header(int(nd), int(ne)))
            for _, bts in parts:
                hsh.update(bts)
            parity["image_sha256"] = hsh.hexdigest()
            parity["match"] = parity["image_sha256"] == gold["image_sha256"] and int(nd) == gold["ndistinct"] and int(ne) == gold["nelts"] and sweeps == gold["denoise_rounds"]
    if rank == 0 and os.environ.get("SHQ_MG_TIMING"):
        import sys
        print("mg timing (ms, warmup+timed steps):", {k: round(v, 2) for k, v in sc._timing.items()}, {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")}, file=sys.stderr)
    if rank == 0:
        ms_step = float(ms.item()) / args.steps
        value = n_total / (ms_step / 1e3)
        cfg = dict(config)
        cfg["workload"] += " | filter sharded by quotient range, pull exchange over peer memory, every rank materialises its own blocks (max-plus carry + boundary stitch)"
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        a_ins = w["L"] / (w["L"] - w["K"] + 1) + 32
        line = {"metric": "canonical k-mers inserted/sec", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": n_total / (float(ms_e2e.item()) / 1e3), "unit": "k-mers/s", "h2d_bytes_per_step": int(io[0].item()),
                        "d2h_bytes_per_step": int(io[1].item()), "ms_per_step": float(ms_e2e.item()),
                        "steps": e2e_steps,
                        "note": "every rank: " + host_note + " -> H2D in 64 MiB slices on a side stream -> exchange + insert as the slices land -> its blocks of the .cqf image -> D2H; wall clock between barriers, max over ranks; bytes summed over ranks"},
                "gpu_launches": int(launches.item()),
                "roofline": {"bound": "hbm", "kernel": "whole sharded insert path", "achieved": value * a_ins / 1e9 / world, "peak": peak, "unit": "GB/s",
                             "frac": value * a_ins / 1e9 / world / peak, "traffic": None, "note": "per GPU, against SURVEY 8(d) A_ins",
                             "phase_ms_per_step_rank0": {k: st[k] / args.steps for k in st if k.startswith("ms_")},
                             "host_syncs_per_step_rank0": st["host_syncs"] / args.steps, "batches_per_step": st["batches"] / args.steps,
                             "host_wall_ms_per_step_rank0": wall},
                "result": {"ndistinct": int(nd), "nelts": int(ne), "denoise_rounds": int(sweeps), "kmers_per_step": int(n_total),
                           "boundary_search": {"staged_intervals": search[0], "rejected": search[1]}},
                "parity": parity}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
