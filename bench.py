#!/usr/bin/env python
"""bench.py -- canonical k-mers inserted per second into the CQF (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (libshq.so, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload at N=1 = BASELINE.json configs[1]: synthetic E. coli-size genome (4.6 Mb), 100x 150 bp
reads, 0.2 % substitutions, K=31, CQF-deNoise ... --deNoise=2 --endDeNoise (effective: 1 mid-stream
round + the end round, qb=25).  A step = one whole counting job over that read set.
For N>1 the genome grows with N (N x 4.6 Mb, 100x) and the filter is sharded by quotient range:
weak scaling, every rank hashes 1/N of the reads and owns 1/N of the quotient space.

`value`  : device-resident FASTQ text -> complete CQF image in HBM (CUDA events on the handle's stream).
`e2e`    : same job through the C ABI with HOST buffers: chunks cut by the reader, pinned staging,
           H2D copies and the D2H export of the .cqf image inside the timed region.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_INS_BYTES = None  # filled from the workload: L/(L-K+1) + 32 + 2*T/N  


# ----------------------------------------------------------------------------- data
def gen_fastq(G, cov, L, err, seed, rank=0, world=1):
    """SURVEY.md 8(d) synthetic reads: iid genome, uniform positions, strand flip p=.5, substitutions.
    Rank r of `world` produces every world-th block of reads of the same global read set."""
    rng_g = np.random.default_rng(seed)
    genome = np.frombuffer(b"ACGT", np.uint8)[rng_g.integers(0, 4, G, dtype=np.uint8)]
    nreads = int(G * cov / L)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    code = np.zeros(256, np.uint8)
    for i, a in enumerate(b"ACGT"):
        code[a] = i
    acgt = np.frombuffer(b"ACGT", np.uint8)
    hdr_w = len(str(nreads))
    rec = 2 + hdr_w + 1 + L + 1 + 2 + L + 1
    block = 200000
    parts = []
    for bi, b0 in enumerate(range(0, nreads, block)):
        if bi % world != rank:
            continue
        n = min(block, nreads - b0)
        rng = np.random.default_rng([seed, 1, bi])
        pos = rng.integers(0, G - L + 1, n)
        reads = genome[pos[:, None] + np.arange(L)[None, :]]
        flip = rng.random(n) < 0.5
        reads[flip] = comp[reads[flip][:, ::-1]]
        if err > 0:
            mask = rng.random(reads.shape) < err
            reads[mask] = acgt[(code[reads[mask]] + rng.integers(1, 4, int(mask.sum()), dtype=np.uint8)) % 4]
        out = np.empty((n, rec), np.uint8)
        out[:, 0] = ord("@")
        out[:, 1] = ord("r")
        ids = np.arange(b0, b0 + n)
        for d in range(hdr_w):
            out[:, 2 + hdr_w - 1 - d] = ord("0") + (ids // 10 ** d) % 10
        o = 2 + hdr_w
        out[:, o] = 10
        out[:, o + 1:o + 1 + L] = reads
        out[:, o + 1 + L] = 10
        out[:, o + 2 + L] = ord("+")
        out[:, o + 3 + L] = 10
        out[:, o + 4 + L:o + 4 + 2 * L] = ord("I")
        out[:, o + 4 + 2 * L] = 10
        parts.append(out.reshape(-1))
    text = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    nloc = len(text) // rec
    return text, nloc * (L - 0) , nloc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, world):
    G = args.genome * world
    nreads = int(G * args.cov / args.read_len)
    n_total = nreads * (args.read_len - args.k + 1)
    return G, nreads, n_total


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU code (gqf.c + nthash.hpp, unmodified) driven by the std::thread restatement of its
    worker loop (oracle/ref_shim.cpp), all host threads, on a time-bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    shq = importlib.import_module("sh-assembly_b200")
    G, nreads, n_total = workload(args, world)
    p = shq.param_block(args.k, G, n_total, alpha=args.err, deNoise=args.denoise)
    text, _, nloc = gen_fastq(G, args.cov, args.read_len, args.err, args.seed, 0, 1)
    rd = shq.Reader(buffers=[text])
    ends, off = [], 0
    while True:
        r = rd.next_raw()
        if r is None:
            break
        off += r[1]
        ends.append(off)
    rd.close()
    cores = os.cpu_count() or 1
    threads = cores
    if po.ref() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libshq_ref.so missing (built only where /root/reference exists)"}))
        return
    budget = args.cpu_seconds
    times, rates, sample = [], [], ""
    for it in range(args.warmup + args.steps):
        r = po.RefCqf(int(p.qb), int(p.hb))
        secs, kmers, chunks, rounds = r.count_mt(text.ctypes.data, ends, args.k, threads, trigger=int(p.trigger), rounds=int(p.num_deNoise),
                                                 end_denoise=args.end_denoise, max_seconds=budget)
        r.close()
        if it >= args.warmup:
            times.append(secs)
            rates.append(kmers / secs)
        sample = "first %d of %d chunks (%.3g of %.3g k-mers), %.1f s budget per step, table sized for the full set (qb=%d)" % (
            chunks, len(ends), kmers, n_total, budget, p.qb)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": "canonical k-mers inserted/sec", "value": v, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_dict(args, world, p),
            "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(args, world, p):
    return {"workload": "BASELINE configs[1]: synthetic %.3g Mb genome, %dx %d bp reads, %.2g%% subst, k=%d, --deNoise=%d%s (effective %d mid-stream round(s)), qb=%d hb=%d"
            % (args.genome * world / 1e6, args.cov, args.read_len, 100 * args.err, args.k, args.denoise, " --endDeNoise" if args.end_denoise else "",
               p.num_deNoise, p.qb, p.hb),
            "genome_bp": args.genome * world, "coverage": args.cov, "read_len": args.read_len, "error_rate": args.err, "K": args.k,
            "deNoise": args.denoise, "endDeNoise": bool(args.end_denoise), "qb": int(p.qb), "trigger": int(p.trigger),
            "l2_policy": "inputs (>= 1 GB of FASTQ text per step) are larger than the 126 MB L2; no explicit flush",
            "shards": world}


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    shq = importlib.import_module("sh-assembly_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libshq has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    G, nreads, n_total = workload(args, world)
    p = shq.param_block(args.k, G, n_total, alpha=args.err, deNoise=args.denoise)
    text, _, nloc = gen_fastq(G, args.cov, args.read_len, args.err, args.seed, rank, world)
    n_local = nloc * (args.read_len - args.k + 1)
    # reference chunking (host reader, 8 MiB parts)
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
    if world > 1:
        from importlib import import_module
        mg = import_module("sh-assembly_b200.multigpu")
        return mg.bench_sharded(args, rank, world, local_rank, text, ends, p, n_local, n_total, config_dict(args, world, p), ClockSampler)

    d_text = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    h_text = torch.from_numpy(text).pin_memory()
    d_text[:len(text)].copy_(h_text)
    torch.cuda.synchronize()
    q = shq.CQF(int(p.qb), int(p.hb), device=local_rank)
    stream = torch.cuda.ExternalStream(q.stream(), device=torch.device("cuda", local_rank))

    def step_device():
        q.reset()
        q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), args.end_denoise)
        q.insert_fastq_device(d_text.data_ptr(), len(text), ends, args.k)
        q.finish()
        q.materialize()

    # destination of the exported .cqf image: page-locked, so the D2H copy runs at PCIe speed (a pageable destination
    # is staged through a bounce buffer at a fraction of it)
    img_t = torch.empty(128 + q.nblocks * 89, dtype=torch.uint8).pin_memory()
    img = img_t.numpy()
    h_ptr = h_text.data_ptr()

    def step_e2e():
        q.reset()
        q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), args.end_denoise)
        if args.e2e_per_chunk_calls:
            o = 0
            for e in ends:
                q.insert_fastq_pinned_raw(h_ptr + o, int(e) - o, args.k)
                o = int(e)
        else:       # one C call for all reference chunks of the pinned buffer (no interpreter time between chunks)
            q.insert_fastq_pinned_chunks(h_ptr, len(text), ends, args.k)
        q.finish()
        w = C.c_size_t()
        q._check(q.L.shq_export_cqf(q.h, img.ctypes.data, img.nbytes, C.byref(w)))

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    q.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    stats = q.stats()
    ndistinct, nelts = q.counts()
    rounds = q.rounds_done()

    # end to end through host buffers
    step_e2e()
    torch.cuda.synchronize()
    q.reset_stats()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e3.record(stream)
    torch.cuda.synchronize()
    ms_e2e = max(e2.elapsed_time(e3), 1000 * (time.perf_counter() - t0))
    st2 = q.stats()

    ms_step = ms / args.steps
    value = n_total / (ms_step / 1e3)
    # roofline of the dominant phase (CUDA-event time per launch, accumulated inside libshq on its stream)
    table_bytes = q.nblocks * 89
    a_ins = args.read_len / (args.read_len - args.k + 1) + 32 + 2 * table_bytes / n_total
    phases = {k: stats[k] / args.steps for k in ("ms_hash_count", "ms_scatter", "ms_aggregate", "ms_merge", "ms_layout", "ms_denoise")}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    dom = max(phases, key=phases.get)
    alg = phase_bytes(dom, args, n_total, text.nbytes, ndistinct, table_bytes, rounds)
    achieved = alg / (phases[dom] / 1e3) / 1e9
    nl = PHASE_LAUNCHES[dom]
    roofline = {"bound": "hbm", "kernel": PHASE_KERNEL[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": measured_traffic(PHASE_KERNEL[dom]), "peak_source": peak_src, "algorithmic_bytes_per_step": alg,
                "algorithmic_bytes_per_launch": alg / nl, "launches_per_step": nl, "ms_per_launch": phases[dom] / nl,
                "note": "integer/byte path bound by instruction issue and shared-memory latency, not by HBM (DESIGN.md 4); "
                        "achieved = algorithmic bytes / CUDA-event time of the phase on the handle's stream",
                "phase_ms_per_step": phases,
                "pipeline": {"bytes_per_kmer": a_ins, "achieved": value * a_ins / 1e9, "frac": value * a_ins / 1e9 / peak,
                             "note": "whole insert path against SURVEY 8(d) A_ins"}}
    line = {"metric": "canonical k-mers inserted/sec", "value": value, "unit": "k-mers/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": config_dict(args, 1, p), "clocks": clocks,
            "e2e": {"value": n_total / (ms_e2e / args.steps / 1e3), "unit": "k-mers/s", "h2d_bytes_per_step": int(st2["h2d_bytes"] // args.steps),
                    "d2h_bytes_per_step": int(st2["d2h_bytes"] // args.steps), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(stats["kernels_launched"]), "roofline": roofline,
            "result": {"ndistinct": int(ndistinct), "nelts": int(nelts), "denoise_rounds": int(rounds), "kmers_per_step": int(n_total)}}
    if args.unitigs:
        line["unitig"] = bench_unitigs(args, shq, q, text, ends, h_ptr)
    line["cpu_baseline"] = cpu_baseline(args, text, ends, p, n_total)
    print(json.dumps(line))


def bench_unitigs(args, shq, q, text, ends, h_ptr):
    """Second half of the BASELINE metric: seconds to build the unitig graph (Contiger) from the filter just
    counted, reads given as host chunks; beside it the sequential CPU restatement on the same input (1 core)."""
    import ctypes as C
    L = shq.lib()
    opts = shq.shq_unitig_opts(args.k, 2, 2, 1000000)
    best = None
    info = {}
    for rep in range(2):
        u = C.c_void_p()
        t0 = time.perf_counter()
        q._check(L.shq_unitigs_begin(q.h, C.byref(opts), C.byref(u)))
        o = 0
        for e in ends:
            q._check(L.shq_unitigs_add_reads(u, h_ptr + o, int(e) - o))
            o = int(e)
        q._check(L.shq_unitigs_build(u))
        dt = time.perf_counter() - t0
        n, tot = C.c_size_t(), C.c_size_t()
        L.shq_unitigs_count(u, C.byref(n), C.byref(tot))
        L.shq_unitigs_free(u)
        best = dt if best is None else min(best, dt)
        info = {"n_unitigs": n.value, "total_bases": tot.value}
    out = {"seconds": best, "unit": "s", "config": "Contiger -k %d -s 2 -x 2 -X 1000000 on the filter of this run, reads from pinned host chunks" % args.k}
    out.update(info)
    if args.unitig_cpu:
        from oracle import pyoracle as po
        img = q.export()
        o = po.OracleCqf.from_bytes(img)
        t0 = time.perf_counter()
        fa, nq = po.contiger_run(o, [text.tobytes()], args.k)
        out["cpu_baseline"] = {"seconds": time.perf_counter() - t0, "cores": 1, "kind": "port",
                               "sample": "whole read set, sequential restatement of Contiger over a hash-map CQF model",
                               "n_unitigs": len(po.parse_unitigs(fa)), "cqf_queries": int(nq)}
    return out


PHASE_KERNEL = {"ms_hash_count": "roll_count_kernel", "ms_scatter": "roll_scatter_kernel", "ms_aggregate": "bin_aggregate_kernel",
                "ms_merge": "merge_tiles_kernel", "ms_layout": "layout_scan_kernel + image_tiles_kernel", "ms_denoise": "denoise_tiles_kernel"}
# launches of the phase's dominant kernel per step on the default workload (4 device batches; the aggregator runs once
# more for the split at the mid-stream deNoise boundary)
PHASE_LAUNCHES = {"ms_hash_count": 4, "ms_scatter": 4, "ms_aggregate": 5, "ms_merge": 10, "ms_layout": 1, "ms_denoise": 2}
# (ms_hash_count is 0 on the default path: bins have a fixed capacity and the counting pass only runs as the exact fallback)


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture
    (profiles/r1_traffic.json, written by tools/ncu_summary.py from the same bench command); None if absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return t.get(kernel)
    except Exception:
        return None


def phase_bytes(phase, args, n_kmers, text_bytes, ndistinct, table_bytes, rounds):
    """ALGORITHMIC bytes one step moves in each phase ('kernels and rooflines')."""
    if phase == "ms_hash_count":
        return text_bytes                                  # read every FASTQ byte once
    if phase == "ms_scatter":
        return text_bytes + 4 * n_kmers                    # re-read text, write one packed 32-bit word per k-mer
    if phase == "ms_aggregate":
        return 4 * n_kmers + 14 * 2 * ndistinct            # read packed words; write + compact (key,count,chunk)
    if phase == "ms_merge":
        return 2 * 12 * 2 * ndistinct
    if phase == "ms_layout":
        return 12 * ndistinct + table_bytes
    return (rounds) * (2 * 12 * ndistinct)


def cpu_baseline(args, text, ends, p, n_total):
    from oracle import pyoracle as po
    if po.ref() is None:
        return {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    cores = os.cpu_count() or 1
    r = po.RefCqf(int(p.qb), int(p.hb))
    secs, kmers, chunks, rounds = r.count_mt(text.ctypes.data, ends, args.k, cores, trigger=int(p.trigger), rounds=int(p.num_deNoise),
                                             end_denoise=args.end_denoise, max_seconds=args.cpu_seconds)
    r.close()
    return {"value": kmers / secs, "unit": "k-mers/s", "cores": cores, "kind": "reference",
            "sample": "first %d of %d chunks (%.3g of %.3g k-mers) in %.1f s; unmodified gqf.c+nthash.hpp, std::thread worker loop" % (
                chunks, len(ends), kmers, n_total, secs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=4600000)
    ap.add_argument("--cov", type=int, default=100)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--err", type=float, default=0.002)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("--denoise", type=int, default=2)
    ap.add_argument("--end-denoise", type=int, default=1)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--unitigs", type=int, default=1, help="also time the Contiger device path (unitig build seconds)")
    ap.add_argument("--e2e-per-chunk-calls", type=int, default=0, help="e2e leg: one Python->C call per reference chunk instead of one for all")
    ap.add_argument("--unitig-cpu", type=int, default=0, help="also time the sequential CPU restatement of Contiger")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
