#!/usr/bin/env python
"""bench.py -- canonical k-mers inserted per second into the CQF (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (libshq.so, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref), no libshq

Workload at N=1 = BASELINE.json configs[2] ("celegans"): synthetic 100 Mb genome, 100x 150 bp reads, 0.234 %
substitutions, K=47, CQF-deNoise -n 1.2e8 -e 0.00234 with automatic deNoise rounds (3 mid-stream rounds, qb=29):
6.93e9 k-mers, 21 GB of FASTQ text resident in HBM.  A step = the whole counting job over that read set.
BASELINE configs[1] (E. coli-size, --deNoise=2 --endDeNoise) is measured in the same run and reported under
"secondary" (plus the unitig build on its filter).  For N>1 the genome and the filter grow with N (weak scaling):
every rank generates and hashes 1/N of the reads and owns 1/N of the quotient space.

`value`  : device-resident FASTQ text -> complete CQF image in HBM (CUDA events on the handle's stream).
`e2e`    : same job through the C ABI with HOST buffers: pinned host text, H2D copies and the D2H export of the
           .cqf image inside the timed region.
`parity` : sha256 of the exported .cqf against tests/golden/digests.json, which the UNMODIFIED reference core
           (oracle/_ref at t=1) produced for the same synthetic read set (tests/golden/make_digests.py).
"""
import argparse
import ctypes as C
import hashlib
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

import synth  # noqa: E402
from synth import workloads as wl  # noqa: E402

METRIC = "canonical k-mers inserted/sec"


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
        # the median under load: samples at (near) idle clocks between steps would only blur it
        busy = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def golden_digest(name, world):
    try:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))
        return d.get(name if world == 1 else "%s@%d" % (name, world))
    except Exception:
        return None


def config_dict(w, world, p, name):
    return {"workload": wl.describe(w, p), "name": name, "genome_bp": w["G"], "coverage": w["cov"], "read_len": w["L"],
            "error_rate": w["err"], "K": w["K"], "n_true": w["n_true"], "deNoise": w["deNoise"], "endDeNoise": bool(w["end"]),
            "num_deNoise_effective": int(p.num_deNoise), "qb": int(p.qb), "trigger": int(p.trigger), "seed": wl.SEED,
            "l2_policy": "inputs (>= 1 GB of FASTQ text per step) are larger than the 126 MB L2; no explicit flush",
            "shards": world}


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        pass
    return 6650.0, "6650 GB/s (B200_PROFILING.md fallback)"


# ----------------------------------------------------------------------------- CPU legs (oracle/_ref)
def cpu_sample(w, sp, p, seconds, cores, po):
    """A bounded sample of the workload through the reference core with `cores` host threads.
    Returns (k-mers/s, seconds, kmers, description)."""
    # the reference core does ~5e6 k-mers/s per thread: generate about 1.5x what the budget can absorb
    per_read = w["L"] - w["K"] + 1
    want = int(min(sp.n_reads, max(200000, seconds * cores * 7.5e6 / per_read)))
    text = synth.cpu_fill(sp, 0, want)
    ends = po.chunk_ends(text.ctypes.data, text.nbytes)
    r = po.RefCqf(int(p.qb), int(p.hb))
    secs, kmers, chunks, rounds = r.count_mt(text.ctypes.data, ends, w["K"], cores, trigger=int(p.trigger), rounds=int(p.num_deNoise),
                                             end_denoise=bool(w["end"]) and want == sp.n_reads, max_seconds=seconds)
    r.close()
    desc = ("first %d of %d reads generated, %d chunks (%.3g of %.3g k-mers) counted in %.1f s; table sized for the full set (qb=%d); "
            "unmodified gqf.c + nthash.hpp under the std::thread worker loop of oracle/ref_shim.cpp" % (
                want, sp.n_reads, chunks, kmers, wl.n_kmers(w), secs, p.qb))
    return kmers / secs, secs, kmers, desc


def run_reference(args, rank, world):
    """The reference's own CPU code (gqf.c + nthash.hpp, unmodified) driven by the std::thread restatement of its
    worker loop (oracle/ref_shim.cpp), all host threads, on a time-bounded sample of the same workload.
    Nothing of sh-assembly_b200 is imported here: parameters and chunking come from the oracle."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    if po.ref() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libshq_ref.so missing (built only where /root/reference exists)"}))
        return
    w = wl.scaled(args.workload, world)
    sp = synth.make(w["G"], w["cov"], w["L"], w["err"], wl.SEED)
    p = po.param_block(w["K"], w["n_true"], wl.n_kmers(w), alpha=w["err"], deNoise=w["deNoise"])
    cores = os.cpu_count() or 1
    # the whole --steps/--warmup run stays within a few minutes
    budget = max(1.0, min(args.cpu_seconds, 100.0 / max(1, args.steps + args.warmup)))
    times, rates, sample = [], [], ""
    for it in range(args.warmup + args.steps):
        rate, secs, kmers, sample = cpu_sample(w, sp, p, budget, cores, po)
        if it >= args.warmup:
            times.append(secs)
            rates.append(rate)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_dict(w, world, p, args.workload),
            "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- roofline
PHASES = {
    # phase -> (dominant kernel, launches counter, algorithmic bytes from the work counters of shq_stats)
    "ms_hash_count": ("roll_count_kernel", "scatter_launches", lambda s: s["scatter_text_bytes"]),
    # read the text once, write one packed 32-bit word per k-mer
    "ms_scatter": ("roll_scatter_kernel", "scatter_launches", lambda s: s["scatter_text_bytes"] + 4 * s["scatter_words"]),
    # second binning level: read and write every packed word once
    "ms_split": ("split_bins_kernel", "split_launches", lambda s: 8 * s["split_words"]),
    # read the packed words; write (key u64, count u32, first chunk u16) per distinct key, twice (bin order, then dense)
    "ms_aggregate": ("bin_aggregate_kernel", "agg_launches", lambda s: 4 * s["agg_words"] + 2 * 14 * s["agg_distinct"]),
    # read (key u64, count u32) of state and run, write the merged state
    "ms_merge": ("merge_tiles_lookback_kernel", "merge_launches", lambda s: 12 * (s["merge_in"] + s["merge_out"])),
    "ms_layout": ("layout_scan_kernel + image_tiles_kernel", "layout_launches", None),
    # two passes over (key, count), survivors written
    "ms_denoise": ("denoise_tiles_kernel", "denoise_launches", lambda s: 2 * 12 * s["denoise_entries"]),
}


def measured_traffic(kernel, workload=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/r2_traffic.json, written by tools/ncu_summary.py + the workload it was taken on).  Returns (bytes per launch
    or None, capture description): the per-launch figure is only comparable with this run's algorithmic bytes per launch
    when the capture was taken on the same workload -- otherwise `traffic` stays null and the capture is reported beside it."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        return None, None
    if kernel not in t:
        return None, None
    cap = {"workload": t.get("_workload"), "bytes_per_launch": t[kernel], "source": "profiles/r2_traffic.json (ncu --set full)"}
    same = workload is not None and t.get("_workload") == workload
    return (t[kernel] if same else None), cap


def roofline(stats, steps, value, a_ins, table_bytes, workload=None):
    peak, peak_src = peak_hbm()
    phases = {k: stats[k] / steps for k in PHASES}
    dom = max(phases, key=phases.get)
    kernel, lkey, fn = PHASES[dom]
    if fn is None:
        alg = 12 * stats["layout_entries"] + table_bytes * stats["layout_launches"]
    else:
        alg = fn(stats)
    nl = max(1, stats[lkey])
    ms_launch = stats[dom] / nl
    achieved = (alg / nl) / (ms_launch / 1e3) / 1e9 if ms_launch > 0 else 0.0
    traffic, capture = measured_traffic(kernel, workload)
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_capture": capture, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg / nl, "launches_per_step": nl / steps, "ms_per_launch": ms_launch,
            "note": "achieved = algorithmic bytes of the phase's launches (work counters of shq_stats, summed per launch) / CUDA-event time "
                    "of the phase on the handle's stream; the path is bound by instruction issue and shared-memory latency (DESIGN.md 4)",
            "phase_ms_per_step": phases,
            "pipeline": {"bytes_per_kmer": a_ins, "achieved": value * a_ins / 1e9, "frac": value * a_ins / 1e9 / peak,
                         "note": "whole insert path against SURVEY 8(d) A_ins = L/(L-K+1) + 32 + 2T/N"}}


# ----------------------------------------------------------------------------- our arm, one GPU
def run_workload(args, name, steps, warmup, local_rank, want_unitigs, want_cpu):
    import torch
    shq = importlib.import_module("sh-assembly_b200")
    w = wl.scaled(name, 1)
    sp = synth.make(w["G"], w["cov"], w["L"], w["err"], wl.SEED)
    n_total = wl.n_kmers(w)
    p = shq.param_block(w["K"], w["n_true"], n_total, alpha=w["err"], deNoise=w["deNoise"])
    K = w["K"]
    dev = torch.device("cuda", local_rank)
    n_bytes = sp.n_reads * sp.rec_bytes
    # synthetic reads straight into HBM, then one copy to pinned host memory (the e2e leg and the chunker read it there)
    d_text = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)
    synth.gpu_fill(sp, 0, sp.n_reads, d_text.data_ptr(), torch.cuda.current_stream().cuda_stream)
    h_text = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_text.copy_(d_text[:n_bytes])
    torch.cuda.synchronize()
    text = h_text.numpy()
    # reference chunking (host reader, 8 MiB parts cut at a record)
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

    q = shq.CQF(int(p.qb), int(p.hb), device=local_rank)
    stream = torch.cuda.ExternalStream(q.stream(), device=dev)

    def step_device():
        q.reset()
        q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), w["end"])
        q.insert_fastq_device(d_text.data_ptr(), n_bytes, ends, K)
        q.finish()
        q.materialize()

    # destination of the exported .cqf image: page-locked, so the D2H copy runs at PCIe speed
    img_t = torch.empty(128 + q.nblocks * 89, dtype=torch.uint8).pin_memory()
    img = img_t.numpy()
    h_ptr = h_text.data_ptr()

    def step_e2e():
        q.reset()
        q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), w["end"])
        q.insert_fastq_pinned_chunks(h_ptr, n_bytes, ends, K)
        q.finish()
        wr = C.c_size_t()
        q._check(q.L.shq_export_cqf(q.h, img.ctypes.data, img.nbytes, C.byref(wr)))

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    q.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    stats = q.stats()
    ndistinct, nelts = q.counts()
    rounds = q.rounds_done()

    # end to end through host buffers
    del d_text
    torch.cuda.empty_cache()
    step_e2e()
    torch.cuda.synchronize()
    q.reset_stats()
    e2e_steps = max(1, min(steps, args.e2e_steps))
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e3.record(stream)
    torch.cuda.synchronize()
    ms_e2e = max(e2.elapsed_time(e3), 1000 * (time.perf_counter() - t0)) / e2e_steps
    st2 = q.stats()

    # the floor of the e2e leg on this box: the same bytes through the same host link with nothing else to do
    scratch = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record()
    for off in range(0, n_bytes, 1 << 30):
        m = min(1 << 30, n_bytes - off)
        scratch[:m].copy_(h_text[off:off + m], non_blocking=True)
    f1.record()
    torch.cuda.synchronize()
    h2d_floor_ms = f0.elapsed_time(f1)
    del scratch

    # parity: the image the e2e leg just exported against the digest of the unmodified reference core
    gold = golden_digest(name, 1)
    sha = hashlib.sha256(memoryview(img)).hexdigest()
    parity = {"image_sha256": sha, "expected": gold["image_sha256"] if gold else None,
              "match": bool(gold) and sha == gold["image_sha256"] and int(ndistinct) == gold["ndistinct"] and int(nelts) == gold["nelts"]
              and int(rounds) == gold["denoise_rounds"],
              "checker": gold["made_by"] if gold else "no committed digest for this workload (tests/golden/digests.json)"}
    if gold and not parity["match"]:
        parity["expected_counters"] = {k: gold[k] for k in ("ndistinct", "nelts", "denoise_rounds")}
        print("bench.py: PARITY FAILED on %s: %s" % (name, json.dumps(parity)), file=sys.stderr)

    ms_step = ms / steps
    value = n_total / (ms_step / 1e3)
    table_bytes = q.nblocks * 89
    a_ins = w["L"] / (w["L"] - K + 1) + 32 + 2 * table_bytes / n_total
    line = {"metric": METRIC, "value": value, "unit": "k-mers/s", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": config_dict(w, 1, p, name), "clocks": clocks,
            "e2e": {"value": n_total / (ms_e2e / 1e3), "unit": "k-mers/s", "h2d_bytes_per_step": int(st2["h2d_bytes"] // e2e_steps),
                    "d2h_bytes_per_step": int(st2["d2h_bytes"] // e2e_steps), "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "h2d_floor_ms": h2d_floor_ms, "batches_per_step": st2["batches"] / e2e_steps, "wide_batches": int(st2["wide_batches"]),
                    "wide_fallbacks": int(st2["wide_fallbacks"]), "spec_rejected": int(st2["spec_rejected"]),
                    "gpu_busy_ms_per_step": sum(st2[k] for k in st2 if k.startswith("ms_")) / e2e_steps},
            "gpu_launches": int(stats["kernels_launched"]), "roofline": roofline(stats, steps, value, a_ins, table_bytes, name),
            "result": {"ndistinct": int(ndistinct), "nelts": int(nelts), "denoise_rounds": int(rounds), "kmers_per_step": int(n_total),
                       "batches_per_step": stats["batches"] / steps, "host_syncs_per_step": stats["host_syncs"] / steps,
                       "bin_fallbacks": int(stats["bin_fallbacks"]), "wide_batches": int(stats["wide_batches"]),
                       "wide_fallbacks": int(stats["wide_fallbacks"]), "spec_rejected": int(stats["spec_rejected"])},
            "parity": parity}
    if want_unitigs:
        line["unitig"] = bench_unitigs(args, shq, q, text, ends, h_ptr, K)
        line["cli"] = bench_cli(args, shq, text, w, p, name)
    if want_cpu:
        from oracle import pyoracle as po
        if po.ref() is None:
            line["cpu_baseline"] = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
        else:
            cores = os.cpu_count() or 1
            rate, secs, kmers, desc = cpu_sample(w, sp, p, args.cpu_seconds, cores, po)
            line["cpu_baseline"] = {"value": rate, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": desc}
    q.close()
    del h_text, img_t
    torch.cuda.empty_cache()
    return line


def bench_unitigs(args, shq, q, text, ends, h_ptr, K):
    """Second half of the BASELINE metric: seconds to build the unitig graph (Contiger) from the filter just
    counted, reads given as host chunks; beside it the sequential CPU restatement of Contiger (1 core)."""
    L = shq.lib()
    opts = shq.shq_unitig_opts(K, 2, 2, 1000000)
    times = []
    info = {}
    for rep in range(4):      # the first build allocates the device buffers; they stay with the handle afterwards
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
        times.append(dt)
        info = {"n_unitigs": n.value, "total_bases": tot.value}
    best = min(times[1:])
    steps = max(0, info["total_bases"] - (K - 1) * info["n_unitigs"])       # k-mers appended to unitigs = extension steps
    peak, peak_src = peak_hbm()
    out = {"seconds": best, "unit": "s", "seconds_all_runs": times,
           "config": "Contiger -k %d -s 2 -x 2 -X 1000000 on the filter of this run, reads from pinned host chunks (H2D inside the time)" % K,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "achieved": steps * 623 / best / 1e9, "frac": steps * 623 / best / 1e9 / peak,
                        "extension_steps": steps, "note": "SURVEY 8(d) A_ext = 7 probes x 89 bytes per k-mer appended to a unitig; whole build incl. H2D of the reads"}}
    out.update(info)
    if args.unitig_cpu:
        from oracle import pyoracle as po
        img = q.export()
        o = po.OracleCqf.from_bytes(img)
        t0 = time.perf_counter()
        fa, nq = po.contiger_run(o, [text.tobytes()], K)
        out["cpu_baseline"] = {"seconds": time.perf_counter() - t0, "cores": 1, "kind": "port",
                               "sample": "whole read set; sequential restatement of Contiger (oracle/contiger_oracle.cpp, pinned to the reference "
                                         "binary's output by tests/test_contiger_ref.py) over a hash-map CQF model",
                               "n_unitigs": len(po.parse_unitigs(fa)), "cqf_queries": int(nq)}
    return out


def bench_cli(args, shq, text, w, p, name):
    """file -> .cqf through the drop-in command line (bin/CQF-deNoise): the read set written to tmpfs, wall clock of the
    whole process (CUDA context, reader threads, page-locked chunk buffers, asynchronous copies, counting, shq_save)."""
    import tempfile
    d = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(dir=d)
    try:
        fq = os.path.join(tmp, "reads.fq")
        with open(fq, "wb") as f:
            f.write(memoryview(text))
        with open(os.path.join(tmp, "list.txt"), "w") as f:
            f.write("reads.fq\n")
        exe = os.path.join(ROOT, "sh-assembly_b200", "bin", "CQF-deNoise")
        cmd = [exe, "-k", str(w["K"]), "-n", str(w["n_true"]), "-N", str(wl.n_kmers(w)), "-e", str(w["err"]), "--deNoise", str(w["deNoise"]),
               "-t", "16", "-f", "f", "-i", os.path.join(tmp, "list.txt"), "-o", os.path.join(tmp, "out.cqf")]
        if w["end"]:
            cmd.append("--endDeNoise")
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if r.returncode:
                return {"error": (r.stderr or "")[-300:]}
            best = dt if best is None else min(best, dt)
        sha = hashlib.sha256(open(os.path.join(tmp, "out.cqf"), "rb").read()).hexdigest()
        gold = golden_digest(name, 1)
        return {"seconds": best, "unit": "s", "kmers_per_s": wl.n_kmers(w) / best, "command": " ".join(cmd[:1] + cmd[1:13]) + " ...",
                "input": "plain FASTQ on tmpfs, %d bytes" % text.nbytes, "image_sha256_matches_reference": bool(gold) and sha == gold["image_sha256"]}
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libshq has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        mg = importlib.import_module("sh-assembly_b200.multigpu")
        return mg.bench_sharded(args, rank, world, local_rank, sys.modules[__name__])
    line = run_workload(args, args.workload, args.steps, args.warmup, local_rank,
                        want_unitigs=args.unitigs and args.workload.startswith("ecoli"), want_cpu=True)
    if args.secondary and args.secondary != args.workload:
        s_steps, s_warm = max(1, min(args.steps, 5)), max(3, min(args.warmup, 3))
        sec = run_workload(args, args.secondary, s_steps, s_warm, local_rank, want_unitigs=bool(args.unitigs), want_cpu=False)
        line["secondary"] = {k: sec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e", "roofline", "result", "parity")
                             if k in sec}
        if "unitig" in sec:
            line["unitig"] = sec["unitig"]
            line["unitig"]["workload"] = args.secondary
        if "cli" in sec:
            line["cli"] = sec["cli"]
            line["cli"]["workload"] = args.secondary
    print(json.dumps(line))
    if not line["parity"]["match"] and line["parity"]["expected"]:
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="celegans", choices=sorted(wl.WORKLOADS))
    ap.add_argument("--secondary", default="ecoli", help="second workload reported under 'secondary' at N=1 ('' = none)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed steps of the end-to-end leg (at most --steps)")
    ap.add_argument("--unitigs", type=int, default=1, help="also time the Contiger device path (unitig build seconds)")
    ap.add_argument("--unitig-cpu", type=int, default=1, help="also time the sequential CPU restatement of Contiger (about 20 s on the E. coli-size set)")
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
