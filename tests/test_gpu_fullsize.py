"""BASELINE configurations at their REAL size on the GPU, against digests the unmodified reference core produced
(tests/golden/digests.json <- tests/golden/make_digests.py, oracle/_ref at t=1 in the build container).

configs[0] and configs[1] run whole; configs[2] runs as its 20 Mb twin here (the 100 Mb set is checked by bench.py itself,
same digest file).  Marked `slow`: a few seconds each on a B200, dominated by generating and copying the reads."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import synth
from synth import workloads as wl

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "digests.json")


def gold(name):
    d = json.load(open(GOLD))
    if name not in d:
        pytest.skip("no committed digest for %s" % name)
    return d[name]


def device_text(sp, first=0, n=None):
    import torch
    n = sp.n_reads if n is None else n
    nb = n * sp.rec_bytes
    d = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
    synth.gpu_fill(sp, first, n, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d, nb


def test_gpu_generator_matches_c_twin():
    """the CUDA generator and the C generator emit the same bytes (so the digests apply to what the GPU counts)"""
    import torch
    sp = synth.make(300000, 40, 150, 0.01, 11)
    d, nb = device_text(sp)
    g = d[:nb].cpu().numpy()
    c = synth.cpu_fill(sp, 0, sp.n_reads)
    assert g.tobytes() == c.tobytes()
    # a slice that starts in the middle of the set, odd sizes
    d2, nb2 = device_text(sp, 12345, 777)
    assert d2[:nb2].cpu().numpy().tobytes() == synth.cpu_fill(sp, 12345, 777).tobytes()


@pytest.mark.parametrize("name", ["ecoli0", "ecoli", "celegans20"])
def test_baseline_config_matches_reference_digest(shq, name):
    import torch
    g = gold(name)
    w = wl.scaled(name, 1)
    sp = synth.make(w["G"], w["cov"], w["L"], w["err"], wl.SEED)
    N = wl.n_kmers(w)
    p = shq.param_block(w["K"], w["n_true"], N, alpha=w["err"], deNoise=w["deNoise"])
    assert (int(p.qb), int(p.num_deNoise), int(p.trigger)) == (g["qb"], g["num_deNoise"], g["trigger"])
    d, nb = device_text(sp)
    assert nb == g["text_bytes"]
    host = d[:nb].cpu().numpy()
    assert hashlib.sha256(memoryview(host)).hexdigest() == g["text_sha256"]
    rd = shq.Reader(buffers=[host])
    ends, off = [], 0
    while True:
        r = rd.next_raw()
        if r is None:
            break
        off += r[1]
        ends.append(off)
    rd.close()
    assert len(ends) == g["n_chunks"]
    q = shq.CQF(int(p.qb), int(p.hb), device=0)
    q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), w["end"])
    q.insert_fastq_device(d.data_ptr(), nb, np.array(ends, np.uint64), w["K"])
    q.finish()
    assert q.counts() == (g["ndistinct"], g["nelts"])
    assert q.rounds_done() == g["denoise_rounds"]
    img = q.export()
    assert len(img) == g["image_bytes"]
    assert hashlib.sha256(img).hexdigest() == g["image_sha256"]
    # the same job through the host path of the CLI (reader -> shq_build_kmer_spectrum) gives the same image
    q2 = shq.CQF(int(p.qb), int(p.hb), device=0)
    q2.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), w["end"])
    rd = shq.Reader(buffers=[host])
    q2.build_kmer_spectrum(rd, w["K"])
    rd.close()
    assert hashlib.sha256(q2.export()).hexdigest() == g["image_sha256"]
    q.close()
    q2.close()


def test_ecoli_unitigs_match_oracle(shq, po):
    """Contiger on the full configs[1] filter: canonical unitig set and link structure equal the CPU restatement's"""
    name = "ecoli"
    w = wl.scaled(name, 1)
    sp = synth.make(w["G"], w["cov"], w["L"], w["err"], wl.SEED)
    p = shq.param_block(w["K"], w["n_true"], wl.n_kmers(w), alpha=w["err"], deNoise=w["deNoise"])
    host = synth.cpu_fill(sp, 0, sp.n_reads)
    q = shq.CQF(int(p.qb), int(p.hb), device=0)
    q.set_denoise_schedule(int(p.trigger), int(p.num_deNoise), w["end"])
    rd = shq.Reader(buffers=[host])
    q.build_kmer_spectrum(rd, w["K"])
    rd.close()
    img = q.export()
    chunks = [ch for _, ch in shq.Reader(buffers=[host])]
    fa = shq.build_unitigs(q, chunks, w["K"])
    o = po.OracleCqf.from_bytes(img)
    fo, _ = po.contiger_run(o, [host.tobytes()], w["K"])
    assert po.canonical_unitig_set(fa) == po.canonical_unitig_set(fo)
    assert po.canonical_links(fa) == po.canonical_links(fo)
    q.close()
