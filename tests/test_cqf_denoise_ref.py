"""The UNMODIFIED reference CQF-deNoise command line (oracle/_ref/CQF-deNoise_ref, built against oracle/shim by
`make -C oracle cqf_denoise_ref`) against the oracle: same settings banner, byte-identical .cqf at --deNoise=0 -t 1."""
import os
import subprocess
import tempfile

import pytest

import synth

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "CQF-deNoise_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/CQF-deNoise_ref not built (needs /root/reference)")


@pytest.mark.parametrize("K,G,cov,err", [(31, 200000, 40, 0.004), (47, 120000, 30, 0.002)])
def test_reference_cli_image_equals_oracle(po, K, G, cov, err):
    sp = synth.make(G, cov, 150, err, 7)
    text = synth.cpu_fill(sp, 0, sp.n_reads)
    N = sp.n_reads * (150 - K + 1)
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "reads.fq"), "wb").write(text.tobytes())
        open(os.path.join(tmp, "list.txt"), "w").write("reads.fq\n")
        out = os.path.join(tmp, "ref.cqf")
        r = subprocess.run([BIN, "-k", str(K), "-n", str(G), "-N", str(N), "-e", str(err), "--deNoise", "0", "-t", "1", "-f", "f",
                            "-i", os.path.join(tmp, "list.txt"), "-o", out], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, MALLOC_PERTURB_="255"))
        assert r.returncode == 0, r.stderr[-400:]
        ref = open(out, "rb").read()
    p = po.param_block(K, G, N, alpha=err, deNoise=0)
    assert "qb: %d\n" % p.qb in r.stderr and "hb: %d\n" % p.hb in r.stderr
    assert "number of unique k-mers triggering deNoise: %d\n" % p.trigger in r.stderr
    o, rounds = po.count_run([text.tobytes()], K, int(p.qb), int(p.hb), int(p.trigger), 0, False)
    assert rounds == 0
    assert ref == o.export()          # 128-byte qfmetadata + every block, byte for byte


def test_reference_cli_banner_for_the_baseline_configs(po):
    """parameter block of the BASELINE configurations as the reference binary prints it (it exits on the missing input
    list only after the banner would be printed, so a dummy list with no readable file is enough)"""
    from synth import workloads as wl
    for name in ("ecoli0", "ecoli", "celegans"):
        w = wl.scaled(name, 1)
        N = wl.n_kmers(w)
        p = po.param_block(w["K"], w["n_true"], N, alpha=w["err"], deNoise=w["deNoise"])
        with tempfile.TemporaryDirectory() as tmp:
            open(os.path.join(tmp, "list.txt"), "w").write("missing.fq\n")
            cmd = [BIN, "-k", str(w["K"]), "-n", str(w["n_true"]), "-N", str(N), "-e", str(w["err"]), "--deNoise", str(w["deNoise"]), "-t", "1",
                   "-f", "f", "-i", os.path.join(tmp, "list.txt"), "-o", os.path.join(tmp, "o.cqf")]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=dict(os.environ, MALLOC_PERTURB_="255"))
                err = r.stderr
            except subprocess.TimeoutExpired as e:
                err = (e.stderr or b"").decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
        assert "qb: %d\n" % p.qb in err, (name, err[-500:])
        assert "number of deNoise rounds: %d\n" % p.num_deNoise in err
        assert "number of unique k-mers triggering deNoise: %d\n" % p.trigger in err
