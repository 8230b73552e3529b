"""CPU suite: host-side C++ mirror (reader/chunker, parameter block, CLIs) against the oracle restatement."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reader_matches_oracle_chunker(po, shq):
    files = [po.synth_fastq(G=30000, cov=40, seed=s, L=100 + 10 * s) for s in (1, 2, 3)]
    for part in (1 << 18, 1 << 20, 0):
        want = list(po.chunks(files, part_size=part or (1 << 23)))
        got = list(shq.Reader(buffers=files, part_size=part))
        assert [i for i, _ in got] == [i for i, _ in want]
        assert [c for _, c in got] == [c for _, c in want]


def test_reader_files_plain_and_gzip(po, shq, tmp_path):
    text = po.synth_fastq(G=40000, cov=30, seed=5)
    p1, p2 = str(tmp_path / "a.fq"), str(tmp_path / "a.fq.gz")
    open(p1, "wb").write(text)
    with gzip.open(p2, "wb", compresslevel=1) as f:
        f.write(text)
    plain = b"".join(c for _, c in shq.Reader(paths=[p1], fmode="f"))
    gz = b"".join(c for _, c in shq.Reader(paths=[p2], fmode="g"))
    assert plain == text and gz == text
    # unreadable files are skipped silently, like seqFile_batch (cqf/CQF_mt.h:350-360)
    assert list(shq.Reader(paths=[str(tmp_path / "missing.fq")], fmode="f")) == []


def test_prefetching_file_reader_keeps_the_serial_chunk_order(po, shq, tmp_path):
    """Three gzip files of different lengths, several 8 MiB parts each: the per-file producer threads must hand out
    exactly the chunk sequence (file order and bytes) of the sequential chunker."""
    files = [po.synth_fastq(G=g, cov=40, seed=70 + i) for i, g in enumerate((400000, 150000, 300000))]
    paths = []
    for i, f in enumerate(files):
        paths.append(str(tmp_path / ("r%d.fq.gz" % i)))
        with gzip.open(paths[-1], "wb", compresslevel=1) as fh:
            fh.write(f)
    want = list(po.chunks(files))
    assert len(want) > 6
    for _ in range(2):
        got = list(shq.Reader(paths=paths, fmode="g"))
        assert [i for i, _ in got] == [i for i, _ in want]
        assert [c for _, c in got] == [c for _, c in want]
    # closing a reader whose producers are still ahead must not hang
    r = shq.Reader(paths=paths, fmode="g")
    next(r)
    r.close()


def test_reader_bzip2_via_dlopen(po, shq, tmp_path):
    import bz2
    text = po.synth_fastq(G=20000, cov=20, seed=6)
    p = str(tmp_path / "a.fq.bz2")
    open(p, "wb").write(bz2.compress(text))
    try:
        r = shq.Reader(paths=[p], fmode="b")
    except shq.ShqError:
        pytest.skip("libbz2.so.1 not present")
    assert b"".join(c for _, c in r) == text


@pytest.mark.parametrize("args", [dict(K=31, n=4600000, N=367999920, alpha=0.002, deNoise=0), dict(K=31, n=4600000, N=367999920, alpha=0.002, deNoise=2),
                                  dict(K=47, n=120000000, N=16506371070, alpha=0.00234, deNoise=-1), dict(K=47, n=120000000, N=6930000000, alpha=0.00234, deNoise=-1),
                                  dict(K=31, n=3000000000, N=72000000000, alpha=0.002, deNoise=0), dict(K=21, n=1000, N=50000, alpha=0.01, deNoise=-1, fr=0.001),
                                  dict(K=25, n=119157843, N=16506371070, alpha=0.00234, deNoise=5)])
def test_param_block_matches_oracle(po, shq, args):
    a = shq.param_block(args["K"], args["n"], args["N"], alpha=args["alpha"], deNoise=args["deNoise"], fr=args.get("fr", 0.0))
    b = po.param_block(args["K"], args["n"], args["N"], alpha=args["alpha"], deNoise=args["deNoise"], fr=args.get("fr", 0.0))
    for f in ("num_true_kmers", "num_false_kmers", "num_deNoise", "lower_bound", "upper_bound", "qb", "hb", "trigger"):
        assert getattr(a, f) == getattr(b, f), f
    assert a.fr == b.fr


def test_param_block_error_profile(po, shq, tmp_path):
    errs = np.linspace(0.001, 0.01, 150)
    p = str(tmp_path / "err.txt")
    open(p, "w").write("\n".join("%.6f" % e for e in errs))
    vals = np.array([float("%.6f" % e) for e in errs])      # keep alive across the ctypes call
    t = po.oracle().orc_true2falseKmer_DP(vals.ctypes.data, 150, 31)
    a = shq.param_block(31, 4600000, 367999920, alpha=-1.0, error_profile=p, deNoise=2)
    b = po.param_block(31, 4600000, 367999920, alpha=-1.0, true2false=t, deNoise=2)
    assert (a.qb, a.num_deNoise, a.trigger, a.num_true_kmers) == (b.qb, b.num_deNoise, b.trigger, b.num_true_kmers)


def test_cli_help_and_required_options(shq):
    exe = os.path.join(ROOT, "sh-assembly_b200", "bin", "CQF-deNoise")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "--endDeNoise" in r.stderr and "-k arg" in r.stderr
    r = subprocess.run([exe, "-k", "31"], capture_output=True, text=True)
    assert r.returncode == 0 and "Please specify either <alpha> or <errorProfile>" in r.stderr
    r = subprocess.run([exe, "-k", "31", "-e", "0.002"], capture_output=True, text=True)
    assert r.returncode != 0 and "required" in r.stderr
    exe2 = os.path.join(ROOT, "sh-assembly_b200", "bin", "Contiger")
    r = subprocess.run([exe2, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--solid_abundance_max" in r.stderr and "(=1000000)" in r.stderr
