"""Drop-in boundary end to end: the CQF-deNoise and Contiger command lines (same options as the reference's,
src/CQF-deNoise.cpp:21-34, src/contig_assembly.cpp:25-37) over a list file with a plain and a gzip FASTQ.
The .cqf must be byte-identical to the oracle's run with the same parameter block; the unitig set and links
must equal the sequential restatement's.  -m gpu."""
import gzip
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sh-assembly_b200", "bin")


@pytest.mark.parametrize("fmode", ["f", "g"])
def test_cqf_denoise_then_contiger(po, shq, tmp_path, fmode):
    K = 31
    files = [po.synth_fastq(G=60000, cov=25, err=0.004, n_rate=0.0003, seed=201 + i) for i in range(2)]
    names = []
    for i, f in enumerate(files):
        name = "reads%d.fq%s" % (i, ".gz" if fmode == "g" else "")
        with (gzip.open if fmode == "g" else open)(str(tmp_path / name), "wb") as fh:
            fh.write(f)
        names.append(name)
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    keys = np.concatenate([po.reads_to_keys(f, K, 40) for f in files])
    n_total = len(keys)
    n_true = 2 * 60000
    out = str(tmp_path / "out.cqf")
    r = subprocess.run([os.path.join(BIN, "CQF-deNoise"), "-k", str(K), "-n", str(n_true), "-N", str(n_total), "-e", "0.004",
                        "--deNoise", "1", "--endDeNoise", "-t", "4", "-f", fmode, "-i", str(tmp_path / "list.txt"), "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    p = shq.param_block(K, n_true, n_total, alpha=0.004, deNoise=1)
    o, rounds = po.count_run(files, K, int(p.qb), int(p.hb), int(p.trigger), int(p.num_deNoise), True)
    assert open(out, "rb").read() == o.export()
    # Contiger takes the file names verbatim (contig_assembly.cpp:252-258)
    (tmp_path / "list2.txt").write_text("\n".join(str(tmp_path / n) for n in names) + "\n")
    fa = str(tmp_path / "unitigs.fa")
    r = subprocess.run([os.path.join(BIN, "Contiger"), "-k", str(K), "-i", str(tmp_path / "list2.txt"), "-f", fmode, "-c", out, "-o", fa],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    fa_o, _ = po.contiger_run(o, files, K, 2, 2, 1000000)
    got = open(fa).read()
    assert po.canonical_unitig_set(got) == po.canonical_unitig_set(fa_o)
    assert po.canonical_links(got) == po.canonical_links(fa_o)
