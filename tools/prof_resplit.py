"""Single-GPU stand-in for the owner side of the multi-GPU exchange (for ncu, which cannot wrap a multi-rank job):
bin a batch over the whole key space into coarse bins, then split + insert it as a 1-rank 'exchange'."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
shq = importlib.import_module("sh-assembly_b200")
mg = importlib.import_module("sh-assembly_b200.multigpu")

text, _, _ = bench.gen_fastq(1200000, 100, 150, 0.002, 42)
d = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda"); d[:len(text)].copy_(torch.from_numpy(text)); torch.cuda.synchronize()
fb = int(os.environ.get("FB", "4")); cb = 12 - fb
q = shq.CQF(27, 35)
for it in range(3):
    q.reset(); q.reset_stats()
    wp, op, nw = q.mg_bin_fastq_device(d.data_ptr(), len(text), 31, cb)
    off = mg.dev_view(torch, op, (1 << cb) + 1, 0)
    counts = (off[1:] - off[:-1]).contiguous()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    q.mg_insert_binned(wp, np.zeros(1, np.uint64), counts.data_ptr(), 1, cb, fb, nw)
    q.sync(); t1 = time.perf_counter()
    st = q.stats()
    print("words", nw, "insert_binned %.2f ms" % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")})
