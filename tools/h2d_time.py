"""Pure pinned-host -> device copy time of the bench workload's FASTQ text (the PCIe floor of the e2e leg)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
text, _, _ = bench.gen_fastq(4600000, 100, 150, 0.002, 42)
h = torch.from_numpy(text).pin_memory(); d = torch.empty(len(text), dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("H2D of %d bytes: %.2f ms (%.1f GB/s)" % (len(text), ms, len(text) / ms / 1e6))
