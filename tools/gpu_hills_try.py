"""Hills scan timing on the GPU box: device-resident sqw_hills_scan_dev with CUDA
events vs the CPU oracle on a sample. usage: python tools/gpu_hills_try.py [cfg] [n]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sequnwinder_b200 import synth
from sequnwinder_b200.hills import HillsScanner
from sequnwinder_b200.kmer import num_kmers
from oracle import pyoracle as O

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
c = synth.CONFIGS[cfg]
d = synth.generate_config(cfg, n)
rng = np.random.default_rng(2)
w = rng.normal(0.0, 0.08, num_kmers(c["kmin"], c["kmax"]))
hs = HillsScanner(c["kmin"], c["kmax"], 6, 14, 0.1)
dev = torch.device("cuda:0")
bases, offsets, wd = (torch.from_numpy(a).to(dev) for a in (d.bases, d.offsets, w))
nseq = len(d.offsets) - 1
L = int(np.diff(d.offsets).max())
for mode in (0, 1):
    out = hs.scan_device(bases, offsets, L, wd, mode=mode, cap=64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        out = hs.scan_device(bases, offsets, L, wd, mode=mode, cap=64)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    windows = sum(max(0, L - l + 1) for l in range(6, 15)) * nseq
    print(f"{cfg} mode {mode}: {nseq} x {L} bp, {int(out[3].sum())} hills, {ms:.3f} ms per scan "
          f"({nseq * L / ms / 1e6:.2f} Gbases/s, {windows / ms / 1e6:.1f} G windows/s)")
O.build()
ns = min(nseq, 2000)
t0 = time.time()
O.hills_scan(d.bases[: d.offsets[ns]], d.offsets[: ns + 1], c["kmin"], c["kmax"], w, 6, 14, 0.1, 0, 64)
dt = time.time() - t0
print(f"oracle (1 core): {ns} sequences in {dt:.2f} s -> {dt / ns * nseq:.1f} s for all; "
      f"{ns * L / dt / 1e9:.5f} Gbases/s")
