"""k-mer kernel alone on a stream-sized batch (default 200 000 x 150 bp, k = 4..5: 1 GB of counts):
CUDA events around three launches. SQW_KMER_WARP64=1 selects the 64-bit warp kernel."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sequnwinder_b200 import synth, KmerCounter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
kmin, kmax = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 5)
d = synth.generate(n, 8, seed=7)
kc = KmerCounter(kmin, kmax)
b = torch.from_numpy(d.bases).cuda(); o = torch.from_numpy(d.offsets).cuda()
out, hn, st = kc.count_device(b, o, 150)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): kc.count_device(b, o, 150, out=out, has_n=hn, status=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
nbytes = n * (150 + 4 * kc.numK)
print('n', n, 'k', kmin, kmax, 'kernel %.3f ms  %.1f GB/s  %.2f Gbases/s  row sums ok %s' % (ms, nbytes / ms / 1e6, n * 150 / ms / 1e6, bool((out.sum(1) == sum(150 - k + 1 for k in range(kmin, kmax + 1))).all())))
