"""Host-buffer k-mer call (sqw_kmer_count): ms per call over repeated calls into the same output
matrix and into a fresh one. usage: gpu_kmer_host_try.py [cfg2|cfg3|cfg4 [rows]]  (default cfg2 at
full size; cfg-3 / cfg-4 rows are 87 KB / 21.8 KB of int32 each, so give a row count)."""
import sys, time, os, numpy as np
sys.path.insert(0, '.')
from sequnwinder_b200 import KmerCounter, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
rows = int(sys.argv[2]) if len(sys.argv) > 2 else None
c = synth.CONFIGS[cfg]
d = synth.generate_config(cfg, rows)
kc = KmerCounter(c['kmin'], c['kmax'])
print(cfg, 'rows', len(d.offsets) - 1, 'k', c['kmin'], c['kmax'], 'numK', kc.numK,
      'dense output MB', (len(d.offsets) - 1) * kc.numK * 4 / 1e6, flush=True)
for _ in range(2):
    out, hn = kc.count(d.bases, d.offsets)
for mode in ('fresh output', 'reused output'):
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        out, hn = kc.count(d.bases, d.offsets, out=(out if mode == 'reused output' else None))
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print('%s: median %.2f ms, best %.2f ms -> %.3f Gbases/s, %.2f GB/s of dense output (median)' % (
        mode, ts[len(ts) // 2] * 1e3, ts[0] * 1e3, d.bases.size / ts[len(ts) // 2] / 1e9,
        out.nbytes / ts[len(ts) // 2] / 1e9), flush=True)
