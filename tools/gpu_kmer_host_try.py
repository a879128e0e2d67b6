"""Host-buffer k-mer call (sqw_kmer_count) at cfg-2: ms per call over repeated calls into the same
output matrix and into a fresh one. Tuning hooks: SQW_KMER_NT, SQW_KMER_PIECES, SQW_KMER_THREADS."""
import sys, time, os, numpy as np
sys.path.insert(0, '.')
from sequnwinder_b200 import KmerCounter, synth
d = synth.generate_config('cfg2')
kc = KmerCounter(4, 5)
for _ in range(3):
    out, hn = kc.count(d.bases, d.offsets)
for mode in ('fresh output', 'reused output'):
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        out, hn = kc.count(d.bases, d.offsets, out=(out if mode == 'reused output' else None))
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print('NT=%s PIECES=%s THREADS=%s %s: median %.2f ms, best %.2f ms -> %.2f Gbases/s (median)' % (
        os.environ.get('SQW_KMER_NT', '-'), os.environ.get('SQW_KMER_PIECES', '-'), os.environ.get('SQW_KMER_THREADS', '-'),
        mode, ts[len(ts) // 2] * 1e3, ts[0] * 1e3, d.bases.size / ts[len(ts) // 2] / 1e9), flush=True)
