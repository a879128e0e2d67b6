"""Where the host-side time of bench.py's e2e step goes (cfg-2): wall clock of KmerCounter.count,
of opening the optimizer (problem upload) and of execute, next to the device-timed seconds."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from sequnwinder_b200 import KmerCounter, classifier as clf
d, crs, cls, w, C = bench.build_problem(1)
kc = KmerCounter(bench.KMIN, bench.KMAX)
counts, hn = kc.count(d.bases, d.offsets)
prep = clf.prepare(counts, cls, w, crs, C)
for rep in range(3):
    t0 = time.time(); c, hn = kc.count(d.bases, d.offsets); t1 = time.time()
    lo = bench.make_optimizer(prep, crs, cls, w, C, 8); t2 = time.time()
    lo._open(); t3 = time.time()
    lo.execute(); t4 = time.time()
    _ = float(lo.getX()[0]); t5 = time.time()
    s = lo.stats
    print('rep %d: kmer.count %.3f  make_optimizer %.3f  open(upload) %.3f  execute %.3f (device total %.3f, admm %.3f)  getX %.4f  | e2e %.3f'
          % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, s.seconds_total, s.seconds_admm, t5 - t4, t5 - t0), flush=True)
