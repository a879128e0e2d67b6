"""Throughput of the streamed packed-count kernels (MODE 10) at cfg-3 / cfg-4 feature shapes on
synthetic count matrices (uint8 counts 0..3, a few ADMM iterations, no oracle).

usage: gpu_chunk_perf.py NAME ROWS FEATURES CLASSES [ADMM_ITERS]
       e.g. cfg3 80000 10920 12   |   cfg4 200000 2728 17"""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt, structure
name, n, F, C = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
A = int(sys.argv[5]) if len(sys.argv) > 5 else 1
T = 8
rng = np.random.default_rng(0)
counts = rng.integers(0, 4, size=(n, F), dtype=np.uint8)
mean = np.concatenate([[0.0], np.full(F, 1.5)])
sd = np.concatenate([[1.0], np.full(F, np.sqrt(1.25))])
cls = (np.arange(n) % C).astype(np.int32); rng.shuffle(cls)
w = np.ones(n)
ann = [f"L{c};L{(c+1)%C}" for c in cls]
des = structure.make_design(ann)
crs = structure.ClassRelationStructure(des.design, des.num_layers)
cls = np.array([des.subgroup_names.index(s) for s in des.subgroups_at_sites], dtype=np.int32)
dim = F + 1
x0 = np.zeros(C * dim); smx0 = np.zeros(crs.numNodes * dim)
lo = opt.LOneCuda(x0, smx0, None)
lo.setCountMatrix(counts, mean, sd)
lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(1)
lo.setClassStructure(crs); lo.setClsMembership(cls); lo.setInstanceWeights(w)
lo.setNumClasses(C); lo.setNumPredictors(F); lo.setPho(1.7); lo.set_numThreads(T)
lo.execute()
s = lo.stats
nb = n // T
Cp = (C + 7) // 8 * 8
flop = 4.0 * nb * Cp * ((dim + 15) // 16 * 16) * s.total_evals
print(name, counts.shape, 'C', C, 'mode', s.eval_mode, 'G', s.ctas_per_block, 'evals', s.total_evals,
      'xupdate %.3fs' % s.seconds_eval, 'per eval round %.1f us' % (s.seconds_eval * 1e6 / (s.total_evals / T)),
      'two-pass algorithmic GB/s %.0f' % (s.eval_bytes / s.seconds_eval / 1e9),
      'DMMA TFLOP/s (padded classes) %.1f' % (flop / s.seconds_eval / 1e12),
      'finite', bool(np.isfinite(lo.getX()).all()))
