"""Throughput of the streaming evaluation modes at cfg-3 / cfg-4 shapes (synthetic standardised
matrix generated on the fly; a few ADMM iterations; no oracle)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt, structure
name, n, dim, C, T, A = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), 8, 2
rng = np.random.default_rng(0)
X = rng.standard_normal((n, dim)); X[:, 0] = 1.0
cls = (np.arange(n) % C).astype(np.int32); rng.shuffle(cls)
w = np.ones(n)
ann = [f"L{c};L{(c+1)%C}" for c in cls]
des = structure.make_design(ann)
crs = structure.ClassRelationStructure(des.design, des.num_layers)
cls = np.array([des.subgroup_names.index(s) for s in des.subgroups_at_sites], dtype=np.int32)
x0 = np.zeros(C*dim); smx0 = np.zeros(crs.numNodes*dim)
lo = opt.LOneCuda(x0, smx0, X)
lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(1)
lo.setClassStructure(crs); lo.setClsMembership(cls); lo.setInstanceWeights(w)
lo.setNumClasses(C); lo.setNumPredictors(dim-1); lo.setPho(1.7); lo.set_numThreads(T)
lo.execute()
s = lo.stats
nb = n // T
print(name, X.shape, 'C', C, 'evals', s.total_evals, 'xupdate %.3fs' % s.seconds_eval, 'per eval round %.1f us' % (s.seconds_eval*1e6/(s.total_evals/T)), 'two-pass algorithmic GB/s %.0f' % (s.eval_bytes/s.seconds_eval/1e9), 'finite', np.isfinite(lo.getX()).all())
