import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as O
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
cfg, ns, mf, T, A, S = 'cfg1', 1000, 100, 5, 400, 1
p = make_problem(cfg, ns, max_features=mf)
xo, smo, tr = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(S)
lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1]-1); lo.setPho(1.7); lo.set_numThreads(T)
lo.execute()
order = opt.java_block_order(T)
on = tr.log_nfev[:, order]; gn = lo.log.nfev
m = min(len(on), len(gn))
bad = np.nonzero(~(on[:m] == gn[:m]).all(1))[0]
print('mismatch rows', bad[:10])
r0 = bad[0] if len(bad) else 0
np.set_printoptions(precision=17, linewidth=200)
for r in range(max(0, r0-3), min(m, r0+4)):
    print(r, 'oracle nfev', on[r], 'pho %.6g fold %.4g primal %.17g dual %.17g' % (tr.log_pho[r], tr.log_fold[r], tr.log_primal[r], tr.log_dual[r]))
    print(r, '   gpu nfev', gn[r], 'pho %.6g fold %.4g primal %.17g dual %.17g' % (lo.log.pho[r], lo.log.fold[r], lo.log.primal[r], lo.log.dual[r]))
rel = np.abs(tr.log_primal[:m]-lo.log.primal[:m])/np.abs(tr.log_primal[:m])
print('primal rel err by row (every 20):', rel[:r0+5:20])
