import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
p = make_problem('cfg2', 20000)
lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
lo.setRidge(10.0); lo.setADMMmaxItrs(400); lo.setSeqUnwinderMaxIts(15)
lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1]-1); lo.setPho(1.7); lo.set_numThreads(8)
lo.execute()
nf = lo.log.nfev.astype(float)
print('rows', nf.shape, 'mean evals per block per iteration %.2f' % nf.mean(), 'mean of max %.2f' % nf.max(1).mean(), 'ratio %.3f' % (nf.max(1).mean()/nf.mean()))
print('by outer iteration: ratio', [round(float(nf[lo.log.outer==o].max(1).mean()/nf[lo.log.outer==o].mean()),3) for o in range(lo.stats.outer_iters)])
print('hist of per-row (max-mean):', np.histogram(nf.max(1)-nf.mean(1), bins=[0,1,2,3,4,6,8,12,20,50])[0])
s = lo.stats
print('xupdate total %.3f s; sum over iterations of max evals = %d; per (max) eval %.1f us' % (s.seconds_eval, nf.max(1).sum(), s.seconds_eval*1e6/nf.max(1).sum()))
