"""BASELINE config 5: 16 log-spaced lambda values on the cfg-2 matrix. `sweep` (one training after
the other on the whole GPU, resident matrix) against `sweep_concurrent` (slots trainings side by
side, 148 // slots CTAs each)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p = make_problem('cfg2')
ridges = list(np.logspace(-1, 2, 16))
def make():
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setADMMmaxItrs(400); lo.setSeqUnwinderMaxIts(15)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1); lo.setPho(1.7)
    lo.set_numThreads(8)
    return lo
t0 = time.time(); rs = make().sweep(ridges); t_seq = time.time() - t0
t0 = time.time(); rc = make().sweep_concurrent(ridges, slots=slots); t_con = time.time() - t0
sites_iters = lambda res: sum(20000 * r[2].total_admm_iters for r in res)
print('sweep (sequential): %.2f s, %.3g site*iter/s' % (t_seq, sites_iters(rs) / t_seq))
print('sweep_concurrent(slots=%d): %.2f s, %.3g site*iter/s -> %.2fx' % (slots, t_con, sites_iters(rc) / t_con, t_seq / t_con))
C, dim = p.num_classes, p.data.shape[1]
from oracle import pyoracle as O
for lam, a, b in zip(ridges, rs, rc):
    pa = O.predict(O.destandardize(a[0], C, dim, p.mean, p.sd), p.raw).argmax(1)
    pb = O.predict(O.destandardize(b[0], C, dim, p.mean, p.sd), p.raw).argmax(1)
    print('   lambda %7.3f: outer %d / %d, ADMM iterations %d / %d, accuracy %.4f / %.4f, same arg-max %.4f' % (
        lam, a[2].outer_iters, b[2].outer_iters, a[2].total_admm_iters, b[2].total_admm_iters,
        (pa == p.cls).mean(), (pb == p.cls).mean(), (pa == pb).mean()))
