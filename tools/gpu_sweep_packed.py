"""BASELINE config 5 from the packed count matrix: 16 log-spaced lambda values on the cfg-2 counts.
`sweep` (one training after the other on the whole GPU, rows resident in shared memory, MODE 8)
against `sweep_concurrent` (slots trainings side by side on 148 // slots CTAs each: the streamed
packed kernels, MODE 10, read the 13 MB count matrix from L2).

usage: gpu_sweep_packed.py [slots ...] [--no-seq] [--lambdas N]"""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
args = [a for a in sys.argv[1:] if not a.startswith('--')]
slot_list = [int(a) for a in args] or [3]
nlam = 16
if '--lambdas' in sys.argv:
    nlam = int(sys.argv[sys.argv.index('--lambdas') + 1]); slot_list = [s for s in slot_list if s != nlam] or slot_list
p = make_problem('cfg2')
ridges = list(np.logspace(-1, 2, nlam))
counts = (p.counts[:, p.keep].astype(np.uint8), p.mean, p.sd)
def make():
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None)
    lo.setCountMatrix(*counts)
    lo.setADMMmaxItrs(400); lo.setSeqUnwinderMaxIts(15)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1); lo.setPho(1.7)
    lo.set_numThreads(8)
    return lo
sites_iters = lambda res: sum(20000 * r[2].total_admm_iters for r in res)
rs = None
if '--no-seq' not in sys.argv:
    t0 = time.time(); rs = make().sweep(ridges); t_seq = time.time() - t0
    print('sweep (sequential, mode %d): %.2f s, %.3g site*iter/s' % (rs[0][2].eval_mode, t_seq, sites_iters(rs) / t_seq), flush=True)
C, dim = p.num_classes, p.data.shape[1]
from oracle import pyoracle as O
for slots in slot_list:
    t0 = time.time(); rc = make().sweep_concurrent(ridges, slots=slots); t_con = time.time() - t0
    print('sweep_concurrent(slots=%d, mode %d, %d CTAs per block): %.2f s, %.3g site*iter/s%s' % (
        slots, rc[0][2].eval_mode, rc[0][2].ctas_per_block, t_con, sites_iters(rc) / t_con,
        (' -> %.2fx' % (t_seq / t_con)) if rs else ''), flush=True)
    if rs and slots == slot_list[-1]:
        for lam, a, b in zip(ridges, rs, rc):
            pa = O.predict(O.destandardize(a[0], C, dim, p.mean, p.sd), p.raw).argmax(1)
            pb = O.predict(O.destandardize(b[0], C, dim, p.mean, p.sd), p.raw).argmax(1)
            print('   lambda %7.3f: outer %d / %d, ADMM iterations %d / %d, accuracy %.4f / %.4f, same arg-max %.4f' % (
                lam, a[2].outer_iters, b[2].outer_iters, a[2].total_admm_iters, b[2].total_admm_iters,
                (pa == p.cls).mean(), (pb == p.cls).mean(), (pa == pb).mean()))
