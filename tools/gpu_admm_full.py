import sys, time, numpy as np, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as O
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
cfg = sys.argv[1]; run_oracle = int(sys.argv[2]); ns = int(sys.argv[3]) if len(sys.argv) > 3 else None
T = {'cfg1': 5, 'cfg2': 8}[cfg]; A, S = 400, 15
p = make_problem(cfg, ns)
print(cfg, p.data.shape, 'C', p.num_classes, 'cores', os.cpu_count(), flush=True)
packed = os.environ.get('SQW_FP64') is None   # default: the packed-count form of the same matrix
lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None if packed else p.data)
if packed:
    lo.setCountMatrix(p.counts[:, p.keep].astype(np.uint8), p.mean, p.sd)
lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(S)
lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1]-1); lo.setPho(1.7); lo.set_numThreads(T)
if os.environ.get('SQW_BUDGET'):   # a slot of a concurrent sweep: at most this many CTAs
    lo.setCtaBudget(int(os.environ['SQW_BUDGET']))
t0 = time.time(); lo.execute(); t_gpu = time.time() - t0
s = lo.stats
print('mode', s.eval_mode, end=' '); print('gpu: outer', s.outer_iters, 'conv', s.converged, 'admm', lo.log.admm_iters, 'evals', s.total_evals, 'wall %.2fs dev total %.3fs admm %.3fs xupdate %.3fs launches %d pho %.4g' % (t_gpu, s.seconds_total, s.seconds_admm, s.seconds_eval, s.launches, s.final_pho), flush=True)
nb = p.data.shape[0] // T
print('   per-eval-round us: %.2f  site-iter/s %.4g  eval GB/s %.1f' % (s.seconds_eval*1e6 / (s.total_evals / T), nb*T*s.total_admm_iters / s.seconds_admm, s.eval_bytes / s.seconds_eval / 1e9), flush=True)
xg, smg = lo.getX().copy(), lo.getsmX().copy()
if run_oracle:
    t0 = time.time()
    xo, smo, tr = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=T)
    print('oracle: outer', tr.outer_iters, 'conv', tr.converged, 'admm', tr.admm_iters, 'evals', tr.total_evals, '%.2fs (admm %.2fs)' % (time.time()-t0, tr.seconds_admm), flush=True)
    print('   x rel err %.3e  smx rel err %.3e' % (np.abs(xg-xo).max()/np.abs(xo).max(), np.abs(smg-smo).max()/np.abs(smo).max()))
    m = min(len(tr.log_pho), len(lo.log.pho))
    on = tr.log_nfev[:m][:, opt.java_block_order(T)]
    eq = (on == lo.log.nfev[:m]).all(1)
    print('   rows', len(tr.log_pho), len(lo.log.pho), 'pho equal', np.array_equal(tr.log_pho[:m], lo.log.pho[:m]), 'nfev equal frac', eq.mean(), 'first mismatch', (np.argmin(eq) if not eq.all() else -1))
    C = p.num_classes; dim = p.data.shape[1]
    pg = O.predict(O.destandardize(xg, C, dim, p.mean, p.sd), p.raw).argmax(1)
    po = O.predict(O.destandardize(xo, C, dim, p.mean, p.sd), p.raw).argmax(1)
    print('   argmax equal', (pg == po).mean(), 'acc', (pg == p.cls).mean())
