import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as O
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
for (cfg, n, T, A) in [('cfg3', 2400, 8, 2), ('cfg4', 3400, 8, 2), ('cfg2', 4000, 8, 3)]:
    p = make_problem(cfg, n)
    C = p.num_classes; dim = p.data.shape[1]
    t0 = time.time()
    xo, smo, tr = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, C, num_threads=T, admm_max_itr=A, nodes_max_itr=1, host_threads=8)
    t_or = time.time() - t0
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(1)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(C); lo.setNumPredictors(dim-1); lo.setPho(1.7); lo.set_numThreads(T)
    t0 = time.time(); lo.execute(); t_gpu = time.time() - t0
    order = opt.java_block_order(T)
    print(cfg, p.data.shape, 'C', C, 'oracle %.1fs gpu %.2fs (dev %.3fs)' % (t_or, t_gpu, lo.stats.seconds_total), 'x rel %.3e' % (np.abs(lo.getX()-xo).max()/np.abs(xo).max()), 'nfev equal', np.array_equal(tr.log_nfev[:, order], lo.log.nfev), 'evals', lo.stats.total_evals, flush=True)
