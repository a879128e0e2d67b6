import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as O
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem

# 1. rosenbrock
def rosen(x):
    a, b = x[0::2], x[1::2]; t1 = b - a*a
    f = float(np.sum(100*t1*t1 + (1-a)**2)); g = np.zeros_like(x)
    g[0::2] = -400*a*t1 - 2*(1-a); g[1::2] = 200*t1
    return f, g
for n in (2, 10, 256, 2000):
    x0 = np.tile([-1.2, 1.0], n//2)
    for eps in (0.1, 1e-6):
        xr, ifl, nfev, iters = O.lbfgs_minimize(rosen, x0, eps=eps)
        xg, iflg, nfevg = opt.lbfgs_rosenbrock(x0, eps)
        print('rosen n=%d eps=%g oracle nfev=%d iflag=%d | gpu nfev=%d iflag=%d | maxdiff %.3e' % (n, eps, nfev, ifl, nfevg, iflg, np.abs(xr-xg).max()), flush=True)

# 2. eval parity
p = make_problem('cfg1', 500, max_features=60)
T = 5
n, dim = p.data.shape; C = p.num_classes; NN = p.st.num_nodes
rng = np.random.default_rng(0)
order = opt.java_block_order(T)
print('order', order, O.java_thread_order(T))
cx = rng.normal(size=(T, C*dim))*0.1
smx = rng.normal(size=NN*dim)*0.1
f, g = opt.block_evaluate(p.data, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7)
nb = n // T
for b, t in enumerate(order):
    rows = np.arange(nb)*T + t
    zd = np.zeros(NN*NN*dim)
    fo, go = O.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7, p.st, C)
    print('block', t, 'f gpu %.12g oracle %.12g rel %.2e  g rel %.2e' % (f[b], fo, abs(f[b]-fo)/abs(fo), np.abs(g[b]-go).max()/np.abs(go).max()), flush=True)

# 3. full execute parity
for (cfg, ns, mf, T, A, S) in [('cfg1', 500, 60, 5, 50, 3), ('cfg1', 1000, 100, 5, 400, 15)]:
    p = make_problem(cfg, ns, max_features=mf)
    t0 = time.time()
    xo, smo, tr = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    t_or = time.time() - t0
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1]-1); lo.setPho(1.7); lo.set_numThreads(T); lo.initUandZ()
    t0 = time.time(); lo.execute(); t_gpu = time.time() - t0
    xg, smg = lo.getX(), lo.getsmX()
    print(cfg, ns, mf, 'oracle: outer', tr.outer_iters, 'admm', tr.admm_iters, 'evals', tr.total_evals, '%.2fs' % t_or)
    print('   gpu: outer', lo.stats.outer_iters, 'admm', lo.log.admm_iters, 'evals', lo.stats.total_evals, '%.2fs' % t_gpu, 'dev %.3fs admm %.3fs eval %.3fs launches %d' % (lo.stats.seconds_total, lo.stats.seconds_admm, lo.stats.seconds_eval, lo.stats.launches))
    print('   x rel err %.3e  smx rel err %.3e' % (np.abs(xg-xo).max()/np.abs(xo).max(), np.abs(smg-smo).max()/np.abs(smo).max()))
    m = min(len(tr.log_pho), len(lo.log.pho))
    print('   rows', len(tr.log_pho), len(lo.log.pho), 'pho equal', np.array_equal(tr.log_pho[:m], lo.log.pho[:m]), 'primal rel', np.abs(tr.log_primal[:m]-lo.log.primal[:m]).max()/np.abs(tr.log_primal[:m]).max())
    on = tr.log_nfev[:m][:, opt.java_block_order(T)]
    print('   nfev equal rows', (on == lo.log.nfev[:m]).all(1).mean(), 'first mismatch', np.argmax(~(on == lo.log.nfev[:m]).all(1)) if not (on == lo.log.nfev[:m]).all() else -1, flush=True)
