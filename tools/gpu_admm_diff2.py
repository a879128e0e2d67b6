import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as O
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
cfg = sys.argv[1]; ns = int(sys.argv[2]); mf = int(sys.argv[3]) if sys.argv[3] != '-' else None
T = int(sys.argv[4])
p = make_problem(cfg, ns, max_features=mf)
n, dim = p.data.shape; C = p.num_classes; NN = p.st.num_nodes
print(cfg, p.data.shape)
# eval parity
rng = np.random.default_rng(0)
order = opt.java_block_order(T)
cx = rng.normal(size=(T, C*dim))*0.05
smx = rng.normal(size=NN*dim)*0.05
f, g = opt.block_evaluate(p.data, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7)
nb = n // T
for b, t in enumerate(order):
    rows = np.arange(nb)*T + t
    zd = np.zeros(NN*NN*dim)
    fo, go = O.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7, p.st, C)
    print('block', t, 'f rel %.2e  g rel %.2e' % (abs(f[b]-fo)/abs(fo), np.abs(g[b]-go).max()/np.abs(go).max()), flush=True)
for A in [1, 2, 5, 20, 60]:
    xo, smo, tr = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, C, num_threads=T, admm_max_itr=A, nodes_max_itr=1, host_threads=8)
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(1)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(C); lo.setNumPredictors(dim-1); lo.setPho(1.7); lo.set_numThreads(T)
    lo.execute()
    xg = lo.getX()
    on = tr.log_nfev[:, order]
    print('A=%d x rel %.3e primal rel %.3e' % (A, np.abs(xg-xo).max()/np.abs(xo).max(), np.abs(tr.log_primal-lo.log.primal).max()/np.abs(tr.log_primal).max()))
    print('   oracle nfev', on[-1], 'gpu nfev', lo.log.nfev[-1], 'all equal', np.array_equal(on, lo.log.nfev))
