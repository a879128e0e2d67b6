"""Feasibility probe: Q optimizers with disjoint CTA budgets on Q host threads (one GPU) against
the same Q trainings run one after the other on the whole GPU. cfg-2, different lambda per run."""
import sys, time, threading, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 144 // Q
A = int(sys.argv[3]) if len(sys.argv) > 3 else 400
S = int(sys.argv[4]) if len(sys.argv) > 4 else 15
p = make_problem('cfg2')
lams = list(np.logspace(-1, 2, 16)[::16 // Q][:Q])

def make(lam, b):
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setRidge(float(lam)); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1); lo.setPho(1.7)
    lo.set_numThreads(8); lo.setCtaBudget(b); lo.setResident(True)
    lo._open()
    return lo

# sequential, whole GPU
t_seq, res_seq = 0.0, []
for lam in lams:
    lo = make(lam, 0)
    t0 = time.time(); lo.execute(); t_seq += time.time() - t0
    res_seq.append((lo.getX().copy(), lo.stats.total_evals, lo.stats.total_admm_iters)); lo.close()
print('sequential: %d runs, %.2f s wall (execute only); evals %s' % (Q, t_seq, [r[1] for r in res_seq]), flush=True)

los = [make(lam, budget) for lam in lams]
ths = [threading.Thread(target=lo.execute) for lo in los]
t0 = time.time()
for t in ths: t.start()
for t in ths: t.join()
t_con = time.time() - t0
print('concurrent: %d runs x %d CTAs, %.2f s wall; device seconds %s; evals %s' % (Q, budget, t_con, ['%.2f' % lo.stats.seconds_total for lo in los], [lo.stats.total_evals for lo in los]), flush=True)
for lo, r, lam in zip(los, res_seq, lams):
    x = lo.getX()
    print('   lambda %.3g: admm iters %d vs %d, max rel diff of weights vs whole-GPU run %.2e' % (lam, lo.stats.total_admm_iters, r[2], np.abs(x - r[0]).max() / np.abs(r[0]).max()))
    lo.close()
print('speed-up %.2fx' % (t_seq / t_con))
