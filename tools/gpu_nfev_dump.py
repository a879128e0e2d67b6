"""Dump the per-block L-BFGS evaluation counts of every ADMM iteration of a cfg-2 run
(gpurun_out/nfev_cfg2.npy): input of tools/sim_allocation.py."""
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
np.save('gpurun_out/nfev_cfg2.npy', lo.log.nfev)
np.save('gpurun_out/nfev_cfg2_outer.npy', lo.log.outer)
print(lo.log.nfev.shape, lo.stats.seconds_eval)
