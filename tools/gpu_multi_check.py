"""torchrun --nproc-per-node 2 tools/gpu_multi_check.py : T fixed, blocks spread over ranks;
results must equal the oracle's single-process run (and hence the 1-GPU run) up to rounding."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as O
from sequnwinder_b200 import optimizer as opt, parallel
from helpers import make_problem

info = parallel.rank_info_from_env()
torch.cuda.set_device(info.local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", info.local_rank))
nid = parallel.share_nccl_id()
for (cfg, n, mf, T, A, S) in [("cfg1", 500, 60, 4, 30, 2), ("cfg2", 4000, None, 8, 4, 1), ("cfg1", 400, 200, 2, 400, 15)]:
    p = make_problem(cfg, n, max_features=mf)
    primal = {}
    # both iteration tails: ONE collective (sums incl. sum x_t^2, residual by expansion; the default)
    # and two collectives (residual of the new z from the local blocks, then a second all-reduce)
    for tail in ("one collective", "two collectives"):
        if tail == "two collectives":
            os.environ["SQW_ADMM_UNFUSED"] = "1"
        else:
            os.environ.pop("SQW_ADMM_UNFUSED", None)
        lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None)
        lo.setCountMatrix(p.counts[:, p.keep].astype(np.uint8), p.mean, p.sd)   # packed-count form
        lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setSeqUnwinderMaxIts(S)
        lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
        lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1); lo.setPho(1.7); lo.set_numThreads(T)
        lo.setComm(info.rank, info.world, nid)
        lo.execute()
        primal[tail] = lo.log.primal.copy()
        x = torch.from_numpy(lo.getX().copy()).cuda()
        xs = [torch.empty_like(x) for _ in range(info.world)]
        dist.all_gather(xs, x)
        same = all(torch.equal(xs[0], t) for t in xs)      # replicated consensus state must be bitwise equal
        if info.rank == 0:
            xo, smo, tr = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
            err = np.abs(lo.getX() - xo).max() / np.abs(xo).max()
            perr = np.abs(lo.log.primal - tr.log_primal).max() / np.abs(tr.log_primal).max()
            print(f"{cfg} n={n} T={T} world={info.world} [{tail}]: ranks bitwise equal {same}; admm iters {lo.log.admm_iters} vs oracle {tr.admm_iters}; x rel err {err:.3e}; primal residual log rel err {perr:.3e}; local blocks {lo.log.nfev.shape[1]}", flush=True)
            assert same and lo.log.admm_iters == tr.admm_iters and err < 1e-6 and perr < 1e-6
    if info.rank == 0:
        # cancellation error of the expanded residual against the direct one
        d = np.abs(primal["one collective"] - primal["two collectives"]).max() / np.abs(primal["two collectives"]).max()
        print(f"   expanded vs direct primal residual: max rel diff {d:.3e}", flush=True)
        assert d < 1e-8
dist.barrier()
dist.destroy_process_group()
