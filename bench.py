#!/usr/bin/env python
"""bench.py — SeqUnwinder's two hot paths on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Default workload = BASELINE.json configs[1] ("README typical", cfg-2): 20,000 x 150 bp sites,
8 overlapping labels, k = 4..5, --r 10, T = 8 ADMM blocks. One step = k-mer featurisation of the
batch (device resident) + one full ADMM training (LOne.execute: all outer iterations to the
reference's tolerances) on the resulting matrix. `value` = ADMM site*iter/s = floor(N/T)*T *
(ADMM iterations executed) / seconds inside the ADMM loops, inputs resident in HBM; `e2e` = the
same through the public host-buffer API (KmerCounter.count + LOneCuda with host buffers, H2D/D2H
inside the timed region).

    --scaling weak    (default) 20,000 sites and 8 ADMM blocks PER GPU (T = 8 N)
    --scaling strong  the SAME 20,000-site, T = 8 problem on N GPUs (8 / N blocks per GPU):
                      identical iteration counts at every N, time-to-converge per N
At N > 1 the line of either mode carries the other mode's result as a sub-object, so one scaling
run records both curves.

    --config cfg3 | cfg4   BASELINE configs[2] / [3] at their stated sizes (200,000 x k4..7, 12
                      labels; 1,999,999 x k4..6, 16 labels + Random out-group): device-resident
                      pipeline sequences -> counts -> column plan -> trainer, a FIXED number of ADMM
                      iterations (--admm-iters), T = 8 blocks, rows sharded over the N GPUs.

Every line carries `parity`: before the timed steps the same optimizer runs ONE ADMM iteration,
the ranks' results are compared bitwise and rank 0 compares them with the CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "admm_site_iter_per_s"
UNIT = "site*iter/s"
KMER_REPS = 10        # k-mer passes enqueued per timed step (see resident_step)
KMER_STREAM_X = 10    # the k-mer kernel alone on a batch this many times cfg-2 (output > L2)
LENGTH = 150
RIDGE, PHO = 10.0, 1.7
T_BLOCKS = 8          # ADMM blocks (--threads) of every named configuration

CONFIGS = {
    # name: sites, labels, k range, out-group, ADMM iteration caps (A per outer iteration, S outer)
    "cfg2": dict(sites=20000, labels=8, kmin=4, kmax=5, outgroup=False, A=400, S=15, seed=2),
    "cfg3": dict(sites=200000, labels=12, kmin=4, kmax=7, outgroup=False, A=None, S=1, seed=3),
    "cfg4": dict(sites=1999999, labels=16, kmin=4, kmax=6, outgroup=True, A=None, S=1, seed=4),
}
# Whole-run statistics of the reference algorithm at cfg-2 (T = 8, to the reference tolerances):
# L-BFGS evaluations per block and ADMM iteration, averaged over the 3 600 iterations of a full run
# (423 000 evaluations / (3 600 x 8); profiles/r02_cfg2_trajectory.json, refreshed by every GPU run).
# The first iterations cost 5x that (the verdict of round 1), so a CPU sample of the first
# iterations is converted to a whole-run rate with this figure, never reported raw.
EVALS_PER_BLOCK_ITER_FALLBACK = 14.69


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def whole_run_evals_per_block_iter():
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_cfg2_trajectory.json")))
        return float(t["evals_per_block_iter"]), "profiles/r02_cfg2_trajectory.json"
    except Exception:
        return EVALS_PER_BLOCK_ITER_FALLBACK, "constant in bench.py (full cfg-2 runs on B200)"


def ncu_traffic_per_xupdate():
    """DRAM bytes of the x-update launches of ONE ADMM iteration (the roofline's launch unit), from
    the committed `ncu --set full` capture of this round's kernel
    (profiles/r02_ncu_xupdate_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep).
    None when the file is absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_xupdate_traffic.json")))
        return float(t["dram_bytes_per_xupdate"]), "profiles/r02_ncu_xupdate_traffic.json: " + t.get("source", "")
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------
# problems
# ------------------------------------------------------------------------------------------------
def problem_shape(cfg: str, world: int, scaling: str):
    """(sites, T) of the job: weak scaling multiplies both by the number of GPUs."""
    c = CONFIGS[cfg]
    mult = world if (scaling == "weak" and cfg == "cfg2") else 1
    return c["sites"] * mult, T_BLOCKS * mult


def workload_string(cfg: str, world: int, scaling: str, admm_iters=None) -> str:
    """Identical in both arms (the driver compares it)."""
    c = CONFIGS[cfg]
    sites, T = problem_shape(cfg, world, scaling)
    numK = sum(4 ** k for k in range(c["kmin"], c["kmax"] + 1))
    leaves = c["labels"] + (1 if c["outgroup"] else 0)
    run = (f"A={c['A']}, S={c['S']} to the reference tolerances" if c["A"] is not None
           else f"{admm_iters} ADMM iterations of outer iteration 1 (fixed count)")
    return (f"{cfg}: {sites}x{LENGTH}bp sites, {c['labels']} ring labels"
            f"{' + Random out-group' if c['outgroup'] else ''} (C={leaves}), k={c['kmin']}..{c['kmax']} "
            f"(numK={numK}), r={RIDGE}, pho={PHO}, T={T} ADMM blocks, {run}")


def build_problem(cfg: str, sites: int):
    """Synthetic sequences, labels, design. Host only."""
    from sequnwinder_b200 import structure, synth

    c = CONFIGS[cfg]
    d = synth.generate(sites, c["labels"], length=LENGTH, seed=synth.SEED_BASE + c["seed"],
                       outgroup=c["outgroup"])
    des = structure.make_design(d.annotations)
    crs = structure.ClassRelationStructure(des.design, des.num_layers)
    index = {s: i for i, s in enumerate(des.subgroup_names)}
    cls = np.fromiter((index[s] for s in des.subgroups_at_sites), dtype=np.int32, count=d.n_sites)
    wmap = np.array([des.subgroup_weights[s] for s in des.subgroup_names])
    w = wmap[cls].astype(np.float64)
    return d, crs, cls, w, len(des.subgroup_names)


def make_optimizer(prep, crs, cls, w, C, T, A, S, comm=None, counts=None):
    from sequnwinder_b200.optimizer import LOneCuda

    if counts is not None:   # packed-count form: uint8 counts of the kept k-mers + xMean / xSD
        lo = LOneCuda(prep.x0.copy(), prep.sm_x0.copy(), None)
        lo.setCountMatrix(counts, prep.mean, prep.sd)
    else:
        lo = LOneCuda(prep.x0.copy(), prep.sm_x0.copy(), prep.data)
    lo.setRidge(RIDGE)
    lo.setADMMmaxItrs(A)
    lo.setBGFSmaxItrs(-1)
    lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(crs)
    lo.setClsMembership(cls)
    lo.setInstanceWeights(w)
    lo.setNumClasses(C)
    lo.setNumPredictors(prep.keep.size)
    lo.setPho(PHO)
    lo.set_numThreads(T)
    lo.initUandZ()
    lo.setDebugMode(False)
    if comm is not None:
        lo.setComm(*comm)
    return lo


def standardised(prep, counts):
    """The fp64 training matrix of Classifier.java:373-381 (the CPU legs take it; the device takes
    the counts)."""
    if prep.data is not None:
        return prep.data
    data = np.empty((counts.shape[0], prep.keep.size + 1))
    data[:, 0] = 1.0
    data[:, 1:] = counts[:, prep.keep]
    nz = prep.sd != 0
    data[:, nz] = (data[:, nz] - prep.mean[nz]) / prep.sd[nz]
    prep.data = data
    return data


def initial_weights(crs, cls, C, dim):
    """x0 / sm_x0 of Classifier.java:383-420 from the unweighted class counts."""
    sY = np.bincount(cls, minlength=C).astype(np.float64)
    NN = crs.numNodes
    sm_sY = np.zeros(NN)
    for leaf in crs.leafs:
        sm_sY[leaf.nodeIndex] = sY[leaf.nodeIndex]
    for l in range(1, crs.numLayers):
        for node in crs.layers[l]:
            for ch in node.children:
                sm_sY[node.nodeIndex] += sm_sY[ch]
    x0 = np.zeros(C * dim)
    x0[0::dim] = np.log(sY + 1.0) - np.log(sY[C - 1] + 1.0)
    smx0 = np.zeros(NN * dim)
    smx0[0::dim] = np.log(sm_sY + 1.0) - np.log(sm_sY[C - 1] + 1.0)
    return x0, smx0


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle/ = test infrastructure: the checker and the reported CPU baseline, never the
# product path)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_sample(data, x0, smx0, cls, w, crs, C, T, admm_iters, threads):
    """The oracle (CPU restatement of LOne.java) on a bounded sample: the first `admm_iters` ADMM
    iterations of outer iteration 1, T blocks on `threads` host threads. Returns seconds inside
    the ADMM loop, iterations, block evaluations."""
    from oracle import pyoracle as O

    st = O.StructureArrays(**crs.to_arrays())
    _, _, tr = O.lone_execute(x0, smx0, data, cls, w, st, C, ridge=RIDGE, pho=PHO, num_threads=T,
                              admm_max_itr=admm_iters, nodes_max_itr=1, host_threads=threads)
    return tr.seconds_admm, sum(tr.admm_iters), int(tr.total_evals)


def cpu_whole_run_rate(sites_used, T, sec, iters, evals):
    """Convert a sample of the first ADMM iterations into a whole-run site*iter/s: the CPU's
    block-evaluation rate (what the sample measures; an evaluation costs the same everywhere on the
    trajectory) divided by the whole run's evaluations per ADMM iteration. The raw sample rate would
    under-state the CPU 5x, because the first iterations need 5x the evaluations of the average one."""
    epbi, src = whole_run_evals_per_block_iter()
    evals_per_s = evals / sec
    value = sites_used * evals_per_s / (epbi * T)
    return value, evals_per_s, {"raw_sample_site_iter_per_s": sites_used * iters / sec,
                                "sample_evals_per_admm_iter": evals / max(1, iters),
                                "whole_run_evals_per_admm_iter": epbi * T, "whole_run_source": src}


def oracle_kmer_all_cores(O, d, kmin, kmax, threads):
    """The oracle's counting loop on `threads` host threads (contiguous row ranges; ctypes drops the
    GIL during the call): SURVEY.md 8d asks for the path-1 CPU number single-threaded and on all cores."""
    n = d.n_sites
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)

    def work(a, b):
        if b > a:
            b0 = int(d.offsets[a])
            O.kmer_count(d.bases[b0:int(d.offsets[b])], d.offsets[a:b + 1] - b0, kmin, kmax)

    ths = [threading.Thread(target=work, args=(int(bounds[i]), int(bounds[i + 1]))) for i in range(threads)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return n * LENGTH / (time.time() - t0) / 1e9


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path. The Java reference cannot
    be built in this image (no JDK; Weka and seqcode-core not vendored), so this is the oracle
    port on all host threads it can use: one per ADMM block (LOne.java:599-606 — T = --threads is
    the reference's own limit), at most the box's cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as O
    from sequnwinder_b200 import classifier as clf

    O.build()
    cfg = args.config
    c = CONFIGS[cfg]
    world = max(1, args.gpus)
    sites, T = problem_shape(cfg, world, args.scaling)
    cores = os.cpu_count() or 1
    threads = min(T, cores)
    admm_iters = args.admm_iters if c["A"] is None else None
    # bounded sample: the reference's cost per evaluation is linear in the rows, so the large
    # configurations are sampled on their first rows (same columns, same design)
    sample_sites = sites if cfg == "cfg2" else min(sites, 16000 if cfg == "cfg3" else 64000)
    d, crs, cls, w, C = build_problem(cfg, sample_sites)
    counts, _ = O.kmer_count(d.bases, d.offsets, c["kmin"], c["kmax"])
    prep = clf.prepare(counts, cls, w, crs, C)
    sample_iters = 3 if cfg == "cfg2" else 1
    vals, secs, eps = [], [], []
    extra = {}
    for i in range(args.warmup + args.steps):
        s, its, ev = cpu_oracle_sample(prep.data, prep.x0, prep.sm_x0, cls, w, crs, C, T, sample_iters, threads)
        if cfg == "cfg2":
            v, e, extra = cpu_whole_run_rate((sample_sites // T) * T, T, s, its, ev)
        else:
            # fixed-count runs: seconds scaled to the full row count (the cost of an evaluation is
            # linear in the rows), rate over the sample's iterations
            s_full = s * (sites / sample_sites)
            v, e = (sites // T) * T * its / s_full, ev / s_full
        if i >= args.warmup:
            vals.append(v)
            secs.append(s)
            eps.append(e)
    t0 = time.time()
    O.kmer_count(d.bases, d.offsets, c["kmin"], c["kmax"])
    kmer_gbs = d.n_sites * LENGTH / (time.time() - t0) / 1e9
    kmer_gbs_all = oracle_kmer_all_cores(O, d, c["kmin"], c["kmax"], cores)
    value = float(np.mean(vals))
    if cfg == "cfg2":
        sample = (f"first {sample_iters} ADMM iterations of outer iteration 1 on the full "
                  f"{sample_sites}x{prep.keep.size + 1} matrix, T={T} blocks on {threads} pthreads; value = "
                  f"sites x measured block evaluations/s / whole-run evaluations per ADMM iteration")
    else:
        sample = (f"first ADMM iteration on the first {sample_sites} of {sites} sites "
                  f"({prep.keep.size + 1} columns), T={T} blocks on {threads} pthreads, seconds scaled "
                  f"by {sites / sample_sites:.2f} (cost linear in rows)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_string(cfg, world, args.scaling, admm_iters), "sample": sample},
        "block_evals_per_s": float(np.mean(eps)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample, "host_cores": cores, **extra,
                         "note": "CPU restatement of LOne.java (oracle/, gcc -O2), not the JVM; the "
                                 "reference runs one worker thread per ADMM block (T = --threads), so "
                                 "cores = min(T, host cores); the Java reference additionally sleeps "
                                 ">= 1 s per ADMM run (LOne.java:728-740)"},
        "kmer_gbases_per_s": kmer_gbs,
        "kmer_gbases_per_s_all_cores": {"value": kmer_gbs_all, "cores": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# device arm, cfg-2
# ------------------------------------------------------------------------------------------------
class Cfg2Job:
    """One cfg-2 job (sites, T) on `world` ranks: data, counts, column plan, optimizer factory."""

    def __init__(self, sites, T, info, comm, dev, kc):
        import torch
        from sequnwinder_b200 import classifier as clf

        self.sites, self.T, self.info, self.comm, self.dev, self.kc = sites, T, info, comm, dev, kc
        self.d, self.crs, self.cls, self.w, self.C = build_problem("cfg2", sites)
        self.d_bases = torch.from_numpy(self.d.bases).to(dev)
        self.d_offs = torch.from_numpy(self.d.offsets).to(dev)
        self.counts_t, self.has_n_t, self.status_t = kc.count_device(self.d_bases, self.d_offs, LENGTH)
        torch.cuda.synchronize()
        self.counts = self.counts_t.cpu().numpy()
        self.prep = clf.prepare(self.counts, self.cls, self.w, self.crs, self.C, materialize=False)
        self.cnt8 = (np.ascontiguousarray(self.counts[:, self.prep.keep]).astype(np.uint8)
                     if os.environ.get("SQW_BENCH_FP64") is None else None)
        if self.cnt8 is None:
            standardised(self.prep, self.counts)
        self.dim = self.prep.keep.size + 1
        self.nb = sites // T

    def optimizer(self, A, S):
        c = CONFIGS["cfg2"]
        return make_optimizer(self.prep, self.crs, self.cls, self.w, self.C, self.T,
                              c["A"] if A is None else A, c["S"] if S is None else S, self.comm, self.cnt8)

    def parity(self, world, dist):
        """ONE ADMM iteration through the same optimizer the timed steps use: ranks bitwise equal,
        rank 0 vs the oracle (weights <= 1e-6 relative, identical evaluation counts), k-mer counts
        of a row sample bit-exact."""
        import torch
        from sequnwinder_b200.optimizer import java_block_order

        lo = self.optimizer(1, 1)
        lo.execute()
        x, smx = lo.getX().copy(), lo.getsmX().copy()
        nfev_local = lo.log.nfev.copy()
        out = {"admm_iters": 1, "eval_mode": int(lo.stats.eval_mode)}
        order = [t_ for t_ in java_block_order(self.T) if t_ % world == self.info.rank]
        nfev_by_block = np.zeros(self.T, dtype=np.int32)
        nfev_by_block[order] = nfev_local[0]
        if world > 1:
            t = torch.from_numpy(np.concatenate([x, smx])).to(self.dev)
            allx = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allx, t)
            out["ranks_bitwise_equal"] = bool(all(torch.equal(allx[0], a) for a in allx[1:]))
            nf = torch.from_numpy(nfev_by_block).to(self.dev)   # every block has exactly one owner
            dist.all_reduce(nf)
            nfev_by_block = nf.cpu().numpy()
        if self.info.rank == 0:
            from oracle import pyoracle as O
            O.build()
            data = standardised(self.prep, self.counts)
            st = O.StructureArrays(**self.crs.to_arrays())
            cores = os.cpu_count() or 1
            xo, smo, tr = O.lone_execute(self.prep.x0, self.prep.sm_x0, data, self.cls, self.w, st, self.C,
                                         ridge=RIDGE, pho=PHO, num_threads=self.T, admm_max_itr=1,
                                         nodes_max_itr=1, host_threads=min(self.T, cores))
            out["rel_err_x"] = float(np.abs(x - xo).max() / np.abs(xo).max())
            out["rel_err_sm_x"] = float(np.abs(smx - smo).max() / np.abs(smo).max())
            out["nfev_identical"] = bool(np.array_equal(nfev_by_block, tr.log_nfev[0]))
            rows = min(2000, self.sites)
            want, _ = O.kmer_count(self.d.bases[: rows * LENGTH], self.d.offsets[: rows + 1],
                                   self.kc.kmin, self.kc.kmax)
            out["kmer_counts_bit_exact"] = bool(np.array_equal(want, self.counts[:rows]))
            out["ok"] = bool(out["rel_err_x"] <= 1e-6 and out["rel_err_sm_x"] <= 1e-6 and
                             out["nfev_identical"] and out["kmer_counts_bit_exact"] and
                             out.get("ranks_bitwise_equal", True))
            out["checked_against"] = "CPU oracle (oracle/lone_oracle.c), tolerance 1e-6 relative"
        return out


def timed_training(job, args, world, dist, torch, with_kmer=True):
    """W warm-up + K timed steps of the resident pipeline of `job`; device times max over ranks."""
    lo = job.optimizer(None, None)
    lo.setResident(True)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        # the k-mer pass is 25 us of GPU time, less than the host needs to launch it: it is
        # enqueued KMER_REPS times between the two events (same inputs, same output buffer) so that
        # the event difference is GPU time and not launch latency; the time of ONE pass is what
        # enters the step
        ks = 0.0
        if with_kmer:
            ev0.record()
            for _ in range(KMER_REPS):
                job.kc.count_device(job.d_bases, job.d_offs, LENGTH, out=job.counts_t, has_n=job.has_n_t,
                                    status=job.status_t)
            ev1.record()
        np.copyto(lo.x, job.prep.x0)
        np.copyto(lo.sm_x, job.prep.sm_x0)
        lo.execute()
        torch.cuda.synchronize()
        if with_kmer:
            ks = ev0.elapsed_time(ev1) * 1e-3 / KMER_REPS
        return ks, lo.stats, lo.log

    for _ in range(args.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(job.info.local_rank)
    if job.info.rank == 0:
        sampler.start()
    t_wall0 = time.time()
    acc = dict(kmer_s=0.0, admm_s=0.0, total_s=0.0, eval_s=0.0, iters=0, evals=0, launches=0,
               eval_bytes=0, outer=[], admm_iters=[])
    for _ in range(args.steps):
        ks, st, log = resident_step()
        acc["kmer_s"] += ks
        acc["admm_s"] += st.seconds_admm
        acc["total_s"] += st.seconds_total + ks
        acc["eval_s"] += st.seconds_eval
        acc["iters"] += st.total_admm_iters
        acc["evals"] += st.total_evals
        acc["eval_bytes"] += st.eval_bytes
        acc["launches"] += st.launches + (KMER_REPS if with_kmer else 0)
        acc["outer"].append(int(st.outer_iters))
        acc["admm_iters"].append([int(v) for v in log.admm_iters])
    barrier()
    acc["wall"] = time.time() - t_wall0
    acc["clocks"] = sampler.stop() if job.info.rank == 0 else None
    acc["stats_mode"] = int(lo.stats.eval_mode)
    acc["ctas_per_block"] = int(lo.stats.ctas_per_block)
    # max over ranks of the device-timed seconds; evaluations summed over ranks
    tm = torch.tensor([acc["admm_s"], acc["total_s"], acc["kmer_s"], acc["eval_s"]], dtype=torch.float64,
                      device=job.dev)
    ev = torch.tensor([acc["evals"], acc["eval_bytes"], acc["launches"]], dtype=torch.float64, device=job.dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    acc["admm_s"], acc["total_s"], acc["kmer_s"], acc["eval_s_max"] = [float(v) for v in tm.cpu()]
    acc["evals_all"], acc["eval_bytes_all"], acc["launches_all"] = [float(v) for v in ev.cpu()]
    lo.close()
    return acc


def run_cfg2(args, info, world, dev, comm, dist, torch):
    from sequnwinder_b200 import KmerCounter

    c = CONFIGS["cfg2"]
    kc = KmerCounter(c["kmin"], c["kmax"])
    sites, T = problem_shape("cfg2", world, args.scaling)
    job = Cfg2Job(sites, T, info, comm, dev, kc)
    parity = None if args.no_parity else job.parity(world, dist)
    acc = timed_training(job, args, world, dist, torch)
    nb = job.nb
    value = nb * T * acc["iters"] / acc["admm_s"]
    kmer_gbases = sites * LENGTH * args.steps / acc["kmer_s"] / 1e9

    # the other scaling mode on the same ranks (N > 1 only): one scaling run records both curves
    other = None
    if world > 1 and not args.no_other_scaling:
        omode = "strong" if args.scaling == "weak" else "weak"
        osites, oT = problem_shape("cfg2", world, omode)
        ojob = Cfg2Job(osites, oT, info, comm, dev, kc)
        oacc = timed_training(ojob, args, world, dist, torch, with_kmer=False)
        other = {"scaling": omode, "sites": osites, "admm_blocks": oT,
                 "value": ojob.nb * oT * oacc["iters"] / oacc["admm_s"], "unit": UNIT,
                 "time_to_converge_s": oacc["admm_s"] / args.steps,
                 "admm_iters_per_step": oacc["iters"] / args.steps, "outer_iters": oacc["outer"],
                 "admm_iters_per_outer": oacc["admm_iters"][-1],
                 "evals_per_step": oacc["evals_all"] / args.steps}
        del ojob

    # ---------------- the k-mer kernel alone on a batch whose output exceeds L2
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)

    def kmer_stream():
        per = CONFIGS["cfg2"]["sites"]
        big_bases = job.d_bases[: per * LENGTH].repeat(KMER_STREAM_X)
        n_big = per * KMER_STREAM_X
        big_offs = torch.arange(n_big + 1, dtype=torch.int64, device=dev) * LENGTH
        out = torch.empty((n_big, kc.numK), dtype=torch.int32, device=dev)
        hn = torch.empty(n_big, dtype=torch.uint8, device=dev)
        stt = torch.zeros(1, dtype=torch.int32, device=dev)
        kc.count_device(big_bases, big_offs, LENGTH, out=out, has_n=hn, status=stt)
        torch.cuda.synchronize()
        reps = 3
        ev0.record()
        for _ in range(reps):
            kc.count_device(big_bases, big_offs, LENGTH, out=out, has_n=hn, status=stt)
        ev1.record()
        torch.cuda.synchronize()
        sec = ev0.elapsed_time(ev1) * 1e-3 / reps
        ok = bool((out[:per] == job.counts_t[:per]).all().item()) and int(stt.item()) == 0
        return n_big, sec, ok

    ks_n, ks_sec, ks_ok = kmer_stream() if info.rank == 0 else (0, 1.0, True)

    # ---------------- e2e through the public host-buffer API (H2D/D2H inside the timed region)
    # path 1 shards by rows with no exchange: every rank featurises its own rows
    d = job.d
    r_a, r_b = sites * info.rank // world, sites * (info.rank + 1) // world
    b_a, b_b = int(d.offsets[r_a]), int(d.offsets[r_b])
    my_bases = np.ascontiguousarray(d.bases[b_a:b_b])
    my_offs = np.ascontiguousarray(d.offsets[r_a:r_b + 1] - b_a)
    kmer_host_s = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the caller's count matrix is allocated once and re-used by every step, as the Java side keeps
    # its int[n][numK]: a fresh 102 MB array per call costs 2 ms of first-touch page faults
    # (gpurun_out/r2_kmerhost24.log: 3.6 against 1.6 ms per call)
    host_counts = np.empty((my_offs.size - 1, kc.numK), dtype=np.int32)

    def e2e_step():
        t0 = time.time()
        cnt, hn = kc.count(my_bases, my_offs, out=host_counts)
        t1 = time.time()
        lo2 = job.optimizer(None, None)
        lo2._open()          # handle + problem upload (execute() would do it)
        t2 = time.time()
        lo2.execute()
        _ = float(lo2.getX()[0])
        dt = time.time() - t0
        print(f"[bench] rank {info.rank} e2e step: k-mer (host buffers) {t1 - t0:.3f} s, upload "
              f"{t2 - t1:.3f} s, execute {dt - (t2 - t0):.3f} s (device {lo2.stats.seconds_total:.3f} s)",
              file=sys.stderr)
        kmer_host_s.append(t1 - t0)
        return dt, lo2.stats.total_admm_iters, lo2.h2d_bytes, (cnt.nbytes + hn.nbytes) * world

    barrier()
    e2e_dt, e2e_iters, h2d, d2h = 0.0, 0, 0, 0
    e2e_steps = max(1, min(args.steps, 2))
    if args.warmup > 0:
        e2e_step()  # untimed: the host-buffer path's one-time allocations (streams, staging buffers)
    barrier()
    kmer_host_s.clear()
    for _ in range(e2e_steps):
        dt, it, hb, db = e2e_step()
        e2e_dt += dt
        e2e_iters += it
        h2d, d2h = hb, db
    barrier()
    # the host-buffer k-mer call on its own: five back-to-back calls into the caller's matrix (inside
    # an e2e step the call follows a 3 s training, with cold caches and descheduled host threads)
    host_call_s = []
    for _ in range(5):
        t0 = time.perf_counter()
        kc.count(my_bases, my_offs, out=host_counts)
        host_call_s.append(time.perf_counter() - t0)
    host_call_s.sort()
    barrier()
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nb * T * e2e_iters / float(te.item())
    # whole-job bytes: the k-mer shards of all ranks, the count matrix as bytes (every rank uploads
    # it and keeps the rows of its own blocks), x / sm_x both ways
    h2d += (my_bases.nbytes + my_offs.nbytes) * world + job.prep.x0.nbytes + job.prep.sm_x0.nbytes
    d2h += job.prep.x0.nbytes + job.prep.sm_x0.nbytes
    if info.rank != 0:
        return None

    peak, peak_src = measured_peak_hbm()
    # per GPU: the algorithmic bytes of one rank's evaluations over that rank's summed x-update
    # durations (max over ranks)
    achieved = acc["eval_bytes_all"] / world / acc["eval_s_max"] / 1e9 if acc["eval_s_max"] > 0 else 0.0
    traffic, traffic_src = ncu_traffic_per_xupdate()
    if world != 1 or acc["stats_mode"] not in (8, 9):
        traffic, traffic_src = None, None   # the capture is of the single-GPU packed-count kernel
    kmer_bytes = sites * (LENGTH + 4 * kc.numK) * args.steps
    block_evals_per_s = acc["evals_all"] / acc["admm_s"]
    # whole-run trajectory statistics (the CPU arm converts its sample with them)
    traj = {"workload": workload_string("cfg2", world, args.scaling), "admm_iters": acc["iters"] / args.steps,
            "block_evals": acc["evals_all"] / args.steps, "blocks": T,
            "evals_per_block_iter": acc["evals_all"] / max(1, acc["iters"]) / T}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(traj, open(os.path.join(ROOT, "gpurun_out", f"cfg2_trajectory_n{world}_{args.scaling}.json"), "w"))
    except Exception:
        pass
    packed = acc["stats_mode"] in (8, 9)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": acc["total_s"] / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_string("cfg2", world, args.scaling),
            "sites": sites, "admm_blocks": T, "rows_per_block": nb, "blocks_per_gpu": T // world,
            "representation": ("packed k-mer counts (uint8) + xMean / xSD, X = (n - mean)/sd evaluated in "
                               "fp64 without materialising it (sqw_admm_set_problem_counts)") if packed
                              else "standardised fp64 matrix",
            "l2": ("the CTA's rows of the count matrix (13 MB per 20,000 sites) stay in shared memory for "
                   "the whole launch; nothing is re-read from L2 or HBM between evaluations, so there is "
                   "no cache to flush between timed iterations") if packed else
                  "training matrix re-read every evaluation (104 MB per GPU, HBM speed)",
            "step": "k-mer featurisation (resident) + LOne.execute to the reference tolerances",
        },
        "parity": parity,
        "time_to_converge_s": acc["admm_s"] / args.steps,
        "admm_iters_per_step": acc["iters"] / args.steps,
        "outer_iters": acc["outer"],
        "admm_iters_per_outer": acc["admm_iters"][-1],
        "evals_per_step": acc["evals_all"] / args.steps,
        "block_evals_per_s": block_evals_per_s,
        "kmer_gbases_per_s": kmer_gbases,
        "kmer_roofline": {"bound": "hbm", "achieved": kmer_bytes / acc["kmer_s"] / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": kmer_bytes / acc["kmer_s"] / 1e9 / peak,
                          "note": "cfg-2 batch (102 MB of counts per pass): launch/latency bound"},
        # the same featurisation through the host-buffer call (H2D of the bases, D2H
        # of the count matrix inside the timing), rank 0's rows
        "kmer_host_api_gbases_per_s": (r_b - r_a) * LENGTH / host_call_s[len(host_call_s) // 2] / 1e9,
        "kmer_host_api": {"ms_per_call_median_of_5": host_call_s[len(host_call_s) // 2] * 1e3,
                          "ms_per_call_inside_e2e_steps": [t * 1e3 for t in kmer_host_s],
                          "note": "sqw_kmer_count with host buffers (bases H2D, dense int32 counts D2H into a "
                                  "matrix the caller re-uses), rank 0's rows"},
        "kmer_stream": {"sites": ks_n, "gbases_per_s": ks_n * LENGTH / ks_sec / 1e9,
                        "bound": "hbm", "achieved": ks_n * (LENGTH + 4 * kc.numK) / ks_sec / 1e9,
                        "peak": peak, "unit": "GB/s",
                        "frac": ks_n * (LENGTH + 4 * kc.numK) / ks_sec / 1e9 / peak,
                        "ms_per_launch": ks_sec * 1e3, "matches_cfg2_counts": ks_ok,
                        "note": f"kmer_count kernel alone, {KMER_STREAM_X} x the cfg-2 sequences in "
                                "one launch (output larger than L2), CUDA events"},
        "roofline": {"bound": "hbm", "kernel": "admm_xupdate_kernel", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": acc["eval_bytes_all"] / world / max(1, acc["iters"]),
                     "peak_source": peak_src,
                     "launch_unit": "one ADMM iteration's x-update = the phased cooperative launches of "
                                    "admm_xupdate_kernel; achieved = algorithmic bytes / summed CUDA-event "
                                    "durations of those launches (max over ranks)",
                     "note": "algorithmic bytes = evaluations x (2 n_b dim 8 + 2 n_b C 8 + 12 n_b + 2 C dim 8), "
                             "SURVEY 8d's two-pass fp64 figure, over the summed durations of the x-update "
                             "kernels (which also contain the L-BFGS phases). With the packed-count "
                             "representation the kernel moves almost none of these bytes (traffic): the "
                             "fraction compares the kernel with an HBM-bound fp64 implementation; what "
                             "bounds it is the fp64 pipe (DMMA + softmax) and the solver's exchanges "
                             "(DESIGN.md 2.3)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "seconds_per_step": float(te.item()) / e2e_steps},
        "gpu_launches": int(acc["launches_all"]),
        "clocks": acc["clocks"],
        "wall_s": acc["wall"],
    }
    if other is not None:
        line[other["scaling"] + "_scaling"] = other
    if world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as O
        O.build()
        cores = os.cpu_count() or 1
        threads = min(T, cores)
        data = standardised(job.prep, job.counts)
        s, its, ev = cpu_oracle_sample(data, job.prep.x0, job.prep.sm_x0, job.cls, job.w, job.crs, job.C,
                                       T, 3, threads)
        v, eps, extra = cpu_whole_run_rate(nb * T, T, s, its, ev)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": threads, "kind": "port", "host_cores": cores,
            "block_evals_per_s": eps, **extra,
            "sample": f"first {its} ADMM iterations of outer iteration 1, full cfg2 matrix, T={T} blocks on "
                      f"{threads} pthreads ({s:.1f} s); value = sites x block evaluations/s / whole-run "
                      f"evaluations per ADMM iteration"}
        # the reference's default --threads 5 as well: same matrix, 5 blocks
        s5, its5, ev5 = cpu_oracle_sample(data, job.prep.x0, job.prep.sm_x0, job.cls, job.w, job.crs, job.C,
                                          5, 2, min(5, cores))
        line["cpu_baseline"]["threads5"] = {
            "block_evals_per_s": ev5 / s5, "cores": min(5, cores),
            "sample": f"first {its5} ADMM iterations, T=5 blocks on {min(5, cores)} pthreads ({s5:.1f} s)"}
    return line


# ------------------------------------------------------------------------------------------------
# device arm, cfg-3 / cfg-4 at full size: device-resident pipeline, rows sharded over the ranks
# ------------------------------------------------------------------------------------------------
def run_big(args, info, world, dev, comm, dist, torch):
    from sequnwinder_b200 import KmerCounter, parallel
    from sequnwinder_b200 import classifier as clf
    from sequnwinder_b200.optimizer import LOneCuda

    cfg = args.config
    c = CONFIGS[cfg]
    sites, T = problem_shape(cfg, world, "strong")
    A = args.admm_iters
    t_setup0 = time.time()
    d, crs, cls, w, C = build_problem(cfg, sites)
    nb = sites // T
    # this rank's rows: the rows of its blocks (t = rank mod world), ascending (path 1 shards by rows)
    mine = parallel.local_row_indices(sites, T, info.rank, world)
    kc = KmerCounter(c["kmin"], c["kmax"])
    all_seq = d.bases.reshape(sites, LENGTH)
    d_bases = torch.from_numpy(np.ascontiguousarray(all_seq[mine]).reshape(-1)).to(dev)
    d_offs = torch.arange(mine.size + 1, dtype=torch.int64, device=dev) * LENGTH
    y_t = torch.from_numpy(cls[mine]).to(dev)
    w_t = torch.from_numpy(w[mine]).to(dev)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup0
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev2 = torch.cuda.Event(enable_timing=True)

    # ---- path 1 + the column plan, device resident, timed with CUDA events (the output buffer is
    # allocated and the kernel warmed up first: cudaMalloc of GBs would otherwise sit in the interval)
    counts_t, has_n_t, status_t = kc.count_device(d_bases, d_offs, LENGTH)
    torch.cuda.synchronize()
    ev0.record()
    kc.count_device(d_bases, d_offs, LENGTH, out=counts_t, has_n=has_n_t, status=status_t)
    ev1.record()
    # the rows the blocking drops (i >= floor(N/T) T) still count for the statistics in the
    # reference (Classifier.java:337-372 runs over all instances): rank 0 adds them
    extra = np.arange(nb * T, sites)
    more = None
    if extra.size and info.rank == 0:
        eb = torch.from_numpy(np.ascontiguousarray(all_seq[extra]).reshape(-1)).to(dev)
        eo = torch.arange(extra.size + 1, dtype=torch.int64, device=dev) * LENGTH
        ecounts, _, _ = kc.count_device(eb, eo, LENGTH)
        more = (ecounts, torch.from_numpy(w[extra]).to(dev))
    plan = clf.plan_device(counts_t, w_t, group=True if world > 1 else None, extra=more)
    ev2.record()
    torch.cuda.synchronize()
    kmer_s = ev0.elapsed_time(ev1) * 1e-3
    plan_s = ev1.elapsed_time(ev2) * 1e-3
    assert int(status_t.item()) == 0 and int(has_n_t.sum().item()) == 0
    F = plan.num_kept
    dim = F + 1
    x0, smx0 = initial_weights(crs, cls, C, dim)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the NCCL communicator is created once per process and shared by every trainer handle
    # (sqw_admm_set_comm caches it by ncclUniqueId): time it apart from the per-training stages
    comm_s = 0.0
    if comm is not None:
        from sequnwinder_b200 import _lib as _l
        import ctypes as _C
        t0 = time.time()
        hh = _C.c_void_p()
        _l.check(_l.load().sqw_admm_create(_C.byref(hh)))
        _l.check(_l.load().sqw_admm_set_comm(hh, comm[0], comm[1], (_C.c_uint8 * 128).from_buffer_copy(comm[2])))
        _l.load().sqw_admm_destroy(hh)
        torch.cuda.synchronize()
        comm_s = time.time() - t0

    lo = LOneCuda(x0.copy(), smx0.copy(), None)
    lo.setCountMatrixDevice(counts_t, plan.keep_idx, plan.mean, plan.sd, y_t, w_t, sites, rows_local=True)
    lo.setRidge(RIDGE); lo.setADMMmaxItrs(A); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(1)
    lo.setClassStructure(crs); lo.setClsMembership(cls[:1]); lo.setInstanceWeights(w[:1])
    lo.setNumClasses(C); lo.setNumPredictors(F); lo.setPho(PHO); lo.set_numThreads(T)
    lo.initUandZ(); lo.setDebugMode(False)
    if comm is not None:
        lo.setComm(*comm)
    lo.setResident(True)
    t0 = time.time()
    lo._open()            # counts -> block-major training matrix on the device
    torch.cuda.synchronize()
    build_s = time.time() - t0
    # the int32 count matrix is no longer needed once the trainer holds its own copy
    counts_bytes = counts_t.numel() * 4
    lo._counts_dev = None
    del counts_t
    torch.cuda.empty_cache()

    def step():
        np.copyto(lo.x, x0)
        np.copyto(lo.sm_x, smx0)
        lo.execute()
        return lo.stats

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(info.local_rank)
    if info.rank == 0:
        sampler.start()
    admm_s = eval_s = total_s = 0.0
    iters = evals = launches = eval_bytes = 0
    for _ in range(args.steps):
        st = step()
        admm_s += st.seconds_admm
        eval_s += st.seconds_eval
        total_s += st.seconds_total
        iters += st.total_admm_iters
        evals += st.total_evals
        launches += st.launches
        eval_bytes += st.eval_bytes
    barrier()
    clocks = sampler.stop() if info.rank == 0 else None
    mode = int(lo.stats.eval_mode)
    x_fin = np.concatenate([lo.getX(), lo.getsmX()])
    lo.close()
    tm = torch.tensor([admm_s, total_s, eval_s, kmer_s, plan_s, build_s], dtype=torch.float64, device=dev)
    ev = torch.tensor([evals, eval_bytes, launches], dtype=torch.float64, device=dev)
    same = True
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
        t = torch.from_numpy(x_fin).to(dev)
        allx = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allx, t)
        same = bool(all(torch.equal(allx[0], a) for a in allx[1:]))
    admm_s, total_s, eval_s, kmer_s, plan_s, build_s = [float(v) for v in tm.cpu()]
    evals_all, eval_bytes_all, launches_all = [float(v) for v in ev.cpu()]
    if info.rank != 0:
        return None
    peak, peak_src = measured_peak_hbm()
    achieved = eval_bytes_all / world / eval_s / 1e9
    value = nb * T * iters / admm_s
    # fp64 tensor (DMMA) roofline of the packed-count kernels: useful FLOPs 4 n_b C dim per block
    # evaluation (two products, LOne.java:936-1049) against the DMMA peak this repo measured on B200
    # (profiles/r01_microbench_dmma.log: 36.6 TFLOP/s at 16 cycles per m8n8k4 and sub-partition;
    # MEASURED_PEAKS.json only holds the bf16 figure, which no fp64 path can use)
    dmma_peak = 36.6
    flops_useful = 4.0 * nb * C * dim * evals_all / world
    # classes the fp64 pipe works on: padded to whole 8-class DMMA tiles, except for C = 8 n + 1 in
    # mode 11 (n tiles + the last class as DFMAs)
    Cp = C if mode == 11 else (C + 7) // 8 * 8
    flops_issued = 4.0 * nb * Cp * ((dim + 15) // 16 * 16) * evals_all / world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(cfg, world, "strong", A), "sites": sites, "admm_blocks": T,
                   "rows_per_block": nb, "blocks_per_gpu": T // world, "columns": dim, "classes": C,
                   "eval_mode": mode,
                   "representation": ("packed k-mer counts (uint8) + xMean / xSD, streamed column chunk by "
                                      "column chunk (csrc/admm_eval_packed_chunk.cuh)" if mode in (10, 11) else
                                      "fp64 standardised matrix"),
                   "l2": "the training matrix of a rank (GBs as bytes, 8x that as fp64) is far larger than L2 "
                         "(126 MB): every evaluation streams it from HBM twice",
                   "step": f"{A} ADMM iterations of outer iteration 1 on the resident matrix"},
        "parity": {"ranks_bitwise_equal": same, "finite": bool(np.isfinite(x_fin).all()),
                   "oracle": "reduced-size shapes of this configuration in tests/ (-m gpu); the oracle "
                             "needs hours at this size"},
        "seconds_admm_per_step": admm_s / args.steps, "admm_iters_per_step": iters / args.steps,
        "evals_per_step": evals_all / args.steps, "block_evals_per_s": evals_all / admm_s,
        "ms_per_admm_iteration": admm_s / max(1, iters) * 1e3,
        "pipeline_s": {"host_setup": t_setup, "kmer_count_device": kmer_s,
                       "column_plan_device": plan_s, "counts_to_training_matrix": build_s,
                       "nccl_comm_init_once_per_process": comm_s},
        "kmer_gbases_per_s": sites * LENGTH / kmer_s / 1e9,
        "kmer_roofline": {"bound": "hbm", "achieved": sites * (LENGTH + 4 * kc.numK) / kmer_s / 1e9 / world,
                          "peak": peak, "unit": "GB/s",
                          "frac": sites * (LENGTH + 4 * kc.numK) / kmer_s / 1e9 / world / peak,
                          "count_matrix_bytes_per_gpu": int(counts_bytes)},
        "roofline": {"bound": "hbm", "kernel": "admm_xupdate_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": eval_bytes_all / world / max(1, iters),
                     "note": "per GPU: two-pass fp64 algorithmic bytes (SURVEY 8d) of the rank's evaluations over "
                             "the summed durations of its x-update launches (max over ranks). The packed-count "
                             "kernels move 1/8 of these bytes; what bounds them is roofline_fp64_tensor"},
        "roofline_fp64_tensor": {"bound": "tensor", "kernel": "admm_xupdate_kernel", "unit": "TFLOP/s",
                                 "achieved": flops_useful / eval_s / 1e12, "peak": dmma_peak,
                                 "frac": flops_useful / eval_s / 1e12 / dmma_peak,
                                 "issued": flops_issued / eval_s / 1e12,
                                 "frac_issued": flops_issued / eval_s / 1e12 / dmma_peak,
                                 "peak_source": "profiles/r01_microbench_dmma.log (DMMA m8n8k4, this pool's B200)",
                                 "note": "achieved = 4 n_b C dim FLOPs per block evaluation; issued also counts "
                                         "the classes padded to a multiple of 8 (C = 12 -> 16; C = 17 runs as "
                                         "16 + 1 in mode 11, 24 in mode 10) and columns padded to 16, which "
                                         "the fp64 pipe executes"},
        "e2e": {"value": nb * T * iters / (admm_s + (kmer_s + plan_s + build_s) * args.steps), "unit": UNIT,
                "h2d_bytes_per_step": int(d_bases.numel() * world + 12 * sites),
                "d2h_bytes_per_step": int(2 * (x0.nbytes + smx0.nbytes)),
                "note": "sequences in (H2D) -> counts -> plan -> training matrix -> ADMM -> weights out; "
                        "the pipeline stages are timed once (pipeline_s) and charged to every step"},
        "gpu_launches": int(launches_all), "clocks": clocks,
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--admm-iters", type=int, default=10,
                    help="cfg3 / cfg4: the fixed number of ADMM iterations per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-other-scaling", action="store_true")
    args = ap.parse_args()
    if args.config != "cfg2":
        args.scaling = "strong"
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner
    # under NCCL_DEBUG, for one) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from sequnwinder_b200 import parallel

    info = parallel.rank_info_from_env()
    world = info.world
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(info.local_rank)
    dev = torch.device("cuda", info.local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = (info.rank, world, parallel.share_nccl_id())
    if args.config == "cfg2":
        line = run_cfg2(args, info, world, dev, comm, dist, torch)
    else:
        line = run_big(args, info, world, dev, comm, dist, torch)
    if line is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
