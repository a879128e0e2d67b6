#!/usr/bin/env python
"""bench.py — SeqUnwinder's two hot paths on B200, on BASELINE.json's configs[1]
("README typical": 20,000 x 150 bp sites, 8 overlapping labels, k = 4..5, --r 10).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

One step = k-mer featurisation of the batch (device resident) + one full ADMM training
(LOne.execute: all outer iterations to the reference's tolerances) on the resulting matrix.
`value` = ADMM site*iter/s = floor(N/T)*T * (ADMM iterations executed) / seconds inside the
ADMM loops, inputs resident in HBM; `e2e` = the same through the public host-buffer API
(KmerCounter.count + LOneCuda.execute with H2D/D2H inside the timed region).
N > 1 is weak scaling: 20,000 sites and 8 ADMM blocks PER GPU (T = 8 N), one ncclAllReduce
pair per ADMM iteration for the consensus sums.
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
SITES_PER_GPU = 20000
LABELS = 8
BLOCKS_PER_GPU = 8
KMER_REPS = 10        # k-mer passes enqueued per timed step (see resident_step)
KMER_STREAM_X = 10    # the k-mer kernel alone on a batch this many times cfg-2 (output > L2)
KMIN, KMAX, LENGTH = 4, 5, 150
RIDGE, PHO, ADMM_MAX, NODES_MAX = 10.0, 1.7, 400, 15


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def build_problem(world: int):
    """Synthetic cfg-2 (x world for weak scaling): sequences, labels, design. Host only."""
    from sequnwinder_b200 import structure, synth

    n_sites = SITES_PER_GPU * world
    d = synth.generate(n_sites, LABELS, length=LENGTH, seed=synth.SEED_BASE + 2)
    des = structure.make_design(d.annotations)
    crs = structure.ClassRelationStructure(des.design, des.num_layers)
    cls = np.array([des.subgroup_names.index(s) for s in des.subgroups_at_sites], dtype=np.int32)
    w = np.array([des.subgroup_weights[s] for s in des.subgroups_at_sites], dtype=np.float64)
    return d, crs, cls, w, len(des.subgroup_names)


def make_optimizer(prep, crs, cls, w, C, T, comm=None, counts=None):
    from sequnwinder_b200.optimizer import LOneCuda

    if counts is not None:   # packed-count form: uint8 counts of the kept k-mers + xMean / xSD
        lo = LOneCuda(prep.x0.copy(), prep.sm_x0.copy(), None)
        lo.setCountMatrix(counts, prep.mean, prep.sd)
    else:
        lo = LOneCuda(prep.x0.copy(), prep.sm_x0.copy(), prep.data)
    lo.setRidge(RIDGE)
    lo.setADMMmaxItrs(ADMM_MAX)
    lo.setBGFSmaxItrs(-1)
    lo.setSeqUnwinderMaxIts(NODES_MAX)
    lo.setClassStructure(crs)
    lo.setClsMembership(cls)
    lo.setInstanceWeights(w)
    lo.setNumClasses(C)
    lo.setNumPredictors(prep.data.shape[1] - 1)
    lo.setPho(PHO)
    lo.set_numThreads(T)
    lo.initUandZ()
    lo.setDebugMode(False)
    if comm is not None:
        lo.setComm(*comm)
    return lo


def cpu_oracle_sample(prep_data, x0, smx0, cls, w, crs, C, T, admm_iters, threads):
    """The oracle (CPU restatement of LOne.java) on a bounded sample: `admm_iters` ADMM
    iterations of the first outer iteration, T blocks on `threads` host threads."""
    from oracle import pyoracle as O

    st = O.StructureArrays(**crs.to_arrays())
    t0 = time.time()
    _, _, tr = O.lone_execute(x0, smx0, prep_data, cls, w, st, C, ridge=RIDGE, pho=PHO,
                              num_threads=T, admm_max_itr=admm_iters, nodes_max_itr=1,
                              host_threads=threads)
    wall = time.time() - t0
    nb = prep_data.shape[0] // T
    iters = sum(tr.admm_iters)
    return nb * T * iters / tr.seconds_admm, tr.seconds_admm, iters, wall


def oracle_kmer_all_cores(O, d, threads):
    """The oracle's counting loop on `threads` host threads (contiguous row ranges; ctypes drops the
    GIL during the call): SURVEY.md 8d asks for the path-1 CPU number single-threaded and on all cores."""
    n = d.n_sites
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)

    def work(a, b):
        if b > a:
            b0 = int(d.offsets[a])
            O.kmer_count(d.bases[b0:int(d.offsets[b])], d.offsets[a:b + 1] - b0, KMIN, KMAX)

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
    port, all host threads it can use (one per ADMM block, LOne.java:599-606)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as O
    from sequnwinder_b200 import classifier as clf

    O.build()
    d, crs, cls, w, C = build_problem(1)
    counts, has_n = O.kmer_count(d.bases, d.offsets, KMIN, KMAX)
    prep = clf.prepare(counts, cls, w, crs, C)
    T = BLOCKS_PER_GPU
    cores = os.cpu_count() or 1
    threads = min(T, cores)
    sample_iters = 3
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, s, iters, _ = cpu_oracle_sample(prep.data, prep.x0, prep.sm_x0, cls, w, crs, C, T,
                                           sample_iters, threads)
        if i >= args.warmup:
            vals.append(v)
            secs.append(s)
    # k-mer path of the reference (counting loop only): single thread, and all host cores
    t0 = time.time()
    O.kmer_count(d.bases, d.offsets, KMIN, KMAX)
    kmer_gbs = d.n_sites * LENGTH / (time.time() - t0) / 1e9
    kmer_gbs_all = oracle_kmer_all_cores(O, d, cores)
    value = float(np.mean(vals))
    sample = (f"first {sample_iters} ADMM iterations of outer iteration 1 on the full "
              f"{SITES_PER_GPU}x{prep.data.shape[1]} matrix, T={T} blocks on {threads} pthreads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "cfg2: 20000x150bp sites, 8 ring labels, k=4..5, r=10, T=8 (CPU sample)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample, "host_cores": cores,
                         "note": "CPU restatement of LOne.java (oracle/), not the JVM; the Java "
                                 "reference additionally sleeps >= 1 s per ADMM run (LOne.java:728-740)"},
        "kmer_gbases_per_s": kmer_gbs,
        "kmer_gbases_per_s_all_cores": {"value": kmer_gbs_all, "cores": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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

    from sequnwinder_b200 import KmerCounter, parallel
    from sequnwinder_b200 import classifier as clf

    info = parallel.rank_info_from_env()
    world = info.world
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(info.local_rank)
    dev = torch.device("cuda", info.local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = None
    if world > 1:
        comm = (info.rank, world, parallel.share_nccl_id())

    # ---------------- setup (untimed): data, k-mer matrix, standardised training matrix
    d, crs, cls, w, C = build_problem(world)
    n_sites = d.n_sites
    T = BLOCKS_PER_GPU * world
    kc = KmerCounter(KMIN, KMAX)
    d_bases = torch.from_numpy(d.bases).to(dev)
    d_offs = torch.from_numpy(d.offsets).to(dev)
    counts_t, has_n_t, status_t = kc.count_device(d_bases, d_offs, LENGTH)
    torch.cuda.synchronize()
    counts = counts_t.cpu().numpy()
    prep = clf.prepare(counts, cls, w, crs, C)
    dim = prep.data.shape[1]
    nb = n_sites // T

    packed = os.environ.get("SQW_BENCH_FP64") is None
    cnt8 = np.ascontiguousarray(counts[:, prep.keep]).astype(np.uint8) if packed else None
    lo = make_optimizer(prep, crs, cls, w, C, T, comm, cnt8)
    lo.setResident(True)

    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        """k-mer kernel on resident bases + full ADMM training on the resident matrix."""
        # the k-mer pass is 30 us of GPU time, less than the host needs to launch it: it is
        # enqueued KMER_REPS times between the two events (same inputs, same output buffer) so
        # that the event difference is GPU time and not launch latency; the time of ONE pass is
        # what enters the step
        ev0.record()
        for _ in range(KMER_REPS):
            kc.count_device(d_bases, d_offs, LENGTH, out=counts_t, has_n=has_n_t, status=status_t)
        ev1.record()
        np.copyto(lo.x, prep.x0)
        np.copyto(lo.sm_x, prep.sm_x0)
        lo.execute()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) * 1e-3 / KMER_REPS, lo.stats

    for _ in range(args.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(info.local_rank)
    if info.rank == 0:
        sampler.start()
    t_wall0 = time.time()
    kmer_s, admm_s, total_s, eval_s = 0.0, 0.0, 0.0, 0.0
    iters, evals, launches, eval_bytes = 0, 0, 0, 0
    outer = []
    for _ in range(args.steps):
        ks, st = resident_step()
        kmer_s += ks
        admm_s += st.seconds_admm
        total_s += st.seconds_total + ks
        eval_s += st.seconds_eval
        iters += st.total_admm_iters
        evals += st.total_evals
        eval_bytes += st.eval_bytes
        launches += st.launches + KMER_REPS
        outer.append(st.outer_iters)
    barrier()
    wall = time.time() - t_wall0
    clocks = sampler.stop() if info.rank == 0 else None

    # max over ranks of the device-timed seconds
    tm = torch.tensor([admm_s, total_s, kmer_s, eval_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    admm_s, total_s, kmer_s, eval_s = [float(v) for v in tm.cpu()]
    value = nb * T * iters / admm_s
    kmer_gbases = n_sites * LENGTH * args.steps / kmer_s / 1e9

    # ---------------- the k-mer kernel alone on a batch whose output exceeds L2: 10 x the cfg-2
    # sequences in ONE launch (1 GB of counts), CUDA events around three launches after a warm-up
    def kmer_stream():
        big_bases = d_bases[: SITES_PER_GPU * LENGTH].repeat(KMER_STREAM_X)
        n_big = SITES_PER_GPU * KMER_STREAM_X
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
        ok = bool((out[:SITES_PER_GPU] == counts_t[:SITES_PER_GPU]).all().item()) and int(stt.item()) == 0
        return n_big, sec, ok

    ks_n, ks_sec, ks_ok = kmer_stream() if info.rank == 0 else (0, 1.0, True)

    # ---------------- e2e through the public host-buffer API (H2D/D2H inside the timed region)
    # path 1 shards by rows with no exchange: every rank featurises its own rows
    r_a, r_b = n_sites * info.rank // world, n_sites * (info.rank + 1) // world
    b_a, b_b = int(d.offsets[r_a]), int(d.offsets[r_b])
    my_bases = np.ascontiguousarray(d.bases[b_a:b_b])
    my_offs = np.ascontiguousarray(d.offsets[r_a:r_b + 1] - b_a)

    kmer_host_s = []   # wall seconds of the host-buffer k-mer call of every e2e step (this rank's rows)

    def e2e_step():
        t0 = time.time()
        c, hn = kc.count(my_bases, my_offs)
        t1 = time.time()
        lo2 = make_optimizer(prep, crs, cls, w, C, T, comm, cnt8)
        lo2._open()          # handle + problem upload (execute() would do it)
        t2 = time.time()
        lo2.execute()
        _ = float(lo2.getX()[0])
        dt = time.time() - t0
        print(f"[bench] rank {info.rank} e2e step: k-mer (host buffers) {t1 - t0:.3f} s, upload "
              f"{t2 - t1:.3f} s, execute {dt - (t2 - t0):.3f} s (device {lo2.stats.seconds_total:.3f} s)",
              file=sys.stderr)
        kmer_host_s.append(t1 - t0)
        return dt, lo2.stats.total_admm_iters, lo2.h2d_bytes, (c.nbytes + hn.nbytes) * world

    lo.close()
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
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nb * T * e2e_iters / float(te.item())
    # whole-job bytes: the k-mer shards of all ranks, the training matrix once (each rank uploads
    # only the rows of its own blocks)
    h2d += (my_bases.nbytes + my_offs.nbytes) * world + prep.x0.nbytes + prep.sm_x0.nbytes
    d2h += prep.x0.nbytes + prep.sm_x0.nbytes

    if info.rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_hbm()
    achieved = eval_bytes / eval_s / 1e9 if eval_s > 0 else 0.0
    # DRAM traffic per x-update (the phased launches of one ADMM iteration taken together): ncu
    # --set full capture of round 1 (profiles/r01_ncu_prof_xupdate_r1k.txt): 842.3 MB read + 30.0 MB
    # written for the 8 evaluation rounds of the captured phase-0 launch = 109.0 MB per evaluation
    # round of 8 blocks (one pass over the 104 MB matrix; the exchange buffers stay in L2), scaled
    # to this run's average evaluations per block and ADMM iteration.
    xupdate_launches = max(1, iters)
    traffic = 109.0e6 * (evals / BLOCKS_PER_GPU / world) / xupdate_launches if world == 1 else None
    kmer_bytes = n_sites * (LENGTH + 4 * kc.numK) * args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"cfg2 per GPU: {SITES_PER_GPU}x{LENGTH}bp sites, {LABELS} ring labels "
                        f"(C={C}), k={KMIN}..{KMAX} (numK={kc.numK}, F={dim - 1}), r={RIDGE}, "
                        f"pho={PHO}, A={ADMM_MAX}, S={NODES_MAX}, T={BLOCKS_PER_GPU} blocks per GPU",
            "sites": n_sites, "admm_blocks": T, "rows_per_block": nb,
            "l2": "training matrix (104 MB per GPU) is re-read every evaluation by a persistent "
                  "kernel; it is smaller than L2 (126 MB), no flush between evaluations - the "
                  "algorithm itself re-reads it; microbench read of 104 MB runs at HBM speed",
            "step": "k-mer featurisation (resident) + LOne.execute to the reference tolerances",
        },
        "time_to_converge_s": admm_s / args.steps,
        "admm_iters_per_step": iters / args.steps,
        "outer_iters": outer,
        "evals_per_step": evals / args.steps,
        "kmer_gbases_per_s": kmer_gbases,
        "kmer_roofline": {"bound": "hbm", "achieved": kmer_bytes / kmer_s / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": kmer_bytes / kmer_s / 1e9 / peak,
                          "note": "cfg-2 batch (102 MB of counts per pass): launch/latency bound"},
        # the same featurisation through the host-buffer call (H2D of the bases, D2H
        # of the count matrix to pageable memory inside the timing), rank 0's rows
        "kmer_host_api_gbases_per_s": (r_b - r_a) * LENGTH * len(kmer_host_s) / max(sum(kmer_host_s), 1e-9) / 1e9,
        "kmer_stream": {"sites": ks_n, "gbases_per_s": ks_n * LENGTH / ks_sec / 1e9,
                        "bound": "hbm", "achieved": ks_n * (LENGTH + 4 * kc.numK) / ks_sec / 1e9,
                        "peak": peak, "unit": "GB/s",
                        "frac": ks_n * (LENGTH + 4 * kc.numK) / ks_sec / 1e9 / peak,
                        "ms_per_launch": ks_sec * 1e3, "matches_cfg2_counts": ks_ok,
                        "note": f"kmer_count kernel alone, {KMER_STREAM_X} x the cfg-2 sequences in "
                                "one launch (output larger than L2), CUDA events"},
        "roofline": {"bound": "hbm", "kernel": "admm_xupdate_kernel", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": eval_bytes / max(1, iters),
                     "peak_source": peak_src,
                     "launch_unit": "one ADMM iteration's x-update = 6 phased cooperative launches "
                                    "of admm_xupdate_kernel; achieved = algorithmic bytes / summed "
                                    "CUDA-event durations of those launches",
                     "note": "algorithmic bytes = evaluations x (2 n_b dim 8 + 2 n_b C 8 + 12 n_b "
                             "+ 2 C dim 8) over the summed durations of the x-update kernels, which "
                             "also contain the L-BFGS phases"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "seconds_per_step": float(te.item()) / e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "wall_s": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as O
        O.build()
        cores = os.cpu_count() or 1
        threads = min(T, cores)
        v, s, its, wl = cpu_oracle_sample(prep.data, prep.x0, prep.sm_x0, cls, w, crs, C, T, 3, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": threads, "kind": "port", "host_cores": cores,
            "sample": f"first {its} ADMM iterations of outer iteration 1, full cfg2 matrix, "
                      f"T={T} blocks on {threads} pthreads ({s:.1f} s)"}
        # the reference's default --threads 5 as well: same matrix, 5 blocks
        v5, s5, its5, _ = cpu_oracle_sample(prep.data, prep.x0, prep.sm_x0, cls, w, crs, C, 5, 2,
                                            min(5, cores))
        line["cpu_baseline"]["threads5"] = {
            "value": v5, "cores": min(5, cores),
            "sample": f"first {its5} ADMM iterations, T=5 blocks on {min(5, cores)} pthreads ({s5:.1f} s)"}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
