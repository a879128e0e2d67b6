"""Host-side mirror of the caller of the trainer boundary, ``Classifier.buildClassifier``
(framework/Classifier.java:291-478, relative to
/root/reference/src/org/seqcode/projects/sequnwinder/), without Weka:

* RemoveUseless on numeric attributes == drop columns that are constant over the training rows
  (Classifier.java:305-308);
* instance-weighted mean / SD with the (W-1) denominator and abs() guard, in-place
  standardisation (:337-381);
* null-model initialisation of x and sm_x from unweighted class counts (:383-420);
* optimizer wiring in the reference's order (:422-449) — here always ``LOneCuda``;
* de-standardisation into m_Par / sm_Par (:452-477) and ``evaluateProbability`` (:267-287).

This is host glue (NumPy), not a hot path; the training itself runs on the GPU through
``LOneCuda``. Row sums use NumPy's pairwise summation, so means/SDs agree with the reference's
sequential loops to rounding (checked against the oracle in tests/).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from .optimizer import LOneCuda
from .structure import ClassRelationStructure


@dataclass
class Prepared:
    data: np.ndarray     # standardised, column 0 == 1
    keep: np.ndarray     # indices of the surviving count columns
    mean: np.ndarray
    sd: np.ndarray
    x0: np.ndarray
    sm_x0: np.ndarray


def prepare(counts: np.ndarray, cls: np.ndarray, weights: np.ndarray,
            structure: ClassRelationStructure, num_classes: int, materialize: bool = True) -> Prepared:
    """Classifier.java:300-420 on a dense count matrix. ``materialize=False`` skips the
    standardised fp64 matrix (``data`` is None): the packed-count trainer path only needs the
    kept columns, xMean and xSD."""
    counts = np.asarray(counts)
    n = counts.shape[0]
    keep = np.nonzero(counts.max(axis=0) != counts.min(axis=0))[0]
    nR = keep.size
    data = np.empty((n, nR + 1), dtype=np.float64)
    data[:, 0] = 1.0
    data[:, 1:] = counts[:, keep]
    w = np.asarray(weights, dtype=np.float64)
    tot = float(w.sum())
    if tot <= 1 and n > 1:
        raise ValueError("Sum of weights of instances less than 1, please reweight!")
    mean = np.zeros(nR + 1)
    sd = np.ones(nR + 1)
    xs = data[:, 1:]
    m = (w[:, None] * xs).sum(axis=0) / tot
    if tot > 1:
        s = np.sqrt(np.abs((w[:, None] * xs * xs).sum(axis=0) - tot * m * m) / (tot - 1))
    else:
        s = np.zeros(nR)
    mean[1:], sd[1:] = m, s
    nz = sd != 0
    if materialize:
        data[:, nz] = (data[:, nz] - mean[nz]) / sd[nz]
    else:
        data = None

    dim = nR + 1
    NN = structure.numNodes
    sY = np.bincount(np.asarray(cls), minlength=num_classes).astype(np.float64)
    sm_sY = np.zeros(NN)
    for leaf in structure.leafs:
        sm_sY[leaf.nodeIndex] = sY[leaf.nodeIndex]
    for l in range(1, structure.numLayers):
        for node in structure.layers[l]:
            for c in node.children:
                sm_sY[node.nodeIndex] += sm_sY[c]
    x0 = np.zeros(num_classes * dim)
    x0[0::dim] = np.log(sY + 1.0) - np.log(sY[num_classes - 1] + 1.0)
    sm_x0 = np.zeros(NN * dim)
    sm_x0[0::dim] = np.log(sm_sY + 1.0) - np.log(sm_sY[num_classes - 1] + 1.0)
    return Prepared(data, keep, mean, sd, x0, sm_x0)


def destandardize(v: np.ndarray, ncols: int, dim: int, mean: np.ndarray, sd: np.ndarray):
    """Classifier.java:452-477 -> par[dim, ncols] (m_Par / sm_Par layout)."""
    V = np.asarray(v, dtype=np.float64).reshape(ncols, dim)
    par = V.T.copy()
    nz = sd[1:] != 0
    par[1:][nz] = par[1:][nz] / sd[1:][nz, None]
    par[0] -= (par[1:][nz] * mean[1:][nz, None]).sum(axis=0)
    return par


def evaluate_probability(par: np.ndarray, raw: np.ndarray) -> np.ndarray:
    """Classifier.java:267-287 for many rows: raw has column 0 == 1."""
    v = raw @ par
    v = v - v.max(axis=1, keepdims=True)
    e = np.exp(v)
    return e / e.sum(axis=1, keepdims=True)


class Classifier:
    """Weka-free stand-in for framework/Classifier.java with the options of
    SeqUnwinderConfig.java:388-413 (-R -PHO -threads -A -S)."""

    def __init__(self, structure: ClassRelationStructure, ridge: float = 10.0, pho: float = 1.7,
                 num_threads: int = 5, admm_max_its: int = 400, sequnwinder_max_its: int = 15,
                 debug: bool = False):
        self.sm_ClassStructure = structure
        self.m_Ridge, self.m_ADMM_pho = ridge, pho
        self.m_ADMM_numThreads = num_threads
        self.m_ADMM_MaxIts, self.m_SeqUnwinder_MaxIts = admm_max_its, sequnwinder_max_its
        self.m_BGFS_MaxIts = -1
        self.sm_Debug = debug
        self.m_Par: Optional[np.ndarray] = None
        self.sm_Par: Optional[np.ndarray] = None
        self.prep: Optional[Prepared] = None
        self.optimizer: Optional[LOneCuda] = None
        self.comm = None   # (rank, world, nccl_id) for multi-GPU runs
        self.cta_budget = 0  # > 0: share of the GPU when several classifiers train side by side
        self.use_counts = True   # train from the byte count matrix when the counts allow it

    def buildClassifier(self, counts: np.ndarray, cls: np.ndarray, weights: np.ndarray,
                        num_classes: int):
        counts = np.asarray(counts)
        # k-mer counts fit a byte (MakeArff.java:347-368: at most L - k + 1 per window length): hand
        # the trainer the counts + xMean / xSD instead of the standardised doubles (packed-count path)
        packed = self.use_counts and counts.size > 0 and counts.min() >= 0 and counts.max() <= 255
        prep = prepare(counts, cls, weights, self.sm_ClassStructure, num_classes, materialize=not packed)
        self.prep = prep
        x, sm_x = prep.x0.copy(), prep.sm_x0.copy()
        opt = LOneCuda(x, sm_x, prep.data)                  # Classifier.java:424-425
        if packed:
            opt.setCountMatrix(counts[:, prep.keep].astype(np.uint8), prep.mean, prep.sd)
        opt.setRidge(self.m_Ridge)                          # :427
        opt.setADMMmaxItrs(self.m_ADMM_MaxIts)              # :429
        opt.setBGFSmaxItrs(self.m_BGFS_MaxIts)              # :431
        opt.setSeqUnwinderMaxIts(self.m_SeqUnwinder_MaxIts)  # :432
        opt.setClassStructure(self.sm_ClassStructure)       # :433
        opt.setClsMembership(np.asarray(cls, dtype=np.int32))  # :434
        opt.setInstanceWeights(np.asarray(weights, dtype=np.float64))  # :435
        opt.setNumClasses(num_classes)                      # :436
        opt.setNumPredictors(prep.keep.size)                # :437
        opt.setPho(self.m_ADMM_pho)                         # :439
        opt.set_numThreads(self.m_ADMM_numThreads)          # :440
        opt.initUandZ()                                     # :441
        opt.setDebugMode(self.sm_Debug)                     # :443
        if self.comm is not None:
            opt.setComm(*self.comm)
        if self.cta_budget:
            opt.setCtaBudget(self.cta_budget)
        opt.execute()                                       # :444
        x, sm_x = opt.getX(), opt.getsmX()                  # :447-448
        dim = prep.keep.size + 1
        self.m_Par = destandardize(x, num_classes, dim, prep.mean, prep.sd)
        self.sm_Par = destandardize(sm_x, self.sm_ClassStructure.numNodes, dim, prep.mean, prep.sd)
        self.optimizer = opt
        self.x, self.sm_x = x.copy(), sm_x.copy()
        opt.clearOptimizer()                                # :449
        return self

    def distributionForInstances(self, counts: np.ndarray) -> np.ndarray:
        """Classifier.java:244-265 for a batch of raw count rows."""
        raw = np.empty((counts.shape[0], self.prep.keep.size + 1))
        raw[:, 0] = 1.0
        raw[:, 1:] = counts[:, self.prep.keep]
        return evaluate_probability(self.m_Par, raw)

    def getWeights(self, kmer_names: List[str]) -> Dict[str, np.ndarray]:
        """Classifier.java:173-193: per node, a numK-long vector of k-mer weights."""
        out = {}
        for idx, name in enumerate(self.sm_ClassStructure.node_names()):
            v = np.zeros(len(kmer_names))
            v[self.prep.keep] = self.sm_Par[1:, idx]
            out[name] = v
        return out


# ---------------------------------------------------------------------------------------------
# Cross-validation (SeqUnwinder.java:41: weka Evaluation.crossValidateModel, "--x" folds): the
# model on all sites plus one model per fold, each through Classifier.buildClassifier exactly as
# Weka's loop does (train on the complement of the fold, predict the fold, pool the hits).
@dataclass
class CVResult:
    accuracy: float              # pooled over the test folds (Evaluation.pctCorrect() / 100)
    fold_accuracy: List[float]
    predictions: np.ndarray      # arg-max class of every site that lies in a fold, else -1
    full_model: "Classifier"
    fold_models: List["Classifier"]


def stratified_folds(cls: np.ndarray, num_folds: int, seed: int = 1) -> List[np.ndarray]:
    """Deterministic stratified test-fold index lists: a seeded shuffle, then the sites of every
    class dealt round-robin. These are NOT Weka's folds (Instances.randomize / stratify with
    seed 1; Weka is not available here): a caller that needs Weka's exact folds
    passes its own index lists to cross_validate."""
    cls = np.asarray(cls)
    rng = np.random.default_rng(seed)
    order = rng.permutation(cls.size)
    folds: List[List[int]] = [[] for _ in range(num_folds)]
    nxt = 0
    for c in np.unique(cls):
        for i in order[cls[order] == c]:
            folds[nxt % num_folds].append(int(i))
            nxt += 1
    return [np.sort(np.asarray(f, dtype=np.int64)) for f in folds]


def cross_validate(structure: ClassRelationStructure, counts: np.ndarray, cls: np.ndarray,
                   weights: np.ndarray, num_classes: int, folds: List[np.ndarray], slots: int = 1,
                   **classifier_options) -> CVResult:
    """1 + len(folds) trainings. ``slots`` > 1 runs them side by side on one GPU, each on
    148 // slots CTAs (LOneCuda.setCtaBudget): the trainings are independent, so this
    is the job-level batching the reference's one-JVM-thread loop cannot do. Measured on B200 at
    cfg-2 with 3 folds (count matrix, rows panel by panel through shared memory): 12.5 s with one
    slot, 11.6 s with three, 8.7 s with four (all four trainings at once)."""
    import queue
    import threading

    counts, cls, weights = np.asarray(counts), np.asarray(cls), np.asarray(weights)
    n = counts.shape[0]
    slots = max(1, min(int(slots), 1 + len(folds)))
    budget = 0 if slots == 1 else 148 // slots
    tasks: "queue.SimpleQueue" = queue.SimpleQueue()
    tasks.put((-1, np.arange(n)))
    for f, test in enumerate(folds):
        train = np.setdiff1d(np.arange(n), np.asarray(test), assume_unique=False)
        tasks.put((f, train))
    models: Dict[int, Classifier] = {}
    errors: list = []
    # the CUDA current device is per host thread: workers train on the caller's device
    from . import _lib
    parent_dev = _lib.load().sqw_get_device()

    def worker():
        try:
            if parent_dev >= 0:
                _lib.check(_lib.load().sqw_set_device(parent_dev))
            while not errors:
                try:
                    f, rows = tasks.get_nowait()
                except queue.Empty:
                    return
                clf = Classifier(structure, **classifier_options)
                clf.cta_budget = budget
                clf.buildClassifier(counts[rows], cls[rows], weights[rows], num_classes)
                models[f] = clf
        except Exception as exc:  # surfaced to the caller below
            errors.append(exc)

    if slots == 1:
        worker()              # inline, on the calling thread
    else:
        threads = [threading.Thread(target=worker) for _ in range(slots)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    pred = np.full(n, -1, dtype=np.int64)
    fold_acc = []
    for f, test in enumerate(folds):
        test = np.asarray(test)
        pf = models[f].distributionForInstances(counts[test]).argmax(axis=1)
        pred[test] = pf
        fold_acc.append(float((pf == cls[test]).mean()) if test.size else float("nan"))
    seen = pred >= 0
    acc = float((pred[seen] == cls[seen]).mean()) if seen.any() else float("nan")
    return CVResult(acc, fold_acc, pred, models[-1], [models[f] for f in range(len(folds))])


# ---------------------------------------------------------------------------------------------
# Device-resident versions of the steps either side of the trainer: torch
# tensors are only device memory here; the arithmetic is in csrc/prepare_predict.cu.
# ---------------------------------------------------------------------------------------------
class DevicePrepared:
    """Result of ``prepare_device``: cuda tensors X [n, F+1] fp64, keep_idx [F] int32,
    mean / sd [numK] fp64 (indexed by original column)."""

    def __init__(self, X, keep_idx, mean, sd):
        self.X, self.keep_idx, self.mean, self.sd = X, keep_idx, mean, sd

    @property
    def num_kept(self) -> int:
        return int(self.keep_idx.numel())


def prepare_device(counts, weights) -> DevicePrepared:
    """counts: cuda int32 [n, numK]; weights: cuda fp64 [n]. Classifier.java:300-381 on the GPU."""
    import ctypes as C
    import torch
    from . import _lib

    lib = _lib.load()
    n, numK = counts.shape
    dev = counts.device
    keep = torch.empty(numK, dtype=torch.int32, device=dev)
    mean = torch.empty(numK, dtype=torch.float64, device=dev)
    sd = torch.empty(numK, dtype=torch.float64, device=dev)
    nk = C.c_int32(0)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    tot = float(weights.sum().item())
    rc = lib.sqw_prepare_plan_dev(counts.data_ptr(), n, numK, weights.data_ptr(), tot, keep.data_ptr(),
                                  C.byref(nk), mean.data_ptr(), sd.data_ptr(), stream)
    if rc == _lib.SQW_ERR_INVALID_ARG and b"Sum of weights" in lib.sqw_last_error():
        raise ValueError("Sum of weights of instances less than 1, please reweight!")
    _lib.check(rc)
    F = nk.value
    keep = keep[:F].contiguous()
    X = torch.empty((n, F + 1), dtype=torch.float64, device=dev)
    _lib.check(lib.sqw_prepare_apply_dev(counts.data_ptr(), n, numK, keep.data_ptr(), F,
                                         mean.data_ptr(), sd.data_ptr(), X.data_ptr(), stream))
    return DevicePrepared(X, keep, mean, sd)


def plan_device(counts, weights, total_weight=None, n_total=None, group=None, extra=None) -> DevicePrepared:
    """The column plan only (keep_idx, xMean, xSD; ``X`` is None): what the packed-count trainer
    entry point needs (``LOneCuda.setCountMatrixDevice``). With ``group`` (a torch.distributed
    process group, or True for the default one) the rows are sharded over the ranks: every rank
    passes ITS rows, the per-column statistics are combined with all-reduces (min / max / sum) and
    every rank gets the plan of the whole matrix; ``total_weight`` / ``n_total`` are then the
    global sums (computed with an all-reduce when omitted). ``extra`` = (counts2, weights2): further
    rows of this rank that take part in the statistics (e.g. the rows the ADMM blocking drops)."""
    import ctypes as C
    import torch
    from . import _lib

    lib = _lib.load()
    n, numK = counts.shape
    dev = counts.device
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    mn = torch.empty(numK, dtype=torch.int32, device=dev)
    mx = torch.empty(numK, dtype=torch.int32, device=dev)
    swx = torch.empty(numK, dtype=torch.float64, device=dev)
    swxx = torch.empty(numK, dtype=torch.float64, device=dev)
    _lib.check(lib.sqw_prepare_stats_dev(counts.data_ptr(), n, numK, weights.data_ptr(), mn.data_ptr(),
                                         mx.data_ptr(), swx.data_ptr(), swxx.data_ptr(), stream))
    tot = weights.sum().reshape(1)
    cnt = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    if extra is not None and extra[0].shape[0] > 0:
        c2, w2 = extra
        mn2, mx2, swx2, swxx2 = (torch.empty_like(t) for t in (mn, mx, swx, swxx))
        _lib.check(lib.sqw_prepare_stats_dev(c2.data_ptr(), c2.shape[0], numK, w2.data_ptr(), mn2.data_ptr(),
                                             mx2.data_ptr(), swx2.data_ptr(), swxx2.data_ptr(), stream))
        mn, mx = torch.minimum(mn, mn2), torch.maximum(mx, mx2)
        swx, swxx = swx + swx2, swxx + swxx2
        tot = tot + w2.sum().reshape(1)
        cnt = cnt + float(c2.shape[0])
    if group is not None:
        import torch.distributed as dist
        g = None if group is True else group
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=g)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=g)
        dist.all_reduce(swx, op=dist.ReduceOp.SUM, group=g)
        dist.all_reduce(swxx, op=dist.ReduceOp.SUM, group=g)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=g)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=g)
    tw = float(tot.item()) if total_weight is None else float(total_weight)
    nt = int(cnt.item()) if n_total is None else int(n_total)
    keep = torch.empty(numK, dtype=torch.int32, device=dev)
    mean = torch.empty(numK, dtype=torch.float64, device=dev)
    sd = torch.empty(numK, dtype=torch.float64, device=dev)
    nk = C.c_int32(0)
    rc = lib.sqw_prepare_finish_dev(mn.data_ptr(), mx.data_ptr(), swx.data_ptr(), swxx.data_ptr(), numK,
                                    tw, nt, keep.data_ptr(), C.byref(nk), mean.data_ptr(), sd.data_ptr(),
                                    stream)
    if rc == _lib.SQW_ERR_INVALID_ARG and b"Sum of weights" in lib.sqw_last_error():
        raise ValueError("Sum of weights of instances less than 1, please reweight!")
    _lib.check(rc)
    return DevicePrepared(None, keep[:nk.value].contiguous(), mean, sd)


def predict_device(par, counts, keep_idx, want_prob: bool = True):
    """par: cuda fp64 [dim, C] (m_Par); counts: cuda int32 [n, numK]; keep_idx: cuda int32 [dim-1].
    Returns (prob [n, C] or None, argmax [n] int32). Classifier.java:244-287 on the GPU."""
    import ctypes as C
    import torch
    from . import _lib

    lib = _lib.load()
    dim, ncls = par.shape
    n, numK = counts.shape
    dev = counts.device
    prob = torch.empty((n, ncls), dtype=torch.float64, device=dev) if want_prob else None
    arg = torch.empty(n, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.sqw_predict_dev(par.data_ptr(), dim, ncls, counts.data_ptr(), n, numK,
                                   keep_idx.data_ptr(), prob.data_ptr() if want_prob else None,
                                   arg.data_ptr(), stream))
    return prob, arg
