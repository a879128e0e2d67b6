"""Host-side mirror of the reference's trainer plug-in API.

``Optimizer`` restates the abstract class framework/Optimizer.java:5-27 (same method names and
argument meaning); ``LOneCuda`` is the B200 implementation that takes the place of
framework/LOne.java behind it: ``Classifier.buildClassifier`` constructs it, pushes the setters in
the order of framework/Classifier.java:422-443, calls ``execute()`` and pulls ``getX()`` /
``getsmX()`` (:444-449). All compute goes through the C ABI (``sqw_admm_*``); there is no CPU
path. Citations are relative to /root/reference/src/org/seqcode/projects/sequnwinder/.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib
from .structure import ClassRelationStructure


class Optimizer(ABC):
    """framework/Optimizer.java:5-27"""

    @abstractmethod
    def setBGFSmaxItrs(self, m: int): ...

    @abstractmethod
    def setSeqUnwinderMaxIts(self, m: int): ...

    @abstractmethod
    def setInstanceWeights(self, w): ...

    @abstractmethod
    def setClsMembership(self, c): ...

    @abstractmethod
    def setNumPredictors(self, p: int): ...

    @abstractmethod
    def setNumClasses(self, c: int): ...

    @abstractmethod
    def setClassStructure(self, rel: ClassRelationStructure): ...

    @abstractmethod
    def setNumNodes(self, n: int): ...

    @abstractmethod
    def setRidge(self, r: float): ...

    @abstractmethod
    def setDebugMode(self, debug: bool): ...

    @abstractmethod
    def execute(self): ...

    @abstractmethod
    def getX(self): ...

    @abstractmethod
    def getsmX(self): ...

    @abstractmethod
    def clearOptimizer(self): ...


@dataclass
class AdmmLog:
    outer: np.ndarray
    itr: np.ndarray
    pho: np.ndarray
    fold: np.ndarray
    primal: np.ndarray
    dual: np.ndarray
    nfev: np.ndarray          # [rows, local blocks]
    admm_iters: List[int]     # ADMM_currItr_value at exit, per outer iteration


class LOneCuda(Optimizer):
    """Drop-in for ``new LOne(xinit, sm_xinit, data)`` (LOne.java:110-114).

    ``x`` (numClasses*dim, class-major, intercept first) and ``sm_x`` (numNodes*dim) are the
    caller's arrays and receive the trained values in place, as in the reference;
    ``data`` is the standardised n x dim matrix with column 0 == 1 (Classifier.java:337-381).
    """

    def __init__(self, xinit: np.ndarray, sm_xinit: np.ndarray, data: np.ndarray):
        self.x = xinit
        self.sm_x = sm_xinit
        self.data = data
        # LOne.java:24-51 defaults
        self.ADMM_maxItr = 500
        self.ADMM_pho = 1.7
        self.ADMM_numThreads = 5
        self.BGFS_maxIts = -1
        self.NODES_maxItr = 10
        self.regularization = 0.0
        self.numPredictors = None
        self.numClasses = None
        self.numNodes = None
        self.classStructure: Optional[ClassRelationStructure] = None
        self.weights = None
        self.cls = None
        self.sm_Debug = False
        self._rank, self._world, self._nccl_id = 0, 1, None
        self._handle = None
        self._resident = False
        self._cta_budget = 0
        self.h2d_bytes = 0
        self.stats: Optional[_lib.AdmmStats] = None
        self.log: Optional[AdmmLog] = None
        self._local_blocks = 0
        self._counts = None       # (uint8 counts [n, dim-1], xMean [dim], xSD [dim]): packed-count form
        self._counts_dev = None   # (cuda int32 counts, keep_idx, mean, sd, y, w, rows_local)

    def setCountMatrix(self, counts: np.ndarray, mean: np.ndarray, sd: np.ndarray):
        """The training matrix as the k-mer COUNTS it was standardised from (uint8 n x (dim-1), the
        features in column order) plus xMean / xSD (dim entries, entry 0 = 0 / 1,
        Classifier.java:362-371). execute() then goes through ``sqw_admm_set_problem_counts``: the
        trainer evaluates on X = (counts - xMean) / xSD without materialising it (packed-count
        kernels; shapes they do not cover are expanded to fp64 on the device). ``data`` may be None."""
        counts = np.ascontiguousarray(counts, dtype=np.uint8)
        self._counts = (counts, np.ascontiguousarray(mean, dtype=np.float64),
                        np.ascontiguousarray(sd, dtype=np.float64))

    def setCountMatrixDevice(self, counts, keep_idx, mean, sd, y, w, n_rows: int, rows_local: bool = False):
        """Device-resident form (cuda tensors): the dense int32 count matrix of
        ``KmerCounter.count_device`` [rows, numK], the column plan of ``prepare_device``
        (keep_idx int32 [F], mean / sd fp64 [numK] by ORIGINAL column), class ids int32 and weights
        fp64 per row. ``n_rows`` is the global row count; with ``rows_local`` the tensors hold only
        this rank's rows (see include/sequnwinder_b200.h)."""
        self._counts_dev = (counts, keep_idx, mean, sd, y, w, int(n_rows), bool(rows_local))

    # ---- setters, LOne.java:92-104 ----
    def setBGFSmaxItrs(self, m): self.BGFS_maxIts = int(m)
    def setADMMmaxItrs(self, m): self.ADMM_maxItr = int(m)
    def setSeqUnwinderMaxIts(self, m): self.NODES_maxItr = int(m)
    def setInstanceWeights(self, w): self.weights = w
    def setClsMembership(self, c): self.cls = c
    def setNumPredictors(self, p): self.numPredictors = int(p)
    def setNumClasses(self, c): self.numClasses = int(c)

    def setClassStructure(self, rel):
        self.classStructure = rel
        self.setNumNodes(rel.numNodes)

    def setNumNodes(self, n): self.numNodes = int(n)
    def setRidge(self, r): self.regularization = float(r)
    def setDebugMode(self, debug): self.sm_Debug = bool(debug)
    def setPho(self, ph): self.ADMM_pho = float(ph)
    def set_numThreads(self, nt): self.ADMM_numThreads = int(nt)

    def initUandZ(self):
        """z, zold, u start at zero (LOne.java:80-85); allocated on the device by execute()."""

    # ---- multi-GPU placement (no reference counterpart: T stays the numerics knob) ----
    def setComm(self, rank: int, world: int, nccl_id: Optional[bytes]):
        """Blocks t with t % world == rank are solved on this process's GPU; the consensus sums
        (LOne.java:395-417) become ncclAllReduce. ``nccl_id``: 128 bytes from
        ``nccl_unique_id()`` on rank 0, distributed by the caller."""
        self._rank, self._world, self._nccl_id = int(rank), int(world), nccl_id

    # ---- execute / getters ----
    def setResident(self, keep: bool = True):
        """Keep the device copy of the training matrix and the solver workspace alive between
        execute() calls (benchmarks time execute() with inputs already resident in HBM)."""
        self._resident = bool(keep)

    def setCtaBudget(self, max_ctas: int):
        """At most this many CTAs per x-update launch (0 = every SM): lets several optimizers,
        each on its own host thread, train side by side on one GPU (``sweep_concurrent``)."""
        self._cta_budget = int(max_ctas)

    def close(self):
        if self._handle is not None:
            _lib.load().sqw_admm_destroy(self._handle)
            self._handle = None

    def _open(self):
        lib = _lib.load()
        if self.classStructure is None or self.cls is None or self.weights is None:
            raise RuntimeError("setClassStructure / setClsMembership / setInstanceWeights first")
        data = None
        if self._counts_dev is not None:
            n_rows, dim = self._counts_dev[6], int(self._counts_dev[1].numel()) + 1
        elif self._counts is not None:
            n_rows, dim = self._counts[0].shape[0], self._counts[0].shape[1] + 1
        else:
            data = np.ascontiguousarray(self.data, dtype=np.float64)
            n_rows, dim = data.shape
        C_ = self.numClasses if self.numClasses is not None else int(np.max(self.cls)) + 1
        if self.numPredictors is not None and self.numPredictors + 1 != dim:
            raise ValueError("data has %d columns but numPredictors+1 = %d"
                             % (dim, self.numPredictors + 1))
        cls = np.ascontiguousarray(self.cls, dtype=np.int32)
        w = np.ascontiguousarray(self.weights, dtype=np.float64)
        st = self.classStructure.to_arrays()
        NN = st["num_nodes"]
        if self.x.size != C_ * dim or self.sm_x.size != NN * dim:
            raise ValueError("x / sm_x have the wrong length")

        def ip(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            if a.size == 0:
                a = np.zeros(1, dtype=np.int32)
            return a

        names = ("node_index", "is_leaf", "layer", "parent_ptr", "parent_idx", "child_ptr",
                 "child_idx")
        arrs = [ip(st[k]) for k in names]
        h = C.c_void_p()
        _lib.check(lib.sqw_admm_create(C.byref(h)))
        self._handle = h
        try:
            _lib.check(lib.sqw_admm_set_structure(h, NN, st["num_layers"],
                                                  *[a.ctypes.data_as(_lib.i32p) for a in arrs]))
            _lib.check(lib.sqw_admm_set_params(h, self.regularization, self.ADMM_pho,
                                               self.ADMM_numThreads, self.ADMM_maxItr,
                                               self.NODES_maxItr, int(self.sm_Debug)))
            if self._cta_budget:
                _lib.check(lib.sqw_admm_set_cta_budget(h, self._cta_budget))
            if self._world > 1:
                idbuf = (C.c_uint8 * 128).from_buffer_copy(self._nccl_id)
                _lib.check(lib.sqw_admm_set_comm(h, self._rank, self._world, idbuf))
            if self._counts_dev is not None:
                import torch
                cd, keep, mean, sd, yd, wd, _, local = self._counts_dev
                stream = C.c_void_p(torch.cuda.current_stream(cd.device).cuda_stream)
                _lib.check(lib.sqw_admm_set_problem_counts_dev(
                    h, n_rows, cd.shape[1], C_, cd.data_ptr(), keep.data_ptr(), int(keep.numel()),
                    mean.data_ptr(), sd.data_ptr(), yd.data_ptr(), wd.data_ptr(), int(local), stream))
                self.h2d_bytes = 0
            elif self._counts is not None:
                cnt, mean, sd = self._counts
                if mean.size != dim or sd.size != dim or cls.size != n_rows or w.size != n_rows:
                    raise ValueError("count matrix / statistics / labels disagree in size")
                _lib.check(lib.sqw_admm_set_problem_counts(
                    h, n_rows, dim, C_, cnt.ctypes.data_as(_lib.u8p), mean.ctypes.data_as(_lib.f64p),
                    sd.ctypes.data_as(_lib.f64p), cls.ctypes.data_as(_lib.i32p),
                    w.ctypes.data_as(_lib.f64p)))
                self.h2d_bytes = cnt.nbytes + mean.nbytes + sd.nbytes + cls.nbytes + w.nbytes
            else:
                _lib.check(lib.sqw_admm_set_problem(
                    h, n_rows, dim, C_, data.ctypes.data_as(_lib.f64p), cls.ctypes.data_as(_lib.i32p),
                    w.ctypes.data_as(_lib.f64p)))
                self.h2d_bytes = data.nbytes + cls.nbytes + w.nbytes
        except Exception:
            self.close()
            raise

    def execute(self):
        lib = _lib.load()
        if self._handle is None:
            self._open()
        h = self._handle
        try:
            x = np.ascontiguousarray(self.x, dtype=np.float64)
            smx = np.ascontiguousarray(self.sm_x, dtype=np.float64)
            rc = lib.sqw_admm_execute(h, x.ctypes.data_as(_lib.f64p), smx.ctypes.data_as(_lib.f64p))
            self._collect(lib, h)
            _lib.check(rc)
            # results are returned through the caller's arrays, as LOne mutates x / sm_x in place
            if x is not self.x:
                np.copyto(self.x, x.reshape(self.x.shape))
            if smx is not self.sm_x:
                np.copyto(self.sm_x, smx.reshape(self.sm_x.shape))
        finally:
            if not self._resident:
                self.close()

    def _collect(self, lib, h):
        stats = _lib.AdmmStats()
        _lib.check(lib.sqw_admm_get_stats(h, C.byref(stats)))
        self.stats = stats
        rows = C.c_int32(0)
        _lib.check(lib.sqw_admm_get_log(h, 0, C.byref(rows), None, None, None, None, None, None,
                                        None, None))
        r = rows.value
        T = self.ADMM_numThreads
        tl = len([t for t in range(T) if t % self._world == self._rank])
        self._local_blocks = tl
        lo, li = np.zeros(max(r, 1), np.int32), np.zeros(max(r, 1), np.int32)
        lp, lf, lpr, ld = (np.zeros(max(r, 1)) for _ in range(4))
        ln = np.zeros((max(r, 1), tl), np.int32)
        ai = np.zeros(max(1, self.NODES_maxItr), np.int32)
        _lib.check(lib.sqw_admm_get_log(
            h, r, C.byref(rows), lo.ctypes.data_as(_lib.i32p), li.ctypes.data_as(_lib.i32p),
            lp.ctypes.data_as(_lib.f64p), lf.ctypes.data_as(_lib.f64p),
            lpr.ctypes.data_as(_lib.f64p), ld.ctypes.data_as(_lib.f64p),
            ln.ctypes.data_as(_lib.i32p), ai.ctypes.data_as(_lib.i32p)))
        self.log = AdmmLog(lo[:r], li[:r], lp[:r], lf[:r], lpr[:r], ld[:r], ln[:r],
                           [int(v) for v in ai[:stats.outer_iters]])

    def _clone_settings(self, other: "LOneCuda"):
        """Copy every setter's value (not the arrays being trained) onto ``other``."""
        other.ADMM_maxItr, other.ADMM_pho = self.ADMM_maxItr, self.ADMM_pho
        other.ADMM_numThreads, other.BGFS_maxIts = self.ADMM_numThreads, self.BGFS_maxIts
        other.NODES_maxItr, other.regularization = self.NODES_maxItr, self.regularization
        other.numPredictors, other.numClasses, other.numNodes = (self.numPredictors, self.numClasses,
                                                                 self.numNodes)
        other.classStructure, other.weights, other.cls = self.classStructure, self.weights, self.cls
        other.sm_Debug = self.sm_Debug
        other._counts, other._counts_dev = self._counts, self._counts_dev

    def sweep_concurrent(self, ridges, slots: int = 4):
        """The regularisation sweep with ``slots`` trainings side by side on one GPU: ``slots``
        worker threads, each with its own optimizer whose x-update launches are capped at
        148 // slots CTAs (``setCtaBudget``), take the lambda values from a queue. A block's group
        of CTAs shrinks with the budget, and with it the part of every L-BFGS trip that is exchange
        between CTAs; measured on B200 at cfg-2, 16 lambda values: from the count matrix (rows panel
        by panel through shared memory, MODE 12) 28-31 s with 3 slots, 28-32 s with 4, 38.2 s with 6,
        against 38.2 s for ``sweep``; from the fp64 matrix 1.18x (2 slots), 1.29x (3), 1.23x (4),
        0.70x (8) against ``sweep``. Returns [(x, sm_x, stats, log)] in the order of ``ridges``; each entry
        is bit-identical to a single optimizer run with the same CTA budget, and equal to
        ``sweep``'s entry up to the rounding of a different summation order of the partial
        gradients."""
        import queue
        import threading

        if self._world > 1:
            raise RuntimeError("sweep_concurrent runs on one GPU; deal the lambda values out over "
                               "the ranks instead (parallel.local_ridges)")
        ridges = [float(r) for r in ridges]
        slots = max(1, min(int(slots), len(ridges)))
        budget = _lib.B200_SMS // slots
        x0, smx0 = np.array(self.x, dtype=np.float64), np.array(self.sm_x, dtype=np.float64)
        todo: "queue.SimpleQueue" = queue.SimpleQueue()
        for item in enumerate(ridges):
            todo.put(item)
        out: list = [None] * len(ridges)
        errors: list = []
        parent_dev = _lib.load().sqw_get_device()   # CUDA's current device is per host thread

        def worker():
            if parent_dev >= 0:
                _lib.check(_lib.load().sqw_set_device(parent_dev))
            lo = LOneCuda(x0.copy(), smx0.copy(), self.data)
            self._clone_settings(lo)
            lo.setCtaBudget(budget)
            lo.setResident(True)
            lib = _lib.load()
            try:
                while not errors:
                    try:
                        i, lam = todo.get_nowait()
                    except queue.Empty:
                        break
                    np.copyto(lo.x, x0)
                    np.copyto(lo.sm_x, smx0)
                    lo.setRidge(lam)
                    if lo._handle is not None:
                        _lib.check(lib.sqw_admm_set_ridge(lo._handle, lam))
                    lo.execute()
                    out[i] = (np.array(lo.x), np.array(lo.sm_x), lo.stats, lo.log)
            except Exception as exc:  # surfaced to the caller below
                errors.append(exc)
            finally:
                lo.close()

        threads = [threading.Thread(target=worker) for _ in range(slots)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out

    def sweep(self, ridges):
        """Regularisation sweep on one uploaded problem (BASELINE config 5): for every lambda in
        ``ridges`` run execute() from the x / sm_x this optimizer was constructed with, keeping
        the training matrix resident on the device. Returns [(x, sm_x, stats, log)] in order;
        each entry equals what a fresh LOneCuda with setRidge(lambda) returns."""
        lib = _lib.load()
        x0, smx0 = np.array(self.x, dtype=np.float64), np.array(self.sm_x, dtype=np.float64)
        keep, self._resident = self._resident, True
        out = []
        try:
            for lam in ridges:
                np.copyto(self.x, x0.reshape(self.x.shape))
                np.copyto(self.sm_x, smx0.reshape(self.sm_x.shape))
                self.setRidge(float(lam))
                if self._handle is not None:
                    _lib.check(lib.sqw_admm_set_ridge(self._handle, float(lam)))
                self.execute()
                out.append((np.array(self.x), np.array(self.sm_x), self.stats, self.log))
        finally:
            self._resident = keep
            if not keep:
                self.close()
        return out

    def getX(self): return self.x
    def getsmX(self): return self.sm_x

    def clearOptimizer(self):
        """LOne.java:259-263"""
        self.close()
        self.data = None
        self.sm_x = None
        self.x = None


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _lib.check(_lib.load().sqw_nccl_unique_id(buf))
    return bytes(buf)


def block_evaluate(data, cls, weights, structure: ClassRelationStructure, num_classes: int,
                   num_threads: int, cx: np.ndarray, sm_x: np.ndarray, pho: float, counts=None):
    """(f, g) of LOne.OptObject (LOne.java:936-1049) for every ADMM block at block weights
    ``cx[t]`` with z = u = 0: the x-update's evaluation kernel on its own, for parity tests.
    Blocks come back in the library's block order (``java_block_order``)."""
    lib = _lib.load()
    if counts is not None:   # (uint8 counts, xMean, xSD): the packed-count form of `data`
        cnt = np.ascontiguousarray(counts[0], dtype=np.uint8)
        cmean = np.ascontiguousarray(counts[1], dtype=np.float64)
        csd = np.ascontiguousarray(counts[2], dtype=np.float64)
        n_rows, dim = cnt.shape[0], cnt.shape[1] + 1
    else:
        data = np.ascontiguousarray(data, dtype=np.float64)
        n_rows, dim = data.shape
    st = structure.to_arrays()

    def ip(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return a if a.size else np.zeros(1, dtype=np.int32)

    arrs = [ip(st[k]) for k in ("node_index", "is_leaf", "layer", "parent_ptr", "parent_idx",
                                "child_ptr", "child_idx")]
    cls = np.ascontiguousarray(cls, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    cx = np.ascontiguousarray(cx, dtype=np.float64)
    sm_x = np.ascontiguousarray(sm_x, dtype=np.float64)
    f = np.zeros(num_threads)
    g = np.zeros((num_threads, num_classes * dim))
    h = C.c_void_p()
    _lib.check(lib.sqw_admm_create(C.byref(h)))
    try:
        _lib.check(lib.sqw_admm_set_structure(h, st["num_nodes"], st["num_layers"],
                                              *[a.ctypes.data_as(_lib.i32p) for a in arrs]))
        _lib.check(lib.sqw_admm_set_params(h, 0.0, pho, num_threads, 1, 1, 0))
        if counts is not None:
            _lib.check(lib.sqw_admm_set_problem_counts(
                h, n_rows, dim, num_classes, cnt.ctypes.data_as(_lib.u8p),
                cmean.ctypes.data_as(_lib.f64p), csd.ctypes.data_as(_lib.f64p),
                cls.ctypes.data_as(_lib.i32p), w.ctypes.data_as(_lib.f64p)))
        else:
            _lib.check(lib.sqw_admm_set_problem(h, n_rows, dim, num_classes,
                                                data.ctypes.data_as(_lib.f64p),
                                                cls.ctypes.data_as(_lib.i32p),
                                                w.ctypes.data_as(_lib.f64p)))
        _lib.check(lib.sqw_admm_eval(h, cx.ctypes.data_as(_lib.f64p),
                                     sm_x.ctypes.data_as(_lib.f64p), pho,
                                     f.ctypes.data_as(_lib.f64p), g.ctypes.data_as(_lib.f64p)))
    finally:
        lib.sqw_admm_destroy(h)
    return f, g


def java_block_order(T: int) -> List[int]:
    """Order in which the reference sums its per-thread buffers: iteration order of a
    java.util.HashMap keyed "Thread0".."Thread{T-1}" (LOne.java:399,411,427). The library stores
    and reports blocks in this order."""
    from .structure import _java_hashmap_order
    keys = [f"Thread{t}" for t in range(T)]
    return [int(k[6:]) for k in _java_hashmap_order(keys)]


def lbfgs_rosenbrock(x0: np.ndarray, eps: float = 0.1):
    """Runs the device L-BFGS / Moré–Thuente solver on the Rosenbrock chain (test hook)."""
    lib = _lib.load()
    x = np.array(x0, dtype=np.float64, copy=True)
    nfev, iters, iflag = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    _lib.check(lib.sqw_lbfgs_rosenbrock(x.size, x.ctypes.data_as(_lib.f64p), eps,
                                        C.byref(nfev), C.byref(iters), C.byref(iflag)))
    return x, int(iflag.value), int(nfev.value)
