"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): importable from tests/, from
``__graft_entry__.smoke()`` and from bench.py's cpu_baseline / ``--impl reference`` legs.
The product package ``sequnwinder_b200`` never imports this module.

PARITY UNPINNED: the reference ships no golden vectors and cannot run in this image; the
oracle is a line-traceable restatement (citations in the C sources).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _SO


class _Structure(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("num_layers", C.c_int32),
                ("node_index", _i32p), ("is_leaf", _i32p), ("layer", _i32p),
                ("parent_ptr", _i32p), ("parent_idx", _i32p),
                ("child_ptr", _i32p), ("child_idx", _i32p)]


class _Problem(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("dim", C.c_int32), ("num_classes", C.c_int32),
                ("data", _f64p), ("cls", _i32p), ("weights", _f64p), ("st", _Structure),
                ("ridge", C.c_double), ("pho", C.c_double), ("num_threads", C.c_int32),
                ("admm_max_itr", C.c_int32), ("nodes_max_itr", C.c_int32),
                ("host_threads", C.c_int32), ("java_sum_order", C.c_int32)]


class _Trace(C.Structure):
    _fields_ = [("outer_iters", C.c_int32), ("converged", C.c_int32), ("log_cap", C.c_int32),
                ("log_rows", C.c_int32), ("log_outer", _i32p), ("log_itr", _i32p),
                ("log_pho", _f64p), ("log_fold", _f64p), ("log_primal", _f64p),
                ("log_dual", _f64p), ("log_nfev", _i32p), ("admm_iters", _i32p),
                ("final_pho", C.c_double), ("final_fold", C.c_double), ("ran_admm", C.c_int32),
                ("total_evals", C.c_int64), ("seconds_admm", C.c_double),
                ("lbfgs_errors", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_num_kmers.restype = C.c_int64
        L.orc_num_kmers.argtypes = [C.c_int, C.c_int]
        L.orc_kmer_count.restype = C.c_int
        L.orc_kmer_count.argtypes = [C.c_char_p, _i64p, C.c_int64, C.c_int, C.c_int, _i32p, _u8p]
        L.orc_lone_execute.restype = C.c_int
        L.orc_lone_execute.argtypes = [C.POINTER(_Problem), _f64p, _f64p, C.POINTER(_Trace)]
        L.orc_java_thread_order.restype = None
        L.orc_java_thread_order.argtypes = [C.c_int, _i32p]
        L.orc_objective.restype = C.c_double
        L.orc_objective.argtypes = [C.POINTER(_Problem), _f64p, _i32p, _f64p, C.c_int64, _f64p,
                                    _f64p, _f64p, _f64p, C.c_double]
        L.orc_gradient.restype = None
        L.orc_gradient.argtypes = [C.POINTER(_Problem), _f64p, _i32p, _f64p, C.c_int64, _f64p,
                                   _f64p, _f64p, _f64p, C.c_double, _f64p]
        L.orc_standardize.restype = C.c_int
        L.orc_standardize.argtypes = [_f64p, C.c_int64, C.c_int32, _f64p, _f64p, _f64p]
        L.orc_init_params.restype = None
        L.orc_init_params.argtypes = [_i32p, C.c_int64, C.c_int32, C.c_int32,
                                      C.POINTER(_Structure), _f64p, _f64p]
        L.orc_destandardize.restype = None
        L.orc_destandardize.argtypes = [_f64p, C.c_int32, C.c_int32, _f64p, _f64p, _f64p]
        L.orc_predict.restype = None
        L.orc_predict.argtypes = [_f64p, C.c_int32, C.c_int32, _f64p, C.c_int64, _f64p]
        L.orc_lbfgs_minimize.restype = C.c_int
        _lib = L
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty)


def num_kmers(kmin: int, kmax: int) -> int:
    return int(lib().orc_num_kmers(kmin, kmax))


def pack_sequences(seqs):
    """list[str|bytes] -> (concatenated uint8 array, int64 offsets[n+1])."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
    return bases, offsets


def kmer_count(bases: np.ndarray, offsets: np.ndarray, kmin: int, kmax: int):
    """Oracle for MakeArff.KmerCounter: returns (counts[n, numK] int32, has_n[n] uint8)."""
    n = len(offsets) - 1
    numk = num_kmers(kmin, kmax)
    counts = np.empty((n, numk), dtype=np.int32)
    has_n = np.empty(n, dtype=np.uint8)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    buf = bases if bases.size else np.zeros(1, dtype=np.uint8)
    rc = lib().orc_kmer_count(buf.ctypes.data_as(C.c_char_p), _p(offsets, _i64p), n, kmin, kmax,
                              _p(counts, _i32p), _p(has_n, _u8p))
    if rc == -1:
        raise ValueError("Nonstandard nucleotide character encountered")
    if rc != 0:
        raise ValueError(f"orc_kmer_count failed: {rc}")
    return counts, has_n


def hills_scan(bases: np.ndarray, offsets: np.ndarray, kmin: int, kmax: int, weights: np.ndarray,
               min_m: int, max_m: int, thresh: float, mode: int = 0, cap: int = 64):
    """Oracle for Discrim.findSeqlets (mode 0) / findHills (mode 1): returns
    (start[n, cap], length[n, cap], score[n, cap], counts[n], has_n[n])."""
    n = len(offsets) - 1
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    assert weights.size == num_kmers(kmin, kmax)
    start = np.zeros((n, cap), dtype=np.int32)
    length = np.zeros((n, cap), dtype=np.int32)
    score = np.zeros((n, cap), dtype=np.float64)
    counts = np.zeros(n, dtype=np.int32)
    has_n = np.zeros(n, dtype=np.uint8)
    buf = bases if bases.size else np.zeros(1, dtype=np.uint8)
    L = lib()
    L.orc_hills_scan.restype = C.c_int
    L.orc_hills_scan.argtypes = [C.c_char_p, _i64p, C.c_int64, C.c_int, C.c_int, _f64p, C.c_int, C.c_int,
                                 C.c_double, C.c_int, C.c_int32, _i32p, _i32p, _f64p, _i32p, _u8p]
    rc = L.orc_hills_scan(buf.ctypes.data_as(C.c_char_p), _p(offsets, _i64p), n, kmin, kmax,
                          _p(weights, _f64p), min_m, max_m, float(thresh), mode, cap, _p(start, _i32p),
                          _p(length, _i32p), _p(score, _f64p), _p(counts, _i32p), _p(has_n, _u8p))
    if rc == -1:
        raise ValueError("Nonstandard nucleotide character encountered")
    if rc == -4:
        raise OverflowError("more hills in a sequence than cap")
    if rc != 0:
        raise ValueError(f"orc_hills_scan failed: {rc}")
    return start, length, score, counts, has_n


@dataclass
class StructureArrays:
    """CSR form of a ClassRelationStructure, entries in design-file order."""
    num_nodes: int
    num_layers: int
    node_index: np.ndarray
    is_leaf: np.ndarray
    layer: np.ndarray
    parent_ptr: np.ndarray
    parent_idx: np.ndarray
    child_ptr: np.ndarray
    child_idx: np.ndarray

    def as_c(self) -> _Structure:
        for name in ("node_index", "is_leaf", "layer", "parent_ptr", "parent_idx", "child_ptr",
                     "child_idx"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.int32)
            if a.size == 0:
                a = np.zeros(1, dtype=np.int32)
            setattr(self, name, a)
        return _Structure(self.num_nodes, self.num_layers, _p(self.node_index, _i32p),
                          _p(self.is_leaf, _i32p), _p(self.layer, _i32p),
                          _p(self.parent_ptr, _i32p), _p(self.parent_idx, _i32p),
                          _p(self.child_ptr, _i32p), _p(self.child_idx, _i32p))


@dataclass
class Trace:
    outer_iters: int = 0
    converged: bool = False
    admm_iters: list = field(default_factory=list)
    log_outer: np.ndarray = None
    log_itr: np.ndarray = None
    log_pho: np.ndarray = None
    log_fold: np.ndarray = None
    log_primal: np.ndarray = None
    log_dual: np.ndarray = None
    log_nfev: np.ndarray = None
    final_pho: float = 0.0
    final_fold: float = 0.0
    ran_admm: bool = False
    total_evals: int = 0
    seconds_admm: float = 0.0
    lbfgs_errors: int = 0


def _make_problem(data, cls, weights, st: StructureArrays, num_classes, ridge, pho, num_threads,
                  admm_max_itr, nodes_max_itr, host_threads=1, java_sum_order=True):
    data = np.ascontiguousarray(data, dtype=np.float64)
    cls = np.ascontiguousarray(cls, dtype=np.int32)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    n, dim = data.shape
    prob = _Problem(n, dim, num_classes, _p(data, _f64p), _p(cls, _i32p), _p(weights, _f64p),
                    st.as_c(), ridge, pho, num_threads, admm_max_itr, nodes_max_itr, host_threads,
                    1 if java_sum_order else 0)
    return prob, (data, cls, weights, st)


def lone_execute(x, sm_x, data, cls, weights, st: StructureArrays, num_classes, ridge=10.0,
                 pho=1.7, num_threads=5, admm_max_itr=400, nodes_max_itr=15, host_threads=1,
                 java_sum_order=True):
    """Oracle for LOne.execute(): returns (x, sm_x, Trace); inputs are not modified."""
    prob, keep = _make_problem(data, cls, weights, st, num_classes, ridge, pho, num_threads,
                               admm_max_itr, nodes_max_itr, host_threads, java_sum_order)
    x = np.array(x, dtype=np.float64, copy=True)
    sm_x = np.array(sm_x, dtype=np.float64, copy=True)
    cap = max(1, admm_max_itr * nodes_max_itr)
    lo, li = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    lp, lf, lpr, ld = (np.zeros(cap) for _ in range(4))
    ln = np.zeros((cap, num_threads), np.int32)
    ai = np.zeros(max(1, nodes_max_itr), np.int32)
    tr = _Trace()
    tr.log_cap = cap
    tr.log_outer, tr.log_itr = _p(lo, _i32p), _p(li, _i32p)
    tr.log_pho, tr.log_fold = _p(lp, _f64p), _p(lf, _f64p)
    tr.log_primal, tr.log_dual = _p(lpr, _f64p), _p(ld, _f64p)
    tr.log_nfev, tr.admm_iters = _p(ln, _i32p), _p(ai, _i32p)
    rc = lib().orc_lone_execute(C.byref(prob), _p(x, _f64p), _p(sm_x, _f64p), C.byref(tr))
    del keep
    if rc != 0:
        raise RuntimeError(f"orc_lone_execute failed: {rc}")
    r = tr.log_rows
    out = Trace(outer_iters=tr.outer_iters, converged=bool(tr.converged),
                admm_iters=[int(v) for v in ai[:tr.outer_iters]],
                log_outer=lo[:r].copy(), log_itr=li[:r].copy(), log_pho=lp[:r].copy(),
                log_fold=lf[:r].copy(), log_primal=lpr[:r].copy(), log_dual=ld[:r].copy(),
                log_nfev=ln[:r].copy(), final_pho=tr.final_pho, final_fold=tr.final_fold,
                ran_admm=bool(tr.ran_admm), total_evals=int(tr.total_evals),
                seconds_admm=tr.seconds_admm, lbfgs_errors=tr.lbfgs_errors)
    return x, sm_x, out


def java_thread_order(T: int) -> np.ndarray:
    out = np.zeros(T, dtype=np.int32)
    lib().orc_java_thread_order(T, _p(out, _i32p))
    return out


def eval_fg(blk_data, blk_cls, blk_w, cx, smx, z_dense, tu_dense, pho, st: StructureArrays,
            num_classes):
    """One (objective, gradient) evaluation of LOne.OptObject on one block."""
    prob, keep = _make_problem(blk_data, blk_cls, blk_w, st, num_classes, 0.0, pho, 1, 1, 1)
    data, cls, w, _ = keep
    cx = np.ascontiguousarray(cx, dtype=np.float64)
    smx = np.ascontiguousarray(smx, dtype=np.float64)
    z = np.ascontiguousarray(z_dense, dtype=np.float64)
    tu = np.ascontiguousarray(tu_dense, dtype=np.float64)
    nb = data.shape[0]
    f = lib().orc_objective(C.byref(prob), _p(data, _f64p), _p(cls, _i32p), _p(w, _f64p), nb,
                            _p(cx, _f64p), _p(smx, _f64p), _p(z, _f64p), _p(tu, _f64p), pho)
    g = np.zeros(num_classes * data.shape[1])
    lib().orc_gradient(C.byref(prob), _p(data, _f64p), _p(cls, _i32p), _p(w, _f64p), nb,
                       _p(cx, _f64p), _p(smx, _f64p), _p(z, _f64p), _p(tu, _f64p), pho,
                       _p(g, _f64p))
    return float(f), g


_FG = C.CFUNCTYPE(C.c_double, C.c_void_p, _f64p, _f64p)


def lbfgs_minimize(fg, x0, m=5, eps=0.1, xtol=10e-16):
    """Drive the restated LBFGS/Mcsrch exactly as LOne.java:865-896 does, on a Python f/g."""
    x = np.array(x0, dtype=np.float64, copy=True)
    n = x.size

    def _cb(_user, xp, gp):
        xv = np.ctypeslib.as_array(xp, shape=(n,))
        f, g = fg(xv.copy())
        np.ctypeslib.as_array(gp, shape=(n,))[:] = g
        return float(f)

    cb = _FG(_cb)
    nfev, iters = C.c_int(0), C.c_int(0)
    fn = lib().orc_lbfgs_minimize
    fn.argtypes = [C.c_int, C.c_int, _f64p, C.c_double, C.c_double, _FG, C.c_void_p,
                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    iflag = fn(n, m, _p(x, _f64p), eps, xtol, cb, None, C.byref(nfev), C.byref(iters))
    return x, int(iflag), int(nfev.value), int(iters.value)


def standardize(data, weights):
    data = np.array(data, dtype=np.float64, copy=True, order="C")
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    n, dim = data.shape
    mean, sd = np.zeros(dim), np.zeros(dim)
    rc = lib().orc_standardize(_p(data, _f64p), n, dim, _p(weights, _f64p), _p(mean, _f64p),
                               _p(sd, _f64p))
    if rc != 0:
        raise ValueError("Sum of weights of instances less than 1, please reweight!")
    return data, mean, sd


def init_params(cls, dim, num_classes, st: StructureArrays):
    cls = np.ascontiguousarray(cls, dtype=np.int32)
    x = np.zeros(num_classes * dim)
    smx = np.zeros(st.num_nodes * dim)
    cst = st.as_c()
    lib().orc_init_params(_p(cls, _i32p), len(cls), dim, num_classes, C.byref(cst), _p(x, _f64p),
                          _p(smx, _f64p))
    return x, smx


def destandardize(v, ncols, dim, mean, sd):
    v = np.ascontiguousarray(v, dtype=np.float64)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    sd = np.ascontiguousarray(sd, dtype=np.float64)
    par = np.zeros((dim, ncols))
    lib().orc_destandardize(_p(v, _f64p), ncols, dim, _p(mean, _f64p), _p(sd, _f64p),
                            _p(par, _f64p))
    return par


def predict(par, data_raw):
    par = np.ascontiguousarray(par, dtype=np.float64)
    data_raw = np.ascontiguousarray(data_raw, dtype=np.float64)
    dim, ncls = par.shape
    n = data_raw.shape[0]
    prob = np.zeros((n, ncls))
    lib().orc_predict(_p(par, _f64p), dim, ncls, _p(data_raw, _f64p), n, _p(prob, _f64p))
    return prob
