"""GPU parity for the rows either side of the trainer:
counts -> standardised training matrix and batched evaluateProbability, vs the oracle."""
import numpy as np
import pytest

from helpers import make_problem
from sequnwinder_b200 import classifier as clf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,n", [("cfg1", 2000), ("cfg2", 5000), ("cfg4", 1700)])
def test_prepare_device_matches_oracle(oracle, cfg, n):
    import torch
    p = make_problem(cfg, n)
    counts = torch.from_numpy(p.counts).cuda()
    w = torch.from_numpy(p.weights).cuda()
    prep = clf.prepare_device(counts, w)
    assert np.array_equal(prep.keep_idx.cpu().numpy(), p.keep)          # same surviving columns
    mean = prep.mean.cpu().numpy()[p.keep]
    sd = prep.sd.cpu().numpy()[p.keep]
    assert np.allclose(mean, p.mean[1:], rtol=1e-13, atol=0)
    assert np.allclose(sd, p.sd[1:], rtol=1e-12, atol=0)
    X = prep.X.cpu().numpy()
    assert X.shape == p.data.shape and np.all(X[:, 0] == 1.0)
    assert np.allclose(X, p.data, rtol=1e-11, atol=1e-12)              # tolerance: fp64 rounding of mean/sd


def test_prepare_weight_rule():
    import torch
    p = make_problem("cfg1", 200, max_features=8)
    with pytest.raises(ValueError):
        clf.prepare_device(torch.from_numpy(p.counts).cuda(), torch.full((200,), 1e-3, dtype=torch.float64).cuda())


@pytest.mark.parametrize("cfg,n", [("cfg1", 2000), ("cfg2", 5000), ("cfg4", 1700)])
def test_predict_device_matches_oracle(oracle, cfg, n):
    import torch
    p = make_problem(cfg, n)
    C, dim = p.num_classes, p.data.shape[1]
    rng = np.random.default_rng(3)
    x = rng.normal(size=C * dim) * 0.1
    par = oracle.destandardize(x, C, dim, p.mean, p.sd)                 # m_Par [dim, C]
    want = oracle.predict(par, p.raw)
    prob, arg = clf.predict_device(torch.from_numpy(par).cuda(), torch.from_numpy(p.counts).cuda(),
                                   torch.from_numpy(p.keep.astype(np.int32)).cuda())
    assert np.allclose(prob.cpu().numpy(), want, rtol=1e-10, atol=1e-14)
    assert np.array_equal(arg.cpu().numpy(), want.argmax(1))           # identical arg-max predictions


def test_resident_pipeline_sequences_to_predictions(oracle):
    """sequences -> k-mer counts -> training matrix -> ADMM -> predictions without leaving the GPU
    (except the tiny weight vectors), against the same pipeline on the oracle."""
    import ctypes as C_
    import torch
    from sequnwinder_b200 import KmerCounter, _lib, structure, synth
    from sequnwinder_b200.optimizer import java_block_order
    d = synth.generate(600, 3, seed=11)
    des = structure.make_design(d.annotations)
    crs = structure.ClassRelationStructure(des.design, des.num_layers)
    cls = np.array([des.subgroup_names.index(s) for s in des.subgroups_at_sites], dtype=np.int32)
    w = np.array([des.subgroup_weights[s] for s in des.subgroups_at_sites])
    C = len(des.subgroup_names)
    kc = KmerCounter(4, 4)
    counts_t, has_n, status = kc.count_device(torch.from_numpy(d.bases).cuda(), torch.from_numpy(d.offsets).cuda(), 150)
    prep = clf.prepare_device(counts_t, torch.from_numpy(w).cuda())
    # oracle side
    counts, _ = oracle.kmer_count(d.bases, d.offsets, 4, 4)
    assert np.array_equal(counts, counts_t.cpu().numpy())
    hp = clf.prepare(counts, cls, w, crs, C)
    T, A, S = 3, 6, 2
    st = oracle.StructureArrays(**crs.to_arrays())
    xo, smo, tr = oracle.lone_execute(hp.x0, hp.sm_x0, hp.data, cls, w, st, C, num_threads=T, admm_max_itr=A, nodes_max_itr=S)
    # device side: train from the device-resident matrix through the _dev entry point
    lib = _lib.load()
    h = C_.c_void_p()
    _lib.check(lib.sqw_admm_create(C_.byref(h)))
    arr = crs.to_arrays()
    ip = lambda a: np.ascontiguousarray(a if len(a) else np.zeros(1), dtype=np.int32)
    keepalive = [ip(arr[k]) for k in ("node_index", "is_leaf", "layer", "parent_ptr", "parent_idx", "child_ptr", "child_idx")]
    _lib.check(lib.sqw_admm_set_structure(h, arr["num_nodes"], arr["num_layers"], *[a.ctypes.data_as(_lib.i32p) for a in keepalive]))
    _lib.check(lib.sqw_admm_set_params(h, 10.0, 1.7, T, A, S, 0))
    y_t = torch.from_numpy(cls).cuda()
    w_t = torch.from_numpy(w).cuda()
    n, dim = prep.X.shape
    _lib.check(lib.sqw_admm_set_problem_dev(h, n, dim, C, prep.X.data_ptr(), y_t.data_ptr(), w_t.data_ptr()))
    x, smx = hp.x0.copy(), hp.sm_x0.copy()
    _lib.check(lib.sqw_admm_execute(h, x.ctypes.data_as(_lib.f64p), smx.ctypes.data_as(_lib.f64p)))
    lib.sqw_admm_destroy(h)
    assert np.abs(x - xo).max() <= 1e-6 * np.abs(xo).max()
    # the same training straight from the device COUNT matrix + column plan (packed-count form):
    # no fp64 training matrix is materialised anywhere
    h2 = C_.c_void_p()
    _lib.check(lib.sqw_admm_create(C_.byref(h2)))
    _lib.check(lib.sqw_admm_set_structure(h2, arr["num_nodes"], arr["num_layers"], *[a.ctypes.data_as(_lib.i32p) for a in keepalive]))
    _lib.check(lib.sqw_admm_set_params(h2, 10.0, 1.7, T, A, S, 0))
    stream = C_.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.sqw_admm_set_problem_counts_dev(h2, n, counts_t.shape[1], C, counts_t.data_ptr(),
                                                   prep.keep_idx.data_ptr(), prep.num_kept, prep.mean.data_ptr(),
                                                   prep.sd.data_ptr(), y_t.data_ptr(), w_t.data_ptr(), 0, stream))
    x2, smx2 = hp.x0.copy(), hp.sm_x0.copy()
    _lib.check(lib.sqw_admm_execute(h2, x2.ctypes.data_as(_lib.f64p), smx2.ctypes.data_as(_lib.f64p)))
    st2 = _lib.AdmmStats()
    _lib.check(lib.sqw_admm_get_stats(h2, C_.byref(st2)))
    lib.sqw_admm_destroy(h2)
    assert st2.eval_mode in (8, 9)
    assert np.abs(x2 - xo).max() <= 1e-6 * np.abs(xo).max()
    assert np.abs(smx2 - smo).max() <= 1e-6 * np.abs(smo).max()
    par = clf.destandardize(x, C, dim, hp.mean, hp.sd)
    _, arg = clf.predict_device(torch.from_numpy(par).cuda(), counts_t, prep.keep_idx, want_prob=False)
    raw = np.concatenate([np.ones((n, 1)), counts[:, hp.keep]], axis=1)
    want = oracle.predict(oracle.destandardize(xo, C, dim, hp.mean, hp.sd), raw).argmax(1)
    assert np.array_equal(arg.cpu().numpy(), want)
