"""GPU parity, path 2: the device trainer (through the C ABI / LOneCuda) vs the CPU oracle.

Tolerance: north_star asks for weights within 1e-6 relative after a fixed iteration count. The
reference algorithm itself is chaotic beyond a handful of ADMM iterations at the benchmark shapes
(every x-update stops at ||g||/max(1,||x||) <= 0.1, LOne.java:874, so its result depends on the
path; a one-ulp change of the summation order in the ORACLE ALONE grows to 4e-5 relative after 20
ADMM iterations at cfg-1). The fixed counts below are inside the window in
which 1e-6 is meaningful; full-length runs are compared through trajectory-independent
properties instead."""
import numpy as np
import pytest

from helpers import make_problem
from sequnwinder_b200 import optimizer as opt

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6   # north_star: trained weights within 1e-6 relative (fp64)


def packed_counts(p):
    """The training matrix as the trainer's packed-count input: uint8 counts of the kept k-mers +
    xMean / xSD (Classifier.java:362-371)."""
    return p.counts[:, p.keep].astype(np.uint8), p.mean, p.sd


def run_gpu(p, T, A, S, ridge=10.0, pho=1.7, packed=False):
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None if packed else p.data)
    if packed:
        lo.setCountMatrix(*packed_counts(p))
    lo.setRidge(ridge)
    lo.setADMMmaxItrs(A)
    lo.setBGFSmaxItrs(-1)
    lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs)
    lo.setClsMembership(p.cls)
    lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes)
    lo.setNumPredictors(p.data.shape[1] - 1)
    lo.setPho(pho)
    lo.set_numThreads(T)
    lo.initUandZ()
    lo.setDebugMode(False)
    lo.execute()
    return lo


def rosen(x):
    a, b = x[0::2], x[1::2]
    t1 = b - a * a
    g = np.zeros_like(x)
    g[0::2] = -400 * a * t1 - 2 * (1 - a)
    g[1::2] = 200 * t1
    return float(np.sum(100 * t1 * t1 + (1 - a) ** 2)), g


@pytest.mark.parametrize("n", [2, 10, 256, 2000])
@pytest.mark.parametrize("eps", [0.1, 1e-6])
def test_device_lbfgs_matches_oracle_on_rosenbrock(oracle, n, eps):
    x0 = np.tile([-1.2, 1.0], n // 2)
    xr, iflag, nfev, _ = oracle.lbfgs_minimize(rosen, x0, eps=eps)
    xg, iflag_g, nfev_g = opt.lbfgs_rosenbrock(x0, eps)
    assert (iflag_g, nfev_g) == (iflag, nfev)          # same number of f/g evaluations
    assert np.abs(xr - xg).max() <= 1e-8


@pytest.mark.parametrize("cfg,n,mf,T", [("cfg1", 500, 60, 5), ("cfg1", 2000, None, 5),
                                          ("cfg2", 4000, None, 8), ("cfg4", 1700, 200, 4)])
def test_evaluation_kernel_matches_oracle(oracle, cfg, n, mf, T):
    """f and g of LOne.OptObject per block: C = 3, 8 and 17 (one, one and three class tiles; the
    last with the parentless Random leaf)."""
    p = make_problem(cfg, n, max_features=mf)
    nrow, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    rng = np.random.default_rng(5)
    cx = rng.normal(size=(T, C * dim)) * 0.05
    smx = rng.normal(size=NN * dim) * 0.05
    f, g = opt.block_evaluate(p.data, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7)
    nb = nrow // T
    zd = np.zeros(NN * NN * dim)
    for b, t in enumerate(opt.java_block_order(T)):
        rows = np.arange(nb) * T + t
        fo, go = oracle.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7,
                                p.st, C)
        assert abs(f[b] - fo) <= 1e-12 * abs(fo)
        assert np.abs(g[b] - go).max() <= 1e-12 * np.abs(go).max()


@pytest.mark.parametrize("cfg,n,mf,T,A,S", [
    ("cfg1", 500, 60, 5, 50, 3),        # well-conditioned: tracks the oracle for whole runs
    ("cfg1", 2000, None, 5, 5, 1),      # cfg-1 at full shape, fixed iteration count
    ("cfg2", 4000, None, 8, 4, 1),      # cfg-2 shape (8 labels, F=648), reduced rows
    ("cfg4", 1700, 200, 4, 6, 2),       # 17 leaves incl. parentless Random (self edge path)
    ("cfg1", 301, 24, 7, 8, 2),         # N mod T != 0: trailing rows dropped
])
def test_trainer_matches_oracle_fixed_iterations(oracle, cfg, n, mf, T, A, S):
    p = make_problem(cfg, n, max_features=mf)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    lo = run_gpu(p, T, A, S)
    xg, smg = lo.getX(), lo.getsmX()
    assert lo.stats.outer_iters == tr.outer_iters
    assert lo.log.admm_iters == tr.admm_iters
    order = opt.java_block_order(T)
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, order])     # identical L-BFGS evaluation counts
    assert np.array_equal(lo.log.pho, tr.log_pho) and np.array_equal(lo.log.fold, tr.log_fold)
    assert np.abs(xg - xo).max() <= REL_TOL * np.abs(xo).max()
    assert np.abs(smg - smo).max() <= REL_TOL * np.abs(smo).max()
    assert np.abs(lo.log.primal - tr.log_primal).max() <= REL_TOL * np.abs(tr.log_primal).max()
    # identical per-site arg-max predictions (Classifier.java:267-287)
    C, dim = p.num_classes, p.data.shape[1]
    pg = oracle.predict(oracle.destandardize(xg, C, dim, p.mean, p.sd), p.raw).argmax(1)
    po = oracle.predict(oracle.destandardize(xo, C, dim, p.mean, p.sd), p.raw).argmax(1)
    assert np.array_equal(pg, po)


@pytest.mark.parametrize("cfg,n,T,A", [("cfg3", 2400, 8, 2),     # k=4..7: dim 10921, C=12 (column-chunked mode)
                                         ("cfg4", 3400, 8, 2)])    # k=4..6: dim 2729, C=17 incl. Random leaf
def test_long_rows_match_oracle(oracle, cfg, n, T, A):
    """cfg-3 / cfg-4 feature shapes (rows far longer than shared memory) at reduced row counts."""
    p = make_problem(cfg, n)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=1, host_threads=8)
    lo = run_gpu(p, T, A, 1)
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert np.abs(lo.getX() - xo).max() <= REL_TOL * np.abs(xo).max()
    assert np.abs(lo.getsmX() - smo).max() <= REL_TOL * np.abs(smo).max()


@pytest.mark.parametrize("mode", ["6", "1", "2"])
def test_fallback_evaluation_modes_match_oracle(oracle, mode, monkeypatch):
    """The more general evaluation paths (column-chunked TMA tiles, streaming) on a cfg-2 shape."""
    monkeypatch.setenv("SQW_ADMM_EVAL_MODE", mode)
    p = make_problem("cfg2", 4000)
    T, A = 8, 3
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=1, host_threads=8)
    lo = run_gpu(p, T, A, 1)
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert np.abs(lo.getX() - xo).max() <= REL_TOL * np.abs(xo).max()


def test_rho_adaptation_and_convergence_paths(oracle):
    """A run long enough to double rho, converge an ADMM run early and finish the outer loop."""
    p = make_problem("cfg1", 400, max_features=200)
    T, A, S = 2, 400, 15
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S)
    lo = run_gpu(p, T, A, S)
    assert len(set(tr.log_pho.tolist())) > 1 and min(tr.admm_iters) < A     # the case exercises both
    assert lo.stats.outer_iters == tr.outer_iters and bool(lo.stats.converged) == tr.converged
    assert lo.log.admm_iters == tr.admm_iters
    assert np.array_equal(lo.log.pho, tr.log_pho)
    assert bool(lo.stats.ran_admm) == tr.ran_admm
    assert np.abs(lo.getX() - xo).max() <= REL_TOL * np.abs(xo).max()
    assert np.abs(lo.getsmX() - smo).max() <= REL_TOL * np.abs(smo).max()


def test_full_size_cfg1_properties(oracle):
    """cfg-1 at full size, run to the reference's tolerances: trajectory-independent checks."""
    p = make_problem("cfg1")
    T = 5
    lo = run_gpu(p, T, 400, 15)
    x, smx = lo.getX(), lo.getsmX()
    C, dim = p.num_classes, p.data.shape[1]
    assert np.isfinite(x).all() and np.isfinite(smx).all()
    assert lo.stats.converged == 1 and 2 <= lo.stats.outer_iters <= 15
    # leaves of sm_x are x-bar (LOne.java:753-758); labels are lower medians of their children
    assert np.array_equal(smx[:C * dim], x)
    S = smx.reshape(-1, dim)
    for lab in p.crs.layers[1]:
        kids = np.sort(S[lab.children][:, 1:], axis=0)
        k = len(lab.children)
        med = kids[k // 2 - 1] if k % 2 == 0 else kids[k // 2]
        assert np.array_equal(S[lab.nodeIndex, 1:], med)
    # the model learned something: training accuracy well above chance (1/3)
    pred = oracle.predict(oracle.destandardize(x, C, dim, p.mean, p.sd), p.raw).argmax(1)
    assert (pred == p.cls).mean() > 0.45
    # every logged x-update did at least 2 evaluations (no pre-check, LBFGS.java:567-572)
    assert lo.log.nfev.min() >= 2


def test_ridge_sweep_equals_independent_runs(oracle):
    """BASELINE config 5 (regularisation sweep on one matrix): sweep() re-uses the resident
    training matrix; every entry must equal a fresh optimizer's run bit for bit, and the oracle's
    run at that lambda within the stated tolerance."""
    p = make_problem("cfg1", 500, max_features=60)
    T, A, S = 5, 30, 2
    ridges = [0.1, 1.0, 10.0, 100.0]
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setADMMmaxItrs(A); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1)
    lo.setPho(1.7); lo.set_numThreads(T); lo.initUandZ()
    res = lo.sweep(ridges)
    assert lo._handle is None                      # not resident: the handle is released
    norms = []
    for lam, (x, smx, stats, log) in zip(ridges, res):
        one = run_gpu(p, T, A, S, ridge=lam)
        assert np.array_equal(x, one.getX()) and np.array_equal(smx, one.getsmX())
        assert log.admm_iters == one.log.admm_iters
        xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                          ridge=lam, num_threads=T, admm_max_itr=A, nodes_max_itr=S,
                                          host_threads=8)
        assert log.admm_iters == tr.admm_iters
        assert np.abs(x - xo).max() <= REL_TOL * np.abs(xo).max()
        assert np.abs(smx - smo).max() <= REL_TOL * np.abs(smo).max()
        norms.append(np.abs(smx).sum())
    assert norms[0] > norms[-1]                    # stronger L1 penalty, smaller label weights


def test_concurrent_sweep_matches_budgeted_runs_and_oracle(oracle):
    """sweep_concurrent: several trainings side by side on one GPU, each capped at a share of the
    SMs. Every entry must equal a single run with the same CTA budget bit for bit (the schedule is
    a function of evaluation counts, not of timing or of what runs next to it), and the oracle's
    run at that lambda within the stated tolerance."""
    p = make_problem("cfg1", 500, max_features=60)
    T, A, S = 5, 30, 2
    ridges = [0.1, 1.0, 10.0, 100.0, 3.0]
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setADMMmaxItrs(A); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1)
    lo.setPho(1.7); lo.set_numThreads(T); lo.initUandZ()
    res = lo.sweep_concurrent(ridges, slots=2)
    assert all(r is not None for r in res)
    for lam, (x, smx, stats, log) in zip(ridges, res):
        one = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
        lo._clone_settings(one)
        one.setRidge(lam)
        one.setCtaBudget(148 // 2)
        one.execute()
        assert np.array_equal(x, one.getX()) and np.array_equal(smx, one.getsmX())
        assert log.admm_iters == one.log.admm_iters
        xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                          ridge=lam, num_threads=T, admm_max_itr=A, nodes_max_itr=S,
                                          host_threads=8)
        assert log.admm_iters == tr.admm_iters
        assert np.abs(x - xo).max() <= REL_TOL * np.abs(xo).max()
        assert np.abs(smx - smo).max() <= REL_TOL * np.abs(smo).max()


def test_fused_iteration_tail_is_bit_identical(monkeypatch):
    """Single GPU: pack + consensus + decide run as one launch (admm_consensus_fused_kernel); the
    three-kernel tail (the multi-GPU path, SQW_ADMM_UNFUSED here) must give the same bits."""
    p = make_problem("cfg1", 800, max_features=100)
    monkeypatch.delenv("SQW_ADMM_UNFUSED", raising=False)
    fused = run_gpu(p, 5, 25, 3)
    monkeypatch.setenv("SQW_ADMM_UNFUSED", "1")
    plain = run_gpu(p, 5, 25, 3)
    assert np.array_equal(fused.getX(), plain.getX()) and np.array_equal(fused.getsmX(), plain.getsmX())
    assert np.array_equal(fused.log.pho, plain.log.pho) and np.array_equal(fused.log.primal, plain.log.primal)
    assert np.array_equal(fused.log.nfev, plain.log.nfev)
    assert fused.stats.launches < plain.stats.launches


def test_errors_are_reported():
    p = make_problem("cfg1", 100, max_features=8)
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
    lo.setClassStructure(p.crs)
    lo.setClsMembership(p.cls)
    lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes)
    lo.set_numThreads(1000)              # more blocks than rows
    with pytest.raises(RuntimeError):
        lo.execute()


def test_two_gpu_sharding_matches_oracle():
    """T fixed, blocks spread over 2 ranks (ncclAllReduce consensus), both iteration tails (one
    collective with the expanded primal residual; two collectives): same iteration counts, weights
    and residual logs as the oracle, bitwise-identical replicated state on both ranks, and the
    expanded residual within 1e-8 of the direct one. Needs 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered on CPU by tests/test_parallel_gloo.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29533", os.path.join(root, "tools", "gpu_multi_check.py")],
                       cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ranks bitwise equal True") == 6


@pytest.mark.parametrize("slots", [1, 2])
def test_cross_validation_accuracy_matches_oracle(oracle, slots):
    """north_star: identical per-site arg-max predictions and cross-validation accuracy. The
    model on all sites + one model per fold through Classifier.buildClassifier (RemoveUseless,
    standardisation and the null-model start recomputed on every training subset, as Weka's
    crossValidateModel does), on identical fold index lists, against the same loop on the oracle;
    slots = 2 trains the models side by side on one GPU."""
    from sequnwinder_b200 import classifier as clf
    p = make_problem("cfg1", 600)
    counts = p.counts[:, p.keep[:80]]
    T, A, S = 5, 8, 2
    folds = clf.stratified_folds(p.cls, 3, seed=1)
    cv = clf.cross_validate(p.crs, counts, p.cls, p.weights, p.num_classes, folds, slots=slots,
                            ridge=10.0, pho=1.7, num_threads=T, admm_max_its=A,
                            sequnwinder_max_its=S)
    C = p.num_classes
    want = np.full(len(p.cls), -1, dtype=np.int64)
    for f, test in enumerate(folds):
        train = np.setdiff1d(np.arange(len(p.cls)), test)
        ct = counts[train]
        keep = np.nonzero(ct.max(0) != ct.min(0))[0]
        raw = np.concatenate([np.ones((train.size, 1)), ct[:, keep].astype(np.float64)], 1)
        data, mean, sd = oracle.standardize(raw, p.weights[train])
        x0, smx0 = oracle.init_params(p.cls[train], raw.shape[1], C, p.st)
        xo, smo, tr = oracle.lone_execute(x0, smx0, data, p.cls[train], p.weights[train], p.st, C,
                                          ridge=10.0, num_threads=T, admm_max_itr=A,
                                          nodes_max_itr=S, host_threads=8)
        assert np.array_equal(cv.fold_models[f].prep.keep, keep)
        assert np.abs(cv.fold_models[f].x - xo).max() <= REL_TOL * np.abs(xo).max()
        par = oracle.destandardize(xo, C, raw.shape[1], mean, sd)
        raw_test = np.concatenate([np.ones((test.size, 1)), counts[test][:, keep].astype(np.float64)], 1)
        want[test] = oracle.predict(par, raw_test).argmax(1)
    assert np.array_equal(cv.predictions, want)                    # identical arg-max per site
    assert cv.accuracy == float((want == p.cls).mean())            # identical CV accuracy
    assert cv.full_model.m_Par.shape == (81, C)


# --------------------------------------------------------------------------------------------
# The path the benchmark runs: full cfg-2 (20 000 x 649, T = 8, 2 500 rows per block). With 18 CTAs
# per block a CTA walks more tiles per evaluation than its TMA ring holds, so the ring STREAMS
# (refill / stage_wait / drain non-resident branches) - the reduced shapes above keep every CTA's
# rows resident. VERDICT r1 "What's weak" #1.
# --------------------------------------------------------------------------------------------
def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def cfg2_full():
    return make_problem("cfg2")


def test_full_cfg2_one_iteration_matches_oracle(oracle, cfg2_full):
    """A = 1, S = 1 at the headline shape: weights <= 1e-6, identical evaluation counts, and the
    streaming branch of the tile ring was the one that ran."""
    p = cfg2_full
    assert p.data.shape == (20000, 649)
    T = 8
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=1, nodes_max_itr=1, host_threads=8)
    lo = run_gpu(p, T, 1, 1)
    assert lo.stats.eval_mode in (4, 5)
    assert lo.stats.tiles_per_cta > lo.stats.ring_stages          # not resident: the ring streams
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL
    assert _rel(lo.getsmX(), smo) <= REL_TOL


@pytest.fixture(scope="module")
def cfg2_controls(oracle, cfg2_full):
    """Oracle runs of k = 1..3 ADMM iterations at full cfg-2 and, as the control, the same runs
    with ONE matrix element moved by one ulp (two different elements): how far the reference
    algorithm drifts from itself under a rounding-sized perturbation."""
    p = cfg2_full
    T = 8
    perturbed = []
    for (i, j) in ((123, 45), (15000, 600)):
        d = p.data.copy()
        d[i, j] = np.nextafter(d[i, j], np.inf)
        perturbed.append(d)
    out = {}
    for k in (1, 2, 3):
        xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st,
                                          p.num_classes, num_threads=T, admm_max_itr=k,
                                          nodes_max_itr=1, host_threads=8)
        ctrl_err, ctrl_same = 0.0, True
        for d in perturbed:
            xc, _, trc = oracle.lone_execute(p.x0, p.smx0, d, p.cls, p.weights, p.st, p.num_classes,
                                             num_threads=T, admm_max_itr=k, nodes_max_itr=1,
                                             host_threads=8)
            ctrl_err = max(ctrl_err, _rel(xc, xo))
            ctrl_same = ctrl_same and np.array_equal(trc.log_nfev, tr.log_nfev)
        out[k] = (xo, tr, ctrl_err, ctrl_same)
    return out


@pytest.mark.parametrize("packed", [False, True])
def test_full_cfg2_tracks_oracle_like_a_one_ulp_perturbation(cfg2_full, cfg2_controls, packed):
    """DESIGN.md 4 as a test. The reference algorithm amplifies rounding differences (every
    x-update stops at ||g||/max(1,||x||) <= 0.1, LOne.java:874): the oracle ALONE, re-run with ONE
    matrix element moved by one ulp, drifts away from itself by orders of magnitude per ADMM
    iteration. The device paths (fp64 tiles; packed counts) differ from the oracle in rounding
    only, so after k iterations their distance from the oracle must stay within a small factor of
    that control, the evaluation counts must be identical for as long as the control's are, and
    after ONE iteration the weights agree to 1e-6 (north_star's tolerance) with identical counts."""
    p = cfg2_full
    T = 8
    order = opt.java_block_order(T)
    for k in (1, 2, 3):
        xo, tr, ctrl_err, ctrl_same = cfg2_controls[k]
        lo = run_gpu(p, T, k, 1, packed=packed)
        assert (lo.stats.eval_mode in (8, 9)) == packed
        err = _rel(lo.getX(), xo)
        same = np.array_equal(lo.log.nfev, tr.log_nfev[:, order])
        print(f"cfg2 full ({'packed' if packed else 'fp64'}), {k} ADMM iterations: device vs oracle "
              f"{err:.2e} (nfev identical: {same}); oracle vs one-ulp-perturbed oracle {ctrl_err:.2e} "
              f"(nfev identical: {ctrl_same})")
        assert np.array_equal(lo.log.nfev[:1], tr.log_nfev[:1, order])   # iteration 0
        if ctrl_same:
            assert same
        assert err <= max(1e-9, 30.0 * ctrl_err)
        if k == 1:
            assert err <= REL_TOL


def test_mode6_streams_more_tiles_than_ring_stages(oracle):
    """Column-chunked path (cfg-3 feature shape, C = 12) with more tiles per CTA and chunk than
    the ring has stages."""
    p = make_problem("cfg3", 4800, max_features=3000)
    T = 8
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=1, nodes_max_itr=1, host_threads=8)
    lo = run_gpu(p, T, 1, 1)
    assert lo.stats.eval_mode == 6
    assert lo.stats.tiles_per_cta > lo.stats.ring_stages
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL
    assert _rel(lo.getsmX(), smo) <= REL_TOL


def test_two_live_handles_of_different_shapes():
    """the dynamic shared memory attribute is per function and process-wide; a second
    handle with a smaller layout must not break the first one's launches."""
    p1 = make_problem("cfg1", 600)                      # dim 649
    p2 = make_problem("cfg1", 400, max_features=40)     # dim 41
    los = []
    for p in (p1, p2):
        lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), p.data)
        lo.setRidge(10.0); lo.setADMMmaxItrs(3); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(1)
        lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
        lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1)
        lo.setPho(1.7); lo.set_numThreads(5); lo.initUandZ(); lo.setResident(True)
        los.append(lo)
    los[0].execute(); first = los[0].getX().copy()
    los[1].execute()
    np.copyto(los[0].x, p1.x0); np.copyto(los[0].sm_x, p1.smx0)
    los[0].execute()                                    # after the smaller handle was prepared
    assert np.array_equal(los[0].getX(), first)
    for lo in los:
        lo.close()


# --------------------------------------------------------------------------------------------
# Packed-count form (sqw_admm_set_problem_counts, csrc/admm_eval_packed.cuh): the trainer is
# handed the uint8 k-mer counts + xMean / xSD instead of the standardised doubles and evaluates on
# X = (n - xMean) / xSD without materialising it. The oracle always runs the reference's way, on
# the standardised fp64 matrix (Classifier.java:373-381 then LOne.java:936-1049).
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,n,mf,T", [("cfg1", 500, 60, 5), ("cfg1", 2000, None, 5),
                                          ("cfg2", 4000, None, 8), ("cfg2", 20000, None, 8),
                                          ("cfg1", 301, 24, 7)])
def test_packed_evaluation_matches_oracle(oracle, cfg, n, mf, T):
    p = make_problem(cfg, n, max_features=mf)
    nrow, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    rng = np.random.default_rng(5)
    cx = rng.normal(size=(T, C * dim)) * 0.05
    smx = rng.normal(size=NN * dim) * 0.05
    f, g = opt.block_evaluate(None, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7, counts=packed_counts(p))
    nb = nrow // T
    zd = np.zeros(NN * NN * dim)
    for b, t in enumerate(opt.java_block_order(T)):
        rows = np.arange(nb) * T + t
        fo, go = oracle.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7,
                                p.st, C)
        assert abs(f[b] - fo) <= 1e-12 * abs(fo)
        assert np.abs(g[b] - go).max() <= 1e-12 * np.abs(go).max()


@pytest.mark.parametrize("cfg,n,mf,T,A,S", [
    ("cfg1", 500, 60, 5, 50, 3),
    ("cfg1", 2000, None, 5, 5, 1),
    ("cfg2", 4000, None, 8, 4, 1),
    ("cfg1", 301, 24, 7, 8, 2),
    ("cfg2", 20000, None, 8, 1, 1),     # the benchmark's shape: 72 (MODE 9) / 144 (MODE 8) rows per CTA resident
])
@pytest.mark.parametrize("two_ctas_per_sm", [False, True])
def test_packed_trainer_matches_oracle_fixed_iterations(oracle, monkeypatch, cfg, n, mf, T, A, S, two_ctas_per_sm):
    p = make_problem(cfg, n, max_features=mf)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    if two_ctas_per_sm:
        monkeypatch.setenv("SQW_ADMM_PACKED_MODE", "9")
    lo = run_gpu(p, T, A, S, packed=True)
    # the packed kernels ran: MODE 8 (one CTA per SM), or the opt-in MODE 9 (two CTAs per SM)
    assert lo.stats.eval_mode == (9 if two_ctas_per_sm else 8)
    xg, smg = lo.getX(), lo.getsmX()
    assert lo.stats.outer_iters == tr.outer_iters
    assert lo.log.admm_iters == tr.admm_iters
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert np.array_equal(lo.log.pho, tr.log_pho) and np.array_equal(lo.log.fold, tr.log_fold)
    assert np.abs(xg - xo).max() <= REL_TOL * np.abs(xo).max()
    assert np.abs(smg - smo).max() <= REL_TOL * np.abs(smo).max()
    C, dim = p.num_classes, p.data.shape[1]
    pg = oracle.predict(oracle.destandardize(xg, C, dim, p.mean, p.sd), p.raw).argmax(1)
    po = oracle.predict(oracle.destandardize(xo, C, dim, p.mean, p.sd), p.raw).argmax(1)
    assert np.array_equal(pg, po)


def test_counts_with_a_cta_budget_use_the_streamed_packed_kernels(oracle):
    """A handle with a CTA budget (concurrent sweeps) has too few CTAs per block for its rows to stay
    in shared memory: the count matrix passes through the resident form's buffer panel by panel
    (MODE 12; 250 rows per CTA here = two panels)."""
    p = make_problem("cfg2", 4000)
    T, A, S = 8, 3, 1
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None)
    lo.setCountMatrix(*packed_counts(p))
    lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1)
    lo.setPho(1.7); lo.set_numThreads(T); lo.initUandZ(); lo.setCtaBudget(16)
    lo.execute()
    assert lo.stats.eval_mode == 12 and lo.stats.ctas_per_block == 2
    assert lo.stats.tiles_per_cta > lo.stats.ring_stages      # more rows per CTA than one panel holds
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL


# --------------------------------------------------------------------------------------------
# MODE 12: MODE 8's inner loops on rows that pass through shared memory panel by panel. Forced on
# small shapes with SQW_ADMM_PACKED_MODE=12 and a small panel (SQW_ADMM_PANEL_ROWS) so that a CTA
# sees several panels, the last one partial.
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,n,mf,T,panel", [("cfg1", 500, 60, 5, 8), ("cfg2", 4000, None, 8, 8),
                                                ("cfg2", 4000, None, 8, 24), ("cfg1", 301, 24, 7, 8),
                                                ("cfg2", 20000, None, 8, 48)])
def test_packed_panel_evaluation_matches_oracle(oracle, monkeypatch, cfg, n, mf, T, panel):
    monkeypatch.setenv("SQW_ADMM_PACKED_MODE", "12")
    monkeypatch.setenv("SQW_ADMM_PANEL_ROWS", str(panel))
    p = make_problem(cfg, n, max_features=mf)
    nrow, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    rng = np.random.default_rng(5)
    cx = rng.normal(size=(T, C * dim)) * 0.05
    smx = rng.normal(size=NN * dim) * 0.05
    # (the evaluation kernel runs three evaluations and returns NaN unless they agree bit for bit:
    # panel re-loads and the prefetch of panel 0 behind an evaluation are part of the check)
    f, g = opt.block_evaluate(None, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7, counts=packed_counts(p))
    nb = nrow // T
    zd = np.zeros(NN * NN * dim)
    for b, t in enumerate(opt.java_block_order(T)):
        rows = np.arange(nb) * T + t
        fo, go = oracle.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7,
                                p.st, C)
        assert abs(f[b] - fo) <= 1e-12 * abs(fo)
        assert np.abs(g[b] - go).max() <= 1e-12 * np.abs(go).max()


@pytest.mark.parametrize("cfg,n,mf,T,A,S,panel,budget", [
    ("cfg1", 500, 60, 5, 50, 3, 8, 0),
    ("cfg2", 4000, None, 8, 4, 1, 8, 0),
    ("cfg1", 301, 24, 7, 8, 2, 8, 0),
    ("cfg2", 20000, None, 8, 1, 1, 0, 49),     # a slot of a three-way concurrent sweep at the benchmark's shape
    ("cfg2", 20000, None, 8, 1, 1, 0, 24),     # ... of a six-way one: 834 rows per CTA, six panels
])
def test_packed_panel_trainer_matches_oracle(oracle, monkeypatch, cfg, n, mf, T, A, S, panel, budget):
    if panel:
        monkeypatch.setenv("SQW_ADMM_PACKED_MODE", "12")
        monkeypatch.setenv("SQW_ADMM_PANEL_ROWS", str(panel))
    p = make_problem(cfg, n, max_features=mf)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    lo = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None)
    lo.setCountMatrix(*packed_counts(p))
    lo.setRidge(10.0); lo.setADMMmaxItrs(A); lo.setBGFSmaxItrs(-1); lo.setSeqUnwinderMaxIts(S)
    lo.setClassStructure(p.crs); lo.setClsMembership(p.cls); lo.setInstanceWeights(p.weights)
    lo.setNumClasses(p.num_classes); lo.setNumPredictors(p.data.shape[1] - 1)
    lo.setPho(1.7); lo.set_numThreads(T); lo.initUandZ()
    if budget:
        lo.setCtaBudget(budget)
    lo.execute()
    assert lo.stats.eval_mode == 12
    if budget:
        assert lo.stats.tiles_per_cta > lo.stats.ring_stages
    assert lo.stats.outer_iters == tr.outer_iters
    assert lo.log.admm_iters == tr.admm_iters
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL
    assert _rel(lo.getsmX(), smo) <= REL_TOL


def test_packed_ridge_sweep_and_resident_reuse(oracle):
    """A resident packed-count optimizer re-run (sweep) returns what fresh optimizers return."""
    p = make_problem("cfg1", 600, max_features=50)
    ridges = [0.5, 10.0]
    base = opt.LOneCuda(p.x0.copy(), p.smx0.copy(), None)
    base.setCountMatrix(*packed_counts(p))
    base.setADMMmaxItrs(6); base.setBGFSmaxItrs(-1); base.setSeqUnwinderMaxIts(2)
    base.setClassStructure(p.crs); base.setClsMembership(p.cls); base.setInstanceWeights(p.weights)
    base.setNumClasses(p.num_classes); base.setNumPredictors(p.data.shape[1] - 1)
    base.setPho(1.7); base.set_numThreads(5); base.initUandZ()
    out = base.sweep(ridges)
    for lam, (x, smx, stats, log) in zip(ridges, out):
        assert stats.eval_mode in (8, 9)
        fresh = run_gpu(p, 5, 6, 2, ridge=lam, packed=True)
        assert np.array_equal(x, fresh.getX()) and np.array_equal(smx, fresh.getsmX())
        xo, _, _ = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                       ridge=lam, num_threads=5, admm_max_itr=6, nodes_max_itr=2)
        assert _rel(x, xo) <= REL_TOL


# --------------------------------------------------------------------------------------------
# MODE 10: packed counts streamed from HBM, column-chunked (csrc/admm_eval_packed_chunk.cuh) - what
# count matrices run through when their rows do not stay in shared memory: long rows, more than 8
# classes (cfg-3 / cfg-4 shapes), or forced here on small shapes (SQW_ADMM_PACKED_MODE=10).
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,n,mf,T", [("cfg1", 500, 60, 5), ("cfg2", 4000, None, 8),
                                          ("cfg4", 1700, 200, 4), ("cfg3", 1200, 1100, 4),
                                          ("cfg4", 3400, None, 8)])
def test_packed_chunk_evaluation_matches_oracle(oracle, monkeypatch, cfg, n, mf, T):
    monkeypatch.setenv("SQW_ADMM_PACKED_MODE", "10")
    p = make_problem(cfg, n, max_features=mf)
    nrow, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    rng = np.random.default_rng(5)
    cx = rng.normal(size=(T, C * dim)) * 0.05
    smx = rng.normal(size=NN * dim) * 0.05
    f, g = opt.block_evaluate(None, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7, counts=packed_counts(p))
    nb = nrow // T
    zd = np.zeros(NN * NN * dim)
    for b, t in enumerate(opt.java_block_order(T)):
        rows = np.arange(nb) * T + t
        fo, go = oracle.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7,
                                p.st, C)
        assert abs(f[b] - fo) <= 1e-12 * abs(fo)
        assert np.abs(g[b] - go).max() <= 1e-12 * np.abs(go).max()


@pytest.mark.parametrize("cfg,n,mf,T,A,S,force", [
    ("cfg1", 500, 60, 5, 50, 3, True),       # one class tile, forced onto the streamed path
    ("cfg2", 4000, None, 8, 4, 1, True),
    ("cfg4", 1700, 200, 4, 6, 2, True),      # 17 leaves as three class tiles (the padded form, forced)
    ("cfg3", 2400, None, 8, 2, 1, False),    # k = 4..7: 10 921 columns, C = 12 (two class tiles, 22 chunks)
    ("cfg4", 3400, None, 8, 2, 1, True),     # k = 4..6: 2 729 columns, C = 17, padded form
])
def test_packed_chunk_trainer_matches_oracle(oracle, monkeypatch, cfg, n, mf, T, A, S, force):
    if force:
        monkeypatch.setenv("SQW_ADMM_PACKED_MODE", "10")
    p = make_problem(cfg, n, max_features=mf)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    lo = run_gpu(p, T, A, S, packed=True)
    assert lo.stats.eval_mode == 10
    assert lo.log.admm_iters == tr.admm_iters
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL
    assert _rel(lo.getsmX(), smo) <= REL_TOL


# --------------------------------------------------------------------------------------------
# MODE 11: C = 8 n + 1 classes (2^k labels + one outgroup class: 16 + Random at cfg-4, 8 + Random
# here): n class tiles of DMMAs, the last class as DFMAs on the operands the DMMA path converted.
# --------------------------------------------------------------------------------------------
def _extra_class_problem(cfg, n, mf):
    p = make_problem(cfg, n, max_features=mf, outgroup=True)
    assert p.num_classes in (9, 17)
    return p


@pytest.mark.parametrize("cfg,n,mf,T", [("cfg4", 1700, 200, 4), ("cfg4", 3400, None, 8),
                                          ("cfg2", 3000, None, 8), ("cfg2", 900, 100, 5)])
def test_extra_class_evaluation_matches_oracle(oracle, cfg, n, mf, T):
    p = _extra_class_problem(cfg, n, mf)
    nrow, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    rng = np.random.default_rng(7)
    cx = rng.normal(size=(T, C * dim)) * 0.05
    smx = rng.normal(size=NN * dim) * 0.05
    f, g = opt.block_evaluate(None, p.cls, p.weights, p.crs, C, T, cx, smx, 1.7, counts=packed_counts(p))
    nb = nrow // T
    zd = np.zeros(NN * NN * dim)
    for b, t in enumerate(opt.java_block_order(T)):
        rows = np.arange(nb) * T + t
        fo, go = oracle.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], cx[b], smx, zd, zd, 1.7,
                                p.st, C)
        assert abs(f[b] - fo) <= 1e-12 * abs(fo)
        assert np.abs(g[b] - go).max() <= 1e-12 * np.abs(go).max()


@pytest.mark.parametrize("cfg,n,mf,T,A,S", [
    ("cfg4", 1700, 200, 4, 6, 2),      # 17 leaves incl. the parentless Random leaf: 2 class tiles + 1
    ("cfg4", 3400, None, 8, 2, 1),     # k = 4..6: 2 729 columns, 8 chunks
    ("cfg2", 3000, None, 8, 3, 1),     # 8 ring labels + Random = 9: 1 class tile + 1 (slice cache in use)
])
def test_extra_class_trainer_matches_oracle(oracle, cfg, n, mf, T, A, S):
    p = _extra_class_problem(cfg, n, mf)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    lo = run_gpu(p, T, A, S, packed=True)
    assert lo.stats.eval_mode == 11
    assert lo.log.admm_iters == tr.admm_iters
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL
    assert _rel(lo.getsmX(), smo) <= REL_TOL


def test_no_packed_hook_expands_counts_to_fp64(oracle, monkeypatch):
    """SQW_ADMM_NO_PACKED=1: the count form is expanded on the device with the reference's own
    (n - xMean) / xSD (Classifier.java:373-381) and trained by the fp64 kernels."""
    monkeypatch.setenv("SQW_ADMM_NO_PACKED", "1")
    p = make_problem("cfg4", 1700, max_features=200)
    T, A, S = 4, 6, 2
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes,
                                      num_threads=T, admm_max_itr=A, nodes_max_itr=S, host_threads=8)
    lo = run_gpu(p, T, A, S, packed=True)
    assert lo.stats.eval_mode in (4, 5, 6)
    assert np.array_equal(lo.log.nfev, tr.log_nfev[:, opt.java_block_order(T)])
    assert _rel(lo.getX(), xo) <= REL_TOL
