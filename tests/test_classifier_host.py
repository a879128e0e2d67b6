"""CPU: the host-side mirror of Classifier.buildClassifier's arithmetic
(sequnwinder_b200/classifier.py; framework/Classifier.java:300-420,452-477,267-287) against the
oracle's sequential restatement."""
import numpy as np

from helpers import make_problem
from sequnwinder_b200 import classifier as clf


def test_prepare_matches_oracle(oracle):
    p = make_problem("cfg1", 500, max_features=None)
    prep = clf.prepare(p.counts, p.cls, p.weights, p.crs, p.num_classes)
    assert np.array_equal(prep.keep, p.keep)                 # RemoveUseless: constant columns dropped
    assert prep.data.shape == p.data.shape
    assert np.allclose(prep.mean, p.mean, rtol=1e-13, atol=0)
    assert np.allclose(prep.sd, p.sd, rtol=1e-12, atol=0)
    assert np.allclose(prep.data, p.data, rtol=1e-11, atol=1e-12)
    assert np.array_equal(prep.x0, p.x0) and np.array_equal(prep.sm_x0, p.smx0)


def test_destandardize_and_predict_match_oracle(oracle):
    p = make_problem("cfg1", 300, max_features=40)
    C, dim = p.num_classes, p.data.shape[1]
    rng = np.random.default_rng(2)
    x = rng.normal(size=C * dim) * 0.1
    par = clf.destandardize(x, C, dim, p.mean, p.sd)
    assert np.allclose(par, oracle.destandardize(x, C, dim, p.mean, p.sd), rtol=1e-12, atol=1e-14)
    pr = clf.evaluate_probability(par, p.raw)
    assert np.allclose(pr, oracle.predict(par, p.raw), rtol=1e-10, atol=1e-14)
    assert np.array_equal(pr.argmax(1), oracle.predict(par, p.raw).argmax(1))


def test_weight_rule():
    p = make_problem("cfg1", 300, max_features=10)
    w = np.full(len(p.cls), 1e-3)
    try:
        clf.prepare(p.counts, p.cls, w, p.crs, p.num_classes)
        raise AssertionError("expected the 'Sum of weights' error (Classifier.java:359-362)")
    except ValueError as e:
        assert "Sum of weights" in str(e)


def test_stratified_folds_partition_the_sites():
    cls = np.array([0, 1, 2] * 11 + [0] * 4)
    folds = clf.stratified_folds(cls, 3, seed=1)
    assert sorted(np.concatenate(folds).tolist()) == list(range(cls.size))   # a partition
    for c in range(3):                                                       # stratified
        per = [int((cls[f] == c).sum()) for f in folds]
        assert max(per) - min(per) <= 1
    again = clf.stratified_folds(cls, 3, seed=1)
    assert all(np.array_equal(a, b) for a, b in zip(folds, again))           # deterministic


class _StubOptimizer:
    """Stands in for LOneCuda in host-logic tests (no GPU here): records the calls and 'trains'
    by writing a value derived from lambda and the CTA budget into x."""
    made = []

    def __init__(self, x, sm_x, data):
        self.x, self.sm_x, self.data = x, sm_x, data
        self.regularization, self._cta_budget, self._handle = 0.0, 0, None
        self.stats, self.log = None, None
        self.rows = None if data is None else data.shape[0]
        _StubOptimizer.made.append(self)

    def setCountMatrix(self, counts, mean, sd):   # the packed-count form Classifier hands over
        assert counts.dtype == np.uint8 and mean.size == sd.size == counts.shape[1] + 1
        self.rows = counts.shape[0]

    def __getattr__(self, name):          # every other setter is accepted and ignored
        if name.startswith("set") or name in ("initUandZ", "clearOptimizer", "close"):
            return lambda *a, **k: None
        raise AttributeError(name)

    def setRidge(self, r): self.regularization = float(r)
    def setCtaBudget(self, b): self._cta_budget = int(b)

    def execute(self):
        if self.regularization < 0:
            raise RuntimeError("negative lambda")
        self.x[:] = self.regularization
        self.sm_x[:] = self._cta_budget

    def getX(self): return self.x
    def getsmX(self): return self.sm_x


def test_cross_validate_host_logic(monkeypatch):
    """Task queue, CTA budgets, fold bookkeeping and pooled accuracy of cross_validate with a stub
    trainer (the GPU parity of the same loop is tests/test_admm_gpu.py)."""
    p = make_problem("cfg1", 300, max_features=20)
    counts = p.counts[:, p.keep]
    monkeypatch.setattr(clf, "LOneCuda", _StubOptimizer)
    _StubOptimizer.made.clear()
    folds = clf.stratified_folds(p.cls, 3)
    cv = clf.cross_validate(p.crs, counts, p.cls, p.weights, p.num_classes, folds, slots=2, ridge=7.0)
    assert len(_StubOptimizer.made) == 4                          # all sites + one per fold
    assert {o._cta_budget for o in _StubOptimizer.made} == {148 // 2}
    assert sorted(o.rows for o in _StubOptimizer.made) == sorted(
        [300] + [300 - len(f) for f in folds])                    # trained on the complement
    assert (cv.predictions >= 0).all() and len(cv.fold_accuracy) == 3
    hits = sum(a * len(f) for a, f in zip(cv.fold_accuracy, folds))
    assert abs(cv.accuracy - hits / 300) < 1e-12                  # pooled over the folds
    with np.testing.assert_raises(RuntimeError):                  # a worker's error reaches the caller
        clf.cross_validate(p.crs, counts, p.cls, p.weights, p.num_classes, folds, slots=2, ridge=-1.0)
