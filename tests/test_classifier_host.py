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
