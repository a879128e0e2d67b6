"""BASELINE config 2 as the reference runs it (README typical: 20 000 sites, 8 labels, k = 4..5,
--r 10, 3-fold CV): the model on all sites + 3 fold models, one after the other on the whole GPU
(slots=1) and side by side with a CTA budget each (slots=3). Wall clock of classifier.cross_validate."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import classifier as clf
from helpers import make_problem
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p = make_problem('cfg2')
folds = clf.stratified_folds(p.cls, 3, seed=1)
opts = dict(ridge=10.0, pho=1.7, num_threads=8, admm_max_its=400, sequnwinder_max_its=15)
res = {}
for s in (1, slots):
    t0 = time.time()
    cv = clf.cross_validate(p.crs, p.counts, p.cls, p.weights, p.num_classes, folds, slots=s, **opts)
    res[s] = (time.time() - t0, cv)
    print('slots=%d: %.2f s for 1 + 3 trainings, CV accuracy %.4f, fold accuracies %s' % (
        s, res[s][0], cv.accuracy, ['%.4f' % a for a in cv.fold_accuracy]), flush=True)
a, b = res[1][1], res[slots][1]
print('speed-up %.2fx; same CV arg-max %.4f' % (res[1][0] / res[slots][0], (a.predictions == b.predictions).mean()))
