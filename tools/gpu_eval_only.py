import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sequnwinder_b200 import optimizer as opt
from helpers import make_problem
p = make_problem('cfg2', 20000)
T = 8
C, dim = p.num_classes, p.data.shape[1]
rng = np.random.default_rng(0)
cx = rng.normal(0, 0.05, (T, C * dim))
for _ in range(3):
    f, g = opt.block_evaluate(p.data, p.cls, p.weights, p.crs, C, T, cx, p.smx0, 1.7)
print('f', f[:3])
