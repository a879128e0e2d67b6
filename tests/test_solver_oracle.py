"""CPU: the restated LBFGS/Mcsrch (oracle/lbfgs_oracle.c; framework/LBFGS.java:338-592,
framework/Mcsrch.java) driven exactly as the worker drives it (framework/LOne.java:865-896)."""
import numpy as np


def rosen(x):
    a, b = x[0::2], x[1::2]
    t1 = b - a * a
    g = np.zeros_like(x)
    g[0::2] = -400 * a * t1 - 2 * (1 - a)
    g[1::2] = 200 * t1
    return float(np.sum(100 * t1 * t1 + (1 - a) ** 2)), g


def test_rosenbrock_reaches_minimum(oracle):
    x, iflag, nfev, iters = oracle.lbfgs_minimize(rosen, np.tile([-1.2, 1.0], 5), eps=1e-8)
    assert iflag == 0
    assert np.allclose(x, 1.0, atol=1e-6)
    assert 20 < nfev < 200 and iters <= nfev


def test_loose_eps_stops_early_but_after_one_line_search(oracle):
    # LBFGS.java:567-572: the only convergence test sits after a completed line search, so even
    # a start point that already satisfies eps costs at least two evaluations.
    calls = []

    def quad(x):
        calls.append(1)
        return 0.5 * float(x @ x), x.copy()

    x, iflag, nfev, iters = oracle.lbfgs_minimize(quad, np.full(4, 1e-3), eps=0.1)
    assert iflag == 0 and nfev == 2 and iters == 1 and len(calls) == 2


def test_first_step_is_one_over_gnorm(oracle):
    # LBFGS.java:397,503: first trial step = min(1, 1/||g||)
    pts = []

    def quad(x):
        pts.append(x.copy())
        return 0.5 * float(x @ x), x.copy()

    x0 = np.array([30.0, 40.0])        # ||g|| = 50 -> first trial at x0 - g/50
    oracle.lbfgs_minimize(quad, x0, eps=1e-12)
    assert np.allclose(pts[1], x0 - x0 / 50.0, rtol=0, atol=1e-15)


def test_zero_gradient_is_a_line_search_failure(oracle):
    # gnorm == 0: dginit = -0 >= 0 -> "not a descent direction" -> iflag -1, x untouched
    x0 = np.zeros(3)
    x, iflag, nfev, _ = oracle.lbfgs_minimize(lambda x: (0.0, np.zeros_like(x)), x0)
    assert iflag == -1 and nfev == 1 and np.array_equal(x, x0)


def test_step_cap_1000(oracle):
    # Mcsrch.java:621-631: the step multiplier is clipped to 1000. On an almost linear function
    # the extrapolation stp + 4*(stp - stx) grows geometrically until the cap holds it: with
    # ||g|| = 1e-3 and d = -g the first line search can move at most 1000 * 1e-3 = 1.0.
    steps = []

    def f(x):
        steps.append(float(x[0]))
        return -1e-3 * x[0] + 1e-9 * x[0] ** 2, np.array([-1e-3 + 2e-9 * x[0], 0.0])

    oracle.lbfgs_minimize(f, np.zeros(2), eps=1e-30)
    first = steps[1:9]
    assert np.allclose(first[:5], [0.001, 0.005, 0.021, 0.085, 0.341])
    assert max(first) == 1.0 and first.count(1.0) >= 2
