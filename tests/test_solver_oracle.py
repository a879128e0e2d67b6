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


# --------------------------------------------------------------------------------------------
# Independent pin of the solver's control logic. The reference's LBFGS.java / Mcsrch.java are the
# RISO port of Nocedal's L-BFGS with the More-Thuente line search (MCSRCH). SciPy ships MINPACK-2's
# later statement of the same line search (DCSRCH, scipy/optimize/_dcsrch.py) - written by other
# people from the paper, not from the reference - and the two-loop recursion is textbook (Nocedal &
# Wright, alg. 7.4/7.5). A textbook L-BFGS around SciPy's DCSRCH, given the reference's constants
# (m = 5, ftol 1e-4, gtol 0.9, xtol 1e-16, stpmin 1e-5, stpmax 1e20, first step min(1, 1/||g||),
# stop at ||g|| / max(1, ||x||) <= eps after a completed line search), must therefore visit the
# points the restatement visits: same evaluation count, same iteration count, every trial point
# equal up to the rounding of a different operation order.
# --------------------------------------------------------------------------------------------
def _textbook_lbfgs(fg, x0, eps, m=5, ftol=1e-4, gtol=0.9, xtol=1e-16, stpmin=1e-5, stpmax=1e20):
    from scipy.optimize._dcsrch import DCSRCH

    pts = []

    def F(x):
        pts.append(x.copy())
        return fg(x)

    x = np.array(x0, dtype=np.float64)
    f, g = F(x)
    S, Y, it = [], [], 0
    while it < 1000:
        it += 1
        q, al = g.copy(), []
        for s, y in zip(reversed(S), reversed(Y)):
            rho = 1.0 / (y @ s)
            a = rho * (s @ q)
            q = q - a * y
            al.append((a, rho, s, y))
        r = q * ((S[-1] @ Y[-1]) / (Y[-1] @ Y[-1])) if S else q.copy()
        for a, rho, s, y in reversed(al):
            r = r + s * (a - rho * (y @ r))
        d = -r
        stp0 = min(1.0, 1.0 / np.sqrt(g @ g)) if it == 1 else 1.0
        cache = {}

        def ev(a):
            if a not in cache:
                cache[a] = F(x + a * d)
            return cache[a]

        ls = DCSRCH(lambda a: ev(a)[0], lambda a: ev(a)[1] @ d, ftol, gtol, xtol, stpmin, stpmax)
        stp, _, _, task = ls(stp0, phi0=f, derphi0=g @ d, maxiter=200)
        assert stp is not None, task
        xn = x + stp * d
        fn, gn = ev(stp)
        S.append(xn - x)
        Y.append(gn - g)
        if len(S) > m:
            S.pop(0)
            Y.pop(0)
        x, f, g = xn, fn, gn
        if np.sqrt(g @ g) / max(1.0, np.sqrt(x @ x)) <= eps:
            break
    return x, pts, it


def _mlr_problem(seed=0, N=300, D=12, C=4, lam=0.5):
    rng = np.random.default_rng(seed)
    X, yv = rng.normal(size=(N, D)), rng.integers(0, C, N)

    def fg(w):   # ridge-regularised multinomial logistic regression: the kind of objective path 2 solves
        W = w.reshape(C, D)
        Z = X @ W.T
        Z = Z - Z.max(1, keepdims=True)
        P = np.exp(Z)
        P /= P.sum(1, keepdims=True)
        R = P.copy()
        R[np.arange(N), yv] -= 1
        return float(-np.log(P[np.arange(N), yv]).sum() + 0.5 * lam * (w @ w)), (R.T @ X).ravel() + lam * w

    return fg, C * D, rng


def test_control_logic_matches_textbook_lbfgs_around_scipy_dcsrch(oracle):
    pytest = __import__("pytest")
    pytest.importorskip("scipy.optimize._dcsrch")
    mlr, n_mlr, rng = _mlr_problem()
    Aq = np.random.default_rng(1).normal(size=(20, 20))
    H, bq = Aq @ Aq.T + 2 * np.eye(20), np.random.default_rng(2).normal(size=20)
    cases = [
        ("rosenbrock", rosen, np.tile([-1.2, 1.0], 5), 1e-8),        # 49 evaluations, 38 iterations:
                                                                     # real interpolation steps
        ("mlr eps 0.1", mlr, np.zeros(n_mlr), 0.1),                  # the x-update's own tolerance
        ("mlr eps 1e-6", mlr, np.zeros(n_mlr), 1e-6),
        ("mlr far start", mlr, rng.normal(size=n_mlr) * 3, 1e-6),
        ("quadratic", lambda x: (0.5 * float(x @ H @ x) - float(bq @ x), H @ x - bq), np.zeros(20), 1e-6),
    ]
    for name, fg, x0, eps in cases:
        seen = []

        def fg_rec(x, fg=fg):
            seen.append(x.copy())
            return fg(x)

        xo, iflag, nfev, iters = oracle.lbfgs_minimize(fg_rec, x0, eps=eps)
        xt, pts, it = _textbook_lbfgs(fg, x0, eps)
        assert iflag == 0, name
        assert (nfev, iters) == (len(pts), it), (name, nfev, iters, len(pts), it)
        scale = max(1.0, float(np.abs(xo).max()))
        assert max(float(np.abs(a - b).max()) for a, b in zip(seen, pts)) <= 1e-8 * scale, name
        assert float(np.abs(xo - xt).max()) <= 1e-8 * scale, name
