"""CPU: the restated trainer (oracle/lone_oracle.c; framework/LOne.java) — evaluation formulas
against finite differences and an independent vectorised NumPy statement, the documented quirks,
and the caller-side arithmetic of framework/Classifier.java."""
import numpy as np
import pytest

from helpers import make_problem


@pytest.fixture(scope="module")
def small():
    return make_problem("cfg1", 300, max_features=24)


def dense_index(st, NN, dim):
    """edge list [(leaf, parent or leaf)] in leaf order / parents order."""
    edges = []
    for e in range(st.num_nodes):
        if st.is_leaf[e]:
            ps = st.parent_idx[st.parent_ptr[e]:st.parent_ptr[e + 1]]
            leaf = int(st.node_index[e])
            if len(ps) == 0:
                edges.append((leaf, leaf, False))
            for p in ps:
                edges.append((leaf, int(p), True))
    return edges


def numpy_fg(X, y, w, W, smx, z, u, pho, edges, NN, dim):
    """Vectorised statement of LOne.java:936-1049 (different code path from the oracle's loops)."""
    C = W.shape[0]
    S = X @ W.T
    mx = S.max(1, keepdims=True)
    E = np.exp(S - mx)
    den = E.sum(1, keepdims=True)
    f = -np.sum(w * ((S[np.arange(len(y)), y] - mx[:, 0]) - np.log(den[:, 0])))
    P = E / den
    Y = np.zeros_like(P)
    Y[np.arange(len(y)), y] = 1
    G = ((P - Y) * w[:, None]).T @ X
    for leaf, par, has in edges:
        zo = (leaf * NN + par) * dim
        v = W[leaf, 1:] - z[zo + 1:zo + dim] + u[zo + 1:zo + dim]
        if has:
            v = v - smx[par * dim + 1:(par + 1) * dim]
        f += pho / 2 * np.sum(v * v)
        G[leaf, 1:] += pho * v
    return f, G.reshape(-1)


def test_objective_and_gradient(oracle, small):
    p = small
    n, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    rng = np.random.default_rng(1)
    W = rng.normal(size=(C, dim)) * 0.2
    smx = rng.normal(size=NN * dim) * 0.2
    z = rng.normal(size=NN * NN * dim) * 0.1
    u = rng.normal(size=NN * NN * dim) * 0.1
    f, g = oracle.eval_fg(p.data, p.cls, p.weights, W.reshape(-1), smx, z, u, 1.7, p.st, C)
    f2, g2 = numpy_fg(p.data, p.cls, p.weights, W, smx, z, u, 1.7, dense_index(p.st, NN, dim), NN, dim)
    assert abs(f - f2) <= 1e-12 * abs(f2)
    assert np.allclose(g, g2, rtol=1e-11, atol=1e-12)
    # central finite differences on a few coordinates (intercept and non-intercept)
    x = W.reshape(-1)
    for k in [0, 1, dim - 1, dim, 2 * dim + 3]:
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[k] += h
        xm[k] -= h
        fp, _ = oracle.eval_fg(p.data, p.cls, p.weights, xp, smx, z, u, 1.7, p.st, C)
        fm, _ = oracle.eval_fg(p.data, p.cls, p.weights, xm, smx, z, u, 1.7, p.st, C)
        assert abs((fp - fm) / (2 * h) - g[k]) <= 1e-5 * max(1.0, abs(g[k]))


def test_intercept_is_not_penalised(oracle, small):
    p = small
    n, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    W = np.zeros(C * dim)
    smx = np.zeros(NN * dim)
    z = np.zeros(NN * NN * dim)
    u = np.zeros(NN * NN * dim)
    f0, g0 = oracle.eval_fg(p.data, p.cls, p.weights, W, smx, z, u, 1.7, p.st, C)
    z2 = z.copy()
    for leaf in range(C):
        for par in range(NN):
            z2[(leaf * NN + par) * dim] = 123.0      # only intercept slots of z
    f1, g1 = oracle.eval_fg(p.data, p.cls, p.weights, W, smx, z2, u, 1.7, p.st, C)
    assert f0 == f1 and np.array_equal(g0, g1)      # LOne.java:988,994,1037,1043 start at w=1


def test_java_thread_order(oracle):
    # "Thread0".."Thread4": consecutive hash codes -> consecutive buckets of a 16-slot table
    def jhash(s):
        h = 0
        for ch in s:
            h = (31 * h + ord(ch)) & 0xFFFFFFFF
        return (h ^ (h >> 16)) & 15
    T = 5
    want = sorted(range(T), key=lambda t: (jhash(f"Thread{t}"), t))
    assert oracle.java_thread_order(T).tolist() == want
    for T in (1, 8, 10, 13, 30):
        o = oracle.java_thread_order(T).tolist()
        assert sorted(o) == list(range(T))


def test_dropped_remainder_rows(oracle, small):
    """Rows >= floor(N/T)*T never reach a block (LOne.java:337-357): changing them changes
    nothing."""
    p = small
    T = 7
    n = p.data.shape[0]
    assert n % T != 0
    kw = dict(num_threads=T, admm_max_itr=3, nodes_max_itr=1)
    x1, s1, _ = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, **kw)
    d2 = p.data.copy()
    d2[(n // T) * T:] = 99.0
    x2, s2, _ = oracle.lone_execute(p.x0, p.smx0, d2, p.cls, p.weights, p.st, p.num_classes, **kw)
    assert np.array_equal(x1, x2) and np.array_equal(s1, s2)


def test_host_threads_do_not_change_results(oracle, small):
    p = small
    kw = dict(num_threads=5, admm_max_itr=6, nodes_max_itr=2)
    a = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, host_threads=1, **kw)
    b = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, p.num_classes, host_threads=4, **kw)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2].log_nfev, b[2].log_nfev)


def test_first_admm_iteration_by_hand(oracle, small):
    """One ADMM iteration restated independently in NumPy on top of the restated L-BFGS:
    u-update with z = zold = u = 0, x-update, x-bar, z = soft(x-hat + u-bar), residual quirk."""
    p = small
    T, lam, pho, alpha = 3, 10.0, 1.7, 1.9
    n, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    edges = dense_index(p.st, NN, dim)
    xo, smo, tr = oracle.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, C, ridge=lam,
                                      pho=pho, num_threads=T, admm_max_itr=1, nodes_max_itr=1)
    nb = n // T
    zlen = NN * NN * dim
    z0 = np.zeros(zlen)
    xs, us, nfevs = [], [], []
    for t in range(T):
        rows = np.arange(nb) * T + t
        xt = p.x0.copy()
        ut = np.zeros(zlen)
        for leaf, par, has in edges:
            zo = (leaf * NN + par) * dim
            xhat = alpha * xt[leaf * dim:(leaf + 1) * dim] + (1 - alpha) * z0[zo:zo + dim]
            if has:
                xhat = xhat - alpha * p.smx0[par * dim:(par + 1) * dim]
            ut[zo:zo + dim] = (ut[zo:zo + dim] + xhat - z0[zo:zo + dim]) / 1.0
        fg = lambda x: oracle.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], x, p.smx0, z0, ut,
                                      pho, p.st, C)
        xt, iflag, nfev, _ = oracle.lbfgs_minimize(fg, xt, eps=0.1)
        xs.append(xt)
        us.append(ut)
        nfevs.append(nfev)
    assert tr.log_nfev[0].tolist() == nfevs
    order = oracle.java_thread_order(T)
    xbar = sum(xs[t] for t in order) / T
    assert np.allclose(xo, xbar, rtol=0, atol=1e-15)
    ubar = sum(us[t] for t in order) / T
    kappa = 2 * lam / (pho * T)
    prim = 0.0
    for leaf in range(C):
        r_t = 0.0
        for (l2, par, has) in [e for e in edges if e[0] == leaf]:
            zo = (leaf * NN + par) * dim
            xhat = alpha * xbar[leaf * dim:(leaf + 1) * dim]
            sp = p.smx0[par * dim:(par + 1) * dim] if has else 0.0
            zz = xhat - alpha * sp + ubar[zo:zo + dim] if has else xhat + ubar[zo:zo + dim]
            zz = zz - np.sign(zz) * np.minimum(kappa, np.abs(zz))
            for t in range(T):
                r_t += np.sum((xs[t][leaf * dim:(leaf + 1) * dim] - zz - sp) ** 2)
            prim += np.sqrt(r_t)            # accumulates across the leaf's parents (the quirk)
    assert abs(prim - tr.log_primal[0]) <= 1e-10 * prim


def test_standardize_init_destandardize_predict(oracle, small):
    p = small
    raw, w = p.raw, p.weights
    data, mean, sd = oracle.standardize(raw, w)
    W = w.sum()
    m = (raw * w[:, None]).sum(0) / W
    s = np.sqrt(np.abs((raw ** 2 * w[:, None]).sum(0) - W * m * m) / (W - 1))
    assert np.allclose(mean[1:], m[1:], rtol=1e-12) and np.allclose(sd[1:], s[1:], rtol=1e-10)
    assert mean[0] == 0 and sd[0] == 1 and np.all(data[:, 0] == 1)
    C = p.num_classes
    counts = np.bincount(p.cls, minlength=C)
    assert np.allclose(p.x0.reshape(C, -1)[:, 0], np.log(counts + 1.0) - np.log(counts[C - 1] + 1.0))
    assert np.all(p.x0.reshape(C, -1)[:, 1:] == 0)
    # prediction on raw data with de-standardised weights == on standardised data with x
    rng = np.random.default_rng(0)
    x = rng.normal(size=C * raw.shape[1]) * 0.1
    par = oracle.destandardize(x, C, raw.shape[1], mean, sd)
    pr = oracle.predict(par, raw)
    S = data @ x.reshape(C, -1).T
    P = np.exp(S - S.max(1, keepdims=True))
    P /= P.sum(1, keepdims=True)
    assert np.allclose(pr, P, rtol=1e-9, atol=1e-12)
