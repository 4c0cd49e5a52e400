"""Host-side solver glue of the drop-in ALIGNF / NLCK classes (outside the GPU hot path, written independently of the
reference): checked here against a literal restatement of the reference's update rules (ALIGNF.py:60-89,
NLCKernels.py:68-92), with the Gram-side calls replaced by numpy so that no GPU is needed."""
import numpy as np
import pytest
from scipy.optimize import fmin_l_bfgs_b


def _spd(rng, p):
    A = rng.standard_normal((p, p))
    return A @ A.T + p * np.eye(p)


def test_alignf_weights_qp(capsys):
    import ALIGNF as A
    rng = np.random.default_rng(5)
    for p in (2, 3, 6):
        M, a = _spd(rng, p), rng.standard_normal(p) * 3 + 1
        obj = object.__new__(A.ALIGNF)
        obj.M, obj.a, obj.p, obj.Nfeval = M, a, p, 1
        np.random.seed(17)
        got = obj.get_v()
        # ALIGNF.py:81-89 restated: randn start, v >= 0, L-BFGS-B with pgtol 1e-6, unit Euclidean norm
        np.random.seed(17)
        v0 = np.random.randn(p)
        res = fmin_l_bfgs_b(lambda v: np.dot(v.T, np.dot(M, v)) - 2 * np.dot(v, a), v0, fprime=lambda v: 2 * np.dot(M, v) - 2 * a,
                            bounds=[[0, float(np.inf)]] * p, pgtol=1e-6)
        want = res[0] / np.linalg.norm(res[0])
        assert np.allclose(got, want, rtol=0, atol=1e-12), (p, got, want)
        assert np.all(got >= 0) and abs(np.linalg.norm(got) - 1) < 1e-12
        assert obj.loss(want) == pytest.approx(want @ M @ want - 2 * want @ a)
        assert np.allclose(obj.jac(want), 2 * M @ want - 2 * a)
    out = capsys.readouterr().out
    assert "Gradient descent..." in out and "Iteration  1 : loss=" in out and ", tol=" in out


def _reference_fit(p, eps, svm_step, grad, u0, fnorm, n_iter, eta):
    """NLCKernels.py:68-92, restated line by line."""
    def normalize(u):
        u_s = (u - u0)
        return u_s / np.sqrt(np.sum(u_s ** 2)) * fnorm + u0
    u = normalize(np.ones(p))
    u = np.array([0 if u[i] < 0 else u[i] for i in range(p)])
    score_prev = np.inf
    for _ in range(n_iter):
        alpha = svm_step(u)
        g = grad(u, alpha)
        u_next = normalize(u - eta * g)
        u_next = np.array([0 if u_next[i] < 0 else u_next[i] for i in range(p)])
        score = np.linalg.norm(u_next - u, np.inf)
        if score > score_prev:
            eta *= 0.8
        if score < eps:
            return u_next
        u = u_next
        score_prev = score
    return u_next


@pytest.mark.parametrize("u0,fnorm,eps", [(0, 10, 1e-8), (0.5, 1, 1e-3), (0, 1, 0.3)])
def test_nlck_projected_gradient_loop(capsys, u0, fnorm, eps):
    import NLCKernels as N
    rng = np.random.default_rng(3)
    p, n, degree = 4, 30, 2
    Ks = [_spd(rng, n) / n for _ in range(p)]
    y = np.where(rng.random(n) < 0.5, -1.0, 1.0)

    def svm_step(u):  # a stand-in for the QP: deterministic, depends on u
        K = sum(w * k for w, k in zip(u, Ks)) ** degree
        return np.linalg.solve(K + np.eye(n), y)

    def grad(u, alpha):  # NLCKernels.py:61-66 in numpy
        Kt = sum(w * k for w, k in zip(u, Ks)) ** (degree - 1)
        return -degree * np.array([alpha @ (Kt * k) @ alpha for k in Ks])

    obj = object.__new__(N.NLCK)
    obj.p, obj.eps, obj.degree = p, eps, degree
    obj.svm_step, obj.grad = svm_step, grad
    got = obj.fit(u0=u0, fnorm=fnorm, n_iter=12, eta=1)
    want = _reference_fit(p, eps, svm_step, grad, u0, fnorm, 12, 1)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12), (got, want)
    u = rng.standard_normal(p)
    assert np.allclose(obj.normalize(u, u0, fnorm), (u - u0) / np.sqrt(np.sum((u - u0) ** 2)) * fnorm + u0, rtol=1e-14, atol=1e-14)
    assert "Iteration 0, u=" in capsys.readouterr().out


def test_nlck_svm_step_builds_the_reference_qp(monkeypatch):
    """svm_step hands cvxopt the same QP as NLCKernels.py:50-59 (P = K_u, q = -y, G = [diag(y); -diag(y)], h = [C 1; 0]).
    cvxopt is not installed here: a stand-in module records what it is given."""
    import sys
    import types
    import NLCKernels as N
    seen = {}
    fake = types.ModuleType("cvxopt")
    fake.matrix = lambda a, tc='d': np.array(a, dtype=float)

    def spmatrix(vals, rows, cols, tc='d'):
        rows, cols = np.asarray(rows), np.asarray(cols)
        dense = np.zeros((rows.max() + 1, cols.max() + 1))
        dense[rows, cols] = np.asarray(vals, dtype=float)
        return dense
    fake.spmatrix = spmatrix
    fake.solvers = types.SimpleNamespace(options={}, qp=lambda P, q, G, h: seen.update(P=P, q=q, G=G, h=h) or {'x': np.arange(len(q), dtype=float)[:, None]})
    monkeypatch.setitem(sys.modules, "cvxopt", fake)
    rng = np.random.default_rng(9)
    n, p, degree, C = 7, 3, 3, 0.25
    Ks = [_spd(rng, n) for _ in range(p)]
    y = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    u = rng.random(p)
    import pandas as pd
    obj = object.__new__(N.NLCK)
    obj.n, obj.C, obj.degree, obj.y = n, C, degree, pd.Series(y)
    obj._resident = lambda: (Ks, None, None)
    monkeypatch.setattr(N._res, "combine", lambda grams, w, degree, out=None: types.SimpleNamespace(
        to_host=lambda: sum(wi * k for wi, k in zip(w, grams)) ** degree))
    alpha = obj.svm_step(u)
    r, o, z = np.arange(n), np.ones(n), np.zeros(n)
    G = np.zeros((2 * n, n))
    G[np.r_[r, r + n], np.r_[r, r]] = np.r_[y, -y]
    assert np.array_equal(seen["P"], np.sum(np.array(Ks) * u[:, None, None], axis=0) ** degree)
    assert np.array_equal(seen["q"], -y) and np.array_equal(seen["G"], G) and np.array_equal(seen["h"], np.r_[o * C, z])
    assert fake.solvers.options == {'show_progress': False} and np.array_equal(alpha, np.arange(n, dtype=float))
