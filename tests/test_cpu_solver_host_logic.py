"""CPU: the HOST side of the KRR.py / KLR.py drop-ins (kmg/_dual.py: Id lookup, support selection, intercept, predictions,
score, the IRLS loop) against the golden vectors of the unmodified reference (tests/golden/ref_solvers.npz), with the two
device calls -- kmg.host.spd_solve and kmg.resident.DeviceGram -- replaced by numpy stand-ins.  The device algebra itself
is tested on the GPU (tests/test_gpu_solvers.py); nothing here touches libkmg's compute entry points."""
import os

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _NumpyGram:
    """numpy stand-in for kmg.resident.DeviceGram (from_host / matvec / spd_solve)."""

    def __init__(self, K):
        self.K = np.array(K, dtype=np.float64)

    @classmethod
    def from_host(cls, K):
        return cls(K)

    def matvec(self, v):
        return self.K @ v

    def spd_solve(self, b, c, s=None):
        A = self.K if s is None else s[:, None] * self.K * s[None, :]
        return np.linalg.solve(A + c * np.eye(A.shape[0]), b)


def _numpy_spd_solve(K, b, c, idx=None, s=None):
    sub = K if idx is None else K[np.ix_(idx, idx)]
    return _NumpyGram(sub).spd_solve(np.asarray(b, np.float64), c, s=s)


@pytest.fixture()
def stand_ins(monkeypatch):
    from kmg import host, resident
    monkeypatch.setattr(host, "spd_solve", _numpy_spd_solve)
    monkeypatch.setattr(resident, "DeviceGram", _NumpyGram)


@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_solvers.npz"))


def _frames(gs, n_all):
    fit, y = gs["fit_rows"], gs["y"]
    return pd.DataFrame({"Id": fit}), pd.DataFrame({"Id": fit, "Bound": y}), pd.DataFrame({"Id": np.arange(n_all)})


def test_rows_of_follows_the_id_order():
    from kmg._dual import rows_of
    ID = np.array([40, 7, 19, 3, 88])
    assert np.array_equal(rows_of(ID, [3, 40, 88, 7]), [3, 0, 4, 1])
    assert np.array_equal(rows_of(ID, 19), [2])
    with pytest.raises(KeyError):
        rows_of(ID, [3, 5])
    with pytest.raises(KeyError):
        rows_of(ID, [100])


def test_krr_host_logic_matches_the_reference(gs, golden, stand_ins):
    import KRR as ours
    for name, K in (("nlck2", golden["nlck_Km_deg2"]), ("wd5", golden["alignf_K1"])):
        X, Y, Xall = _frames(gs, K.shape[0])
        for lbda in (0.1, 1e-3):
            m = ours.KRR(K.copy(), np.arange(K.shape[0]), lbda=lbda)
            m.fit(X, Y)
            want = gs[f"krr_{name}_l{lbda}_a"]
            assert np.array_equal(m.idx_sv, gs[f"krr_{name}_l{lbda}_sv"])
            assert np.abs(m.a - want).max() <= 1e-9 * np.abs(want).max(), (name, lbda)
            assert abs(m.b - float(gs[f"krr_{name}_l{lbda}_b"])) <= 1e-9 * max(1.0, abs(float(gs[f"krr_{name}_l{lbda}_b"])))
            assert np.array_equal(m.predict(Xall), gs[f"krr_{name}_l{lbda}_pred"])
            assert m.y_fit.shape == m.a.shape == m.y_hat.shape
            assert 0.5 <= m.score(m.predict(X), Y) <= 1.0
            assert m.score(gs["y"], gs["y"]) == 1.0
    with pytest.raises(AssertionError):
        m.score(np.ones(3), np.array([0.0, 1.0, 1.0]))


def test_klr_host_logic_matches_the_reference(gs, golden, stand_ins):
    import KLR as ours
    K = golden["nlck_Km_deg2"]
    X, Y, Xall = _frames(gs, K.shape[0])
    fit, y = gs["fit_rows"], gs["y"]
    m = ours.KLR(K.copy(), np.arange(K.shape[0]), lbda=0.1)
    m.n = fit.size
    sub = np.ascontiguousarray(K[fit][:, fit])
    W, z = m.IRLS(sub, y, np.linspace(-0.01, 0.01, fit.size))
    got, want = m.WKRR(sub, W, z), gs["klr_nlck2_wkrr"]
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    m.fit(X, Y)
    want = gs["klr_nlck2_a"]
    assert np.array_equal(m.idx_sv, gs["klr_nlck2_sv"])
    assert np.abs(m.a - want).max() <= 1e-7 * np.abs(want).max()
    assert abs(m.b - float(gs["klr_nlck2_b"])) <= 1e-7 * max(1.0, abs(float(gs["klr_nlck2_b"])))
    assert np.array_equal(m.predict(Xall), gs["klr_nlck2_pred"])


def test_klr_stops_updating_once_the_step_is_below_tol(golden, gs, stand_ins):
    """The reference leaves alpha alone once ||alpha - alpha_prev|| <= tol (KLR.py:69-77): a huge tol means ONE IRLS step."""
    import KLR as ours
    K = golden["nlck_Km_deg2"]
    X, Y, _ = _frames(gs, K.shape[0])
    fit, y = gs["fit_rows"], gs["y"]
    one = ours.KLR(K.copy(), np.arange(K.shape[0]), lbda=0.1, tol=1e9, eps=0.0)
    one.fit(X, Y)
    ref = ours.KLR(K.copy(), np.arange(K.shape[0]), lbda=0.1)
    ref.n = fit.size
    sub = np.ascontiguousarray(K[fit][:, fit])
    W, z = ref.IRLS(sub, y, np.zeros(fit.size))
    assert np.allclose(one.a, ref.WKRR(sub, W, z), rtol=1e-12, atol=0)


def test_resident_reformat_data_follows_the_reference():
    """kmg.resident.reformat_data against a literal restatement of utils.reformat_data (utils.py:296-312), with a numpy
    stand-in for the device gather (DeviceGram.gather itself: tests/test_gpu_solvers.py)."""
    from kmg import resident

    class G:
        def __init__(self, K):
            self.K = K

        def gather(self, idx):
            return G(self.K[idx][:, idx])

    rng = np.random.default_rng(3)
    n = 40
    ID = rng.permutation(np.arange(100, 100 + n))
    kernels = [rng.standard_normal((n, n)) for _ in range(2)]
    ids = rng.permutation(ID)[:24]

    def frames():
        Xtr, Xva, Xte = (pd.DataFrame({"Id": ids[a:b], "seq": ["A"] * (b - a)}) for a, b in ((0, 12), (12, 18), (18, 24)))
        ytr = pd.DataFrame({"Id": ids[0:12], "Bound": np.ones(12)})
        yva = pd.DataFrame({"Id": ids[12:18], "Bound": -np.ones(6)})
        return Xtr, ytr, Xva, yva, Xte

    # the reference, restated
    Xtr, ytr, Xva, yva, Xte = frames()
    ID_ = np.concatenate((np.array(Xtr.loc[:, 'Id']), np.array(Xva.loc[:, 'Id']), np.array(Xte.loc[:, 'Id'])))
    idx = np.array([np.where(ID == ID_[i])[0] for i in range(len(ID_))]).squeeze()
    want_k = [K[idx][:, idx] for K in kernels]

    got = resident.reformat_data(frames(), [G(K) for K in kernels], ID)
    gXtr, gytr, gXva, gyva, gXte, sub, new_ID = got
    assert np.array_equal(new_ID, np.arange(24))
    for g, w in zip(sub, want_k):
        assert np.array_equal(g.K, w)
    assert list(gXtr.Id) == list(range(0, 12)) and list(gXva.Id) == list(range(12, 18)) and list(gXte.Id) == list(range(18, 24))
    assert list(gytr.Id) == list(range(0, 12)) and list(gyva.Id) == list(range(12, 18))
    with pytest.raises(TypeError):
        resident.reformat_data(frames(), kernels, ID)
