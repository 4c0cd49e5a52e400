"""
kmg/_dual.py -- host-side bookkeeping shared by the dual-form classifiers (KRR.py, KLR.py): which rows of the Gram a
DataFrame's Ids name, which coefficients survive the eps threshold, the intercept, the decision values and the accuracy.
A few gathers and matrix-vector products on nfit-vectors; the K_fit algebra itself is on the device (kmg/host.py
spd_solve, kmg/resident.py DeviceGram).  The attribute names are the reference's (KRR.py:22-66, KLR.py:59-111), because
the callers in utils.py read them.
"""
import numpy as np


def rows_of(ID, wanted):
    """Position in kernel order of every Id in `wanted` (Ids are unique; the reference looks each one up with np.where,
    KRR.py:25): one sort of ID and a binary search per wanted Id."""
    ID, wanted = np.asarray(ID), np.atleast_1d(np.asarray(wanted))
    order = np.argsort(ID, kind="stable")
    pos = np.searchsorted(ID[order], wanted)
    pos = np.minimum(pos, ID.size - 1)
    rows = order[pos]
    if not np.array_equal(ID[rows], wanted):
        raise KeyError("an Id of the data frame is not among the kernel's Ids")
    return rows


class DualClassifier:
    """f(x_i) = sum_j a_j K[sv_j, i] + b over the support rows `idx_sv` of a precomputed Gram `K` with Ids `ID`."""

    def _rows(self, X):
        return rows_of(self.ID, X.loc[:, 'Id'].to_numpy())

    def _start_fit(self, X, y):
        self.X_fit = X
        self.Id_fit = X.loc[:, 'Id'].to_numpy()
        self.idx_fit = rows_of(self.ID, self.Id_fit)
        self.y_fit = y.loc[:, 'Bound'].to_numpy()
        self.n = self.idx_fit.size

    def _finish_fit(self, coef):
        """Keep the coefficients above eps in magnitude; the intercept is the mean residual on those rows."""
        sv = np.flatnonzero(np.abs(coef) > self.eps)
        self.a, self.y_fit, self.idx_sv = coef[sv], self.y_fit[sv], self.idx_fit[sv]
        self.y_hat = self._decision(self.idx_sv)
        self.b = np.mean(self.y_fit - self.y_hat)

    def _decision(self, rows):
        rows = np.atleast_1d(rows)
        return self.a @ self.K[np.ix_(self.idx_sv, rows)]  # sum_j a_j K[sv_j, i] for every i in rows

    def predict(self, X):
        """Predicted labels (-1 / 1) of the rows named by X.Id."""
        self.Id_pred = X.loc[:, 'Id'].to_numpy()
        self.idx_pred = rows_of(self.ID, self.Id_pred)
        return np.sign(self._decision(self.idx_pred) + self.b)

    def score(self, pred, y):
        """Fraction of predictions equal to the labels (a DataFrame with 'Bound', or an array)."""
        truth = y if isinstance(y, np.ndarray) else y.loc[:, 'Bound'].to_numpy()
        if np.any(truth == 0):
            raise AssertionError("Labels must be -1 or 1, not 0 or 1")
        return float(np.mean(pred == truth))
