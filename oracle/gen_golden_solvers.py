"""
gen_golden_solvers.py -- golden vectors for the K_fit algebra of the reference's closed-form solvers,
produced by importing the UNMODIFIED /root/reference/KRR.py and KLR.py (pure numpy) in this container.
*** TEST INFRASTRUCTURE ***  Writes tests/golden/ref_solvers.npz.

Kernels: the NLCK-style combination of tests/golden/ref_vectors.npz (nlck_Km_deg2: 96 x 96, unit diagonal) and the raw
weighted-degree Gram alignf_K1; fit rows / labels as for the ALIGNF goldens (labels mapped 0/1 -> -1/1).
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import KLR as refKLR  # noqa: E402
import KRR as refKRR  # noqa: E402


def main():
    g = np.load(os.path.join(HERE, "..", "tests", "golden", "ref_vectors.npz"))
    fit = g["alignf_fit_rows"]
    y = 2.0 * g["alignf_y"] - 1.0
    n_all = g["nlck_Km_deg2"].shape[0]
    ID = np.arange(n_all)
    X = pd.DataFrame({"Id": fit})
    Y = pd.DataFrame({"Id": fit, "Bound": y})
    Xall = pd.DataFrame({"Id": ID})
    out = {"fit_rows": fit, "y": y}
    for name, K in (("nlck2", g["nlck_Km_deg2"]), ("wd5", g["alignf_K1"])):
        for lbda in (0.1, 1e-3):
            m = refKRR.KRR(K.copy(), ID, lbda=lbda)
            m.fit(X, Y)
            out[f"krr_{name}_l{lbda}_a"] = m.a
            out[f"krr_{name}_l{lbda}_sv"] = m.idx_sv
            out[f"krr_{name}_l{lbda}_b"] = np.array(m.b)
            out[f"krr_{name}_l{lbda}_pred"] = m.predict(Xall)
        m = refKLR.KLR(K.copy(), ID, lbda=0.1)
        m.fit(X, Y)
        out[f"klr_{name}_a"] = m.a
        out[f"klr_{name}_sv"] = m.idx_sv
        out[f"klr_{name}_b"] = np.array(m.b)
        out[f"klr_{name}_pred"] = m.predict(Xall)
        W, z = m.IRLS(K[fit][:, fit], y, np.linspace(-0.01, 0.01, fit.size))
        m.n = fit.size
        out[f"klr_{name}_wkrr"] = m.WKRR(K[fit][:, fit], W, z)
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "ref_solvers.npz"), **out)
    print("wrote ref_solvers.npz:", sorted(out))


if __name__ == "__main__":
    main()
