"""
gen_golden.py -- run the UNMODIFIED reference (`/root/reference/kernels.py`, `ALIGNF.py`,
`NLCKernels.py`) in the build container and record golden input/output vectors under
`tests/golden/`.  Test infrastructure only; needs `/root/reference` (absent on the GPU box, which
uses the committed fixtures).

    python oracle/gen_golden.py            # ~2 min
    python oracle/gen_golden.py --big      # additionally re-derives the SHA-256 KATs of SURVEY App. B
                                           # (spectrum k=6 on 3000 rows takes ~6 min in the reference)

The reference has no tests and no golden vectors of its own, so these outputs
are the only pin there is.
"""
import argparse
import contextlib
import hashlib
import io
import os
import sys
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def load_reference():
    sys.path.insert(0, REF)
    # cvxopt is not installed here; ALIGNF.py/NLCKernels.py/SVM.py only need it at import time for
    # the parts we exercise.
    if "cvxopt" not in sys.modules:
        stub = types.ModuleType("cvxopt")
        stub.__path__ = []
        stub.matrix = stub.spmatrix = lambda *a, **k: None
        solv = types.ModuleType("cvxopt.solvers")
        solv.options = {}
        solv.qp = None
        stub.solvers = solv
        sys.modules["cvxopt"] = stub
        sys.modules["cvxopt.solvers"] = solv
    import tqdm as _tqdm
    _orig = _tqdm.tqdm

    def quiet(*a, **k):
        k["disable"] = True
        return _orig(*a, **k)
    _tqdm.tqdm = quiet
    import kernels as ref
    ref.tqdm = quiet
    return ref


def sha(K):
    return hashlib.sha256(np.ascontiguousarray(K, np.float64).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    sys.path.insert(0, HERE)
    import oracle_np as onp

    frames = [pd.read_csv(f"{REF}/Data/X{s}{i}.csv") for s in ("tr", "te") for i in range(3)]
    allX = pd.concat(frames, axis=0)  # Xtr0,Xtr1,Xtr2,Xte0,Xte1,Xte2 in file order
    codes = onp.encode(allX["seq"])
    assert codes.shape == (9000, 101)
    labels = np.concatenate([pd.read_csv(f"{REF}/Data/Ytr{i}.csv")["Bound"].to_numpy() for i in range(3)])
    np.savez_compressed(os.path.join(OUT, "dna9000.npz"), codes=codes, labels=labels.astype(np.int8))

    X0 = frames[0]
    g = {}
    kat = {}

    def rec(name, K):
        g[name] = np.ascontiguousarray(K)
        print(f"  {name:28s} shape={np.shape(K)} sha={sha(K)[:16]}")

    with contextlib.redirect_stdout(io.StringIO()):
        pass
    # ---- spectrum (kernels.py:28-47)
    for k, n in ((1, 64), (2, 64), (3, 128), (4, 64), (5, 48), (6, 40), (7, 6)):
        rec(f"sp_k{k}_n{n}", ref.get_spectrum_K(X0.iloc[:n], k))
    # ---- weighted degree (kernels.py:84-101)
    for d, n in ((1, 24), (4, 40), (5, 64), (10, 128), (12, 32)):
        rec(f"wd_d{d}_n{n}", ref.get_WD_K(X0.iloc[:n], d))
    g["wd_d4_pair00"] = np.array(ref.get_WD_d(X0.seq[0], X0.seq[0], 4, 101))
    # ---- mismatch (kernels.py:196-217) -- always normalised by the reference
    for (k, m), n in (((3, 0), 16), ((3, 1), 32), ((4, 1), 64), ((4, 2), 24), ((5, 1), 16), ((5, 2), 8), ((6, 1), 5)):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            K = ref.get_mismatch_K(X0.iloc[:n], k, m)
        rec(f"mm_k{k}_m{m}_n{n}", K)
    # ---- normalise / centre (kernels.py:387-415)
    Ksp = ref.get_spectrum_K(X0.iloc[:96], 3)
    rec("norm_in_sp3_n96", Ksp.copy())
    with contextlib.redirect_stdout(io.StringIO()):
        rec("norm_out_sp3_n96", ref.normalize_K(Ksp.copy()))
        Kone = Ksp.copy(); Kone[0, 0] = 1.0
        rec("norm_out_early_n96", ref.normalize_K(Kone))  # early-out: returned unchanged
    rec("center_out_sp3_n96", ref.center_K(Ksp))
    Kwd = ref.get_WD_K(X0.iloc[:96], 5)
    rec("center_in_wd5_n96", Kwd)
    rec("center_out_wd5_n96", ref.center_K(Kwd))
    # ---- local alignment (kernels.py:226-302): degenerate, identically zero
    x, y = X0.seq[0], X0.seq[1]
    g["la_affine_pair01"] = np.array(float(ref.affine_align(x, y, 11, 1, 0.5)))
    g["la_affine_pair01_neg"] = np.array(float(ref.affine_align(x, y, -11, -1, 0.5)))
    g["la_smith_pair01"] = np.array(float(ref.Smith_Waterman(x, y, 11, 1, 0.5)))
    rec("la_eig0_n4", ref.get_LA_K(X0.iloc[:4], 11, 1, 0.5, 0, 0))
    rec("la_smith_eig0_n3", ref.get_LA_K(X0.iloc[:3], 11, 1, 0.5, 1, 0))
    # ---- select_method DSL (kernels.py:461-505)
    with contextlib.redirect_stdout(io.StringIO()):
        rec("sel_SP_k3_n16", ref.select_method(X0.iloc[:16], "SP_k3"))
        rec("sel_WD_d5_n16", ref.select_method(X0.iloc[:16], "WD_d5"))
        rec("sel_MM_k3_m1_n16", ref.select_method(X0.iloc[:16], "MM_k3_m1"))
    # ---- WD with shifts (kernels.py:106-155) -- "next" row of SURVEY section 8(f)
    rec("wds_d3_s2_n12", ref.get_WDShifts_K(X0.iloc[:12], 3, 2))
    rec("wds_d5_s1_n10", ref.get_WDShifts_K(X0.iloc[:10], 5, 1))

    # ---- ALIGNF / NLCK Gram-side algebra on real kernels
    n_all, n_fit = 96, 64
    Xa = X0.iloc[:n_all].copy()
    ya = pd.DataFrame({"Id": np.arange(n_all), "Bound": labels[:n_all].astype(float)})
    rng = np.random.Generator(np.random.PCG64(7))
    fit_rows = np.sort(rng.choice(n_all, n_fit, replace=False))
    ID = np.arange(n_all)
    with contextlib.redirect_stdout(io.StringIO()):
        Ks = [ref.get_spectrum_K(Xa, 3), ref.get_WD_K(Xa, 5), ref.get_mismatch_K(Xa, 3, 1)]
        import ALIGNF as refA
        np.random.seed(11)
        A = refA.ALIGNF(Xa.iloc[fit_rows], ya.iloc[fit_rows], ID, [k.copy() for k in Ks])
        Km = A.get_K()
    g["alignf_fit_rows"] = fit_rows
    g["alignf_y"] = ya["Bound"].to_numpy()[fit_rows]
    for i, k in enumerate(Ks):
        g[f"alignf_K{i}"] = k
    g["alignf_a"] = np.asarray(A.a)
    g["alignf_M"] = np.asarray(A.M)
    g["alignf_u"] = np.asarray(A.u_star)
    g["alignf_Km"] = np.asarray(Km)
    print("  alignf a =", A.a, "u* =", A.u_star)
    import NLCKernels as refN
    u_fix = np.array([0.5, 0.3, 0.8])
    alpha = rng.standard_normal(n_fit)
    for deg in (1, 2, 3):
        with contextlib.redirect_stdout(io.StringIO()):
            N = refN.NLCK(Xa.iloc[fit_rows], ya.iloc[fit_rows], ID, [k.copy() for k in Ks], degree=deg)
            g[f"nlck_grad_deg{deg}"] = N.grad(u_fix, alpha)
            N.fit = lambda *a, **k: u_fix  # skip the cvxopt QP; get_K's own lines 96-99 then run
            g[f"nlck_Km_deg{deg}"] = N.get_K()
    g["nlck_u"] = u_fix
    g["nlck_alpha"] = alpha
    with contextlib.redirect_stdout(io.StringIO()):
        g["nlck_K0_normalized"] = ref.normalize_K(Ks[0].copy())

    # ---- SHA-256 known answers
    with contextlib.redirect_stdout(io.StringIO()):
        kat["sp_k3_Xtr0_256"] = sha(ref.get_spectrum_K(X0.iloc[:256], 3))
        kat["wd_d5_Xtr0_256"] = sha(ref.get_WD_K(X0.iloc[:256], 5))
        kat["mm_k4_m1_Xtr0_64"] = sha(ref.get_mismatch_K(X0.iloc[:64], 4, 1))
        if args.big:
            kat["wd_d10_Xtr0_256"] = sha(ref.get_WD_K(X0.iloc[:256], 10))
            Xc1 = pd.concat((frames[0], frames[3]), axis=0)
            kat["sp_k6_Xtr0_Xte0_3000"] = sha(ref.get_spectrum_K(Xc1, 6))
    # values recorded by the survey run of the same reference calls
    kat.setdefault("wd_d10_Xtr0_256", "f56313a7106d54f09560da13f6d803863851b44abfe74936909b49368bf0daa0")
    kat.setdefault("sp_k6_Xtr0_Xte0_3000", "032431a8831f847159d7df4f72c19d1ba3a44e5e68ef7ee26054ba9348dd1a3c")
    for k_, v in kat.items():
        print("  KAT", k_, v)
    g["kat_names"] = np.array(list(kat.keys()))
    g["kat_sha256"] = np.array(list(kat.values()))
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), **g)
    print("wrote", os.path.join(OUT, "ref_vectors.npz"))


if __name__ == "__main__":
    main()
