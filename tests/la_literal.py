"""la_literal.py -- test helper: a LITERAL log-space evaluation of the intended Vert-Saigo recursion, vectorised over
pairs with numpy.logaddexp.  Written from the reference's affine_align / Smith_Waterman (kernels.py:226-270) with five
DISTINCT state matrices, cells i = 1..n_x, j = 1..n_y and s = S[x[i-1], y[j-1]]; every sum of the
recursion is a chain of two-argument log-add-exps, every product an addition of logs.  It shares no code with the GPU
kernel (scaled linear space, csrc/la_kernel.cu) nor with the oracle (oracle/kmg_oracle.c, oracle_np.py): a third route to
the same numbers for the parity-unpinned kernel."""
import numpy as np

S = np.array([[4, 0, 0, 0], [0, 9, -3, -1], [0, -3, 6, 2], [0, -1, -2, 5]], dtype=np.float64)  # kernels.py:223, [x, y]


def _lae(*v):
    acc = v[0]
    for t in v[1:]:
        acc = np.logaddexp(acc, t)
    return acc


def la_pairs_logspace(xs, ys, e, d, beta, smith=0):
    """xs, ys: (P, L) uint8 codes 0..3; returns (P,) values of K_beta(x_p, y_p)."""
    xs, ys = np.asarray(xs), np.asarray(ys)
    P, L = xs.shape
    NI = np.full(P, -np.inf)
    zero = np.zeros(P)
    bd, be = beta * d, beta * e
    red = (lambda *v: np.maximum.reduce(v)) if smith else _lae
    # previous row (i-1) of every state, indexed by column j = 0..L
    pM = [NI.copy() for _ in range(L + 1)]; pX = [NI.copy() for _ in range(L + 1)]; pY = [NI.copy() for _ in range(L + 1)]
    pX2 = [NI.copy() for _ in range(L + 1)]; pY2 = [NI.copy() for _ in range(L + 1)]
    for i in range(1, L + 1):
        cM = [NI.copy()]; cX = [NI.copy()]; cY = [NI.copy()]; cX2 = [NI.copy()]; cY2 = [NI.copy()]
        xi = xs[:, i - 1]
        for j in range(1, L + 1):
            s = beta * S[xi, ys[:, j - 1]]
            cM.append(s + red(zero, pX[j - 1], pY[j - 1], pM[j - 1]))          # kernels.py:241 / :266
            cX.append(red(bd + pM[j], be + pX[j]))                              # kernels.py:242 / :267
            cY.append(red(bd + cM[j - 1], bd + cX[j - 1], be + cY[j - 1]))      # kernels.py:243 / :268
            cX2.append(red(pM[j], pX2[j]))                                      # kernels.py:244 / :269
            cY2.append(red(cM[j - 1], cX2[j - 1], cY2[j - 1]))                  # kernels.py:245 / :270
        pM, pX, pY, pX2, pY2 = cM, cX, cY, cX2, cY2
    return (1.0 / beta) * red(zero, pX2[L], pY2[L], pM[L])                      # kernels.py:246 / :271
