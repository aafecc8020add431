"""NumPy statement of the column-strip Gauss-Jordan inversion used by the persistent
factor kernel (kore_b200/csrc/kb_chainfac.cu).

The b x b Schur block is split into K column strips of width w, one per CTA; row i of
every strip lives in thread i of the owning CTA and NEVER moves (implicit pivoting).
Strip k is "the panel" of step k: its owner eliminates its columns one by one, each time
choosing the largest entry among the rows that have not been a pivot yet; that leaves
the composite transform G_k (b x w_k) in the strip and the list of pivot rows.  Every
other strip then applies

    A[i, :] <- (0 if i is a pivot row of the step else A[i, :]) + sum_c G[i, c] A[piv_c, :]

After the last step the strips hold Y with Y[piv_c, :] = X[c, :], X = (Pi S)^-1 and Pi the
row permutation that brings row piv_c to position c, hence

    S^-1[c, piv_k] = Y[piv_c, k].

This file is test infrastructure (the checker of tests/test_strip_model.py); the product
path is the CUDA kernel.
"""
import numpy as np


def panel_gj(strip, isfree, wk):
    """In-place elimination of columns 0..wk-1 of strip (b x w).  Returns the pivot rows."""
    piv = []
    for c in range(wk):
        mag = np.where(isfree, np.abs(strip[:, c]), -1.0)
        r = int(np.argmax(mag))
        piv.append(r)
        isfree[r] = False
        prow = strip[r, :].copy()
        pinv = 1.0 / prow[c]
        g = strip[:, c] * pinv
        g[r] = 0.0
        new = strip - np.outer(g, prow)
        new[:, c] = -g
        new[r, :] = prow * pinv
        new[r, c] = pinv
        strip[:, :] = new
    return piv


def strip_invert(S, w):
    b = S.shape[0]
    K = (b + w - 1) // w
    strips = [S[:, k * w:min(b, (k + 1) * w)].astype(np.complex128).copy() for k in range(K)]
    isfree = np.ones(b, dtype=bool)
    rowpiv = np.zeros(b, dtype=int)  # rowpiv[k] = row that pivoted column k
    for k in range(K):
        k0 = k * w
        wk = strips[k].shape[1]
        piv = panel_gj(strips[k], isfree, wk)
        rowpiv[k0:k0 + wk] = piv
        G = strips[k][:, :wk]
        for s in range(K):
            if s == k:
                continue
            A = strips[s]
            R = A[piv, :].copy()
            base = A.copy()
            base[piv, :] = 0.0
            strips[s] = base + G @ R
    Y = np.concatenate(strips, axis=1)
    M = np.empty_like(Y)
    M[:, rowpiv] = Y[rowpiv, :]
    return M
