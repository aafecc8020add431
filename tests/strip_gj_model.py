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


def panel_gj_stream(strip, isfree, wk):
    """The panel as kb_chainfac.cu runs it with the column stream: the same eliminations, but
    every column also emits its ELEMENTARY transform (pivot row r, multipliers g with 1/pivot in
    slot r) -- what the kernel stores as tagged 32-byte elements -- instead of relying on the
    composite left in the strip.  The pivot row restarts from zero and is updated like every
    other row (0 + (1/pivot) * row), the form the kernel uses to keep the update branch-free."""
    stream = []
    for c in range(wk):
        mag = np.where(isfree, np.abs(strip[:, c]), -1.0)
        r = int(np.argmax(mag))
        isfree[r] = False
        prow = strip[r, :].copy()
        pinv = 1.0 / prow[c]
        mult = -strip[:, c] * pinv
        mult[r] = pinv
        elem = -mult
        elem[r] = pinv                      # the streamed element: g_i, or 1/pivot on the pivot row
        stream.append((r, elem.copy()))
        strip[r, :] = 0.0
        strip += np.outer(mult, prow)       # a[j] <- a[j] + mult * prow[j] for every row
        strip[:, c] = mult                  # -g, or 1/pivot on the pivot row
    return stream


def apply_stream(A, stream):
    """A consumer strip applies the streamed columns one by one (kb_chainfac.cu, consumer path):
    A[i,:] <- A[i,:] - g_i A[piv,:] (i != piv), A[piv,:] <- A[piv,:] / pivot."""
    for r, elem in stream:
        prow = A[r, :].copy()
        g = elem.copy()
        g[r] = 0.0
        A -= np.outer(g, prow)
        A[r, :] = prow * elem[r]
    return A


def strip_invert_stream(S, w):
    """strip_invert with the column stream in place of the composite transforms."""
    b = S.shape[0]
    K = (b + w - 1) // w
    strips = [S[:, k * w:min(b, (k + 1) * w)].astype(np.complex128).copy() for k in range(K)]
    isfree = np.ones(b, dtype=bool)
    rowpiv = np.zeros(b, dtype=int)
    for k in range(K):
        k0 = k * w
        wk = strips[k].shape[1]
        stream = panel_gj_stream(strips[k], isfree, wk)
        rowpiv[k0:k0 + wk] = [r for r, _ in stream]
        for s in range(K):
            if s != k:
                apply_stream(strips[s], stream)
    Y = np.concatenate(strips, axis=1)
    M = np.empty_like(Y)
    M[:, rowpiv] = Y[rowpiv, :]
    return M


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


def panel_gj_lazy(strip, isfree, scale, wk):
    """The panel with LAZY normalisation of the pivot rows (what kb_chainfac.cu runs): a row that
    pivots is not divided by its pivot -- it only remembers 1/pivot (`scale`) -- and is afterwards
    updated like every other row, with the multiplier formed from its un-normalised entry:

        true row = stored row * scale,      g~ = stored[i, c] / pivot' = scale^-1 g_true,
        stored[i, :] -= g~ prow'[:]   <=>   true[i, :] -= g_true prow'[:]

    so the scale cancels in every later update and is applied once, when the inverse is stored.
    At its own pivot step the row does not change at all (multiplier 0) except that its entry in
    the pivot column becomes 1 (true value 1/pivot).  This removes the special treatment of the
    pivot row from the update of every column (no zeroing, no scaling pass)."""
    stream = []
    for c in range(wk):
        mag = np.where(isfree, np.abs(strip[:, c]), -1.0)
        r = int(np.argmax(mag))
        isfree[r] = False
        prow = strip[r, :].copy()           # a free row: stored == true
        pinv = 1.0 / prow[c]
        scale[r] = pinv
        mult = -strip[:, c] * pinv
        mult[r] = 0.0
        elem = -mult
        elem[r] = pinv                      # streamed: g~_i, or 1/pivot on the pivot row
        stream.append((r, elem.copy()))
        strip += np.outer(mult, prow)
        strip[:, c] = mult
        strip[r, c] = 1.0
    return stream


def apply_stream_lazy(A, stream):
    """Consumer strips under lazy normalisation: A[i,:] -= g~_i A[piv,:] for i != piv; the pivot
    row itself is left alone (its scale is taken from the stream element)."""
    for r, elem in stream:
        prow = A[r, :].copy()
        g = elem.copy()
        g[r] = 0.0
        A -= np.outer(g, prow)
    return A


def strip_invert_lazy(S, w):
    b = S.shape[0]
    K = (b + w - 1) // w
    strips = [S[:, k * w:min(b, (k + 1) * w)].astype(np.complex128).copy() for k in range(K)]
    isfree = np.ones(b, dtype=bool)
    scale = np.ones(b, dtype=np.complex128)
    rowpiv = np.zeros(b, dtype=int)
    for k in range(K):
        k0 = k * w
        wk = strips[k].shape[1]
        stream = panel_gj_lazy(strips[k], isfree, scale, wk)
        rowpiv[k0:k0 + wk] = [r for r, _ in stream]
        for s in range(K):
            if s != k:
                apply_stream_lazy(strips[s], stream)
    Y = np.concatenate(strips, axis=1) * scale[:, None]   # the scales are applied when M is stored
    M = np.empty_like(Y)
    M[:, rowpiv] = Y[rowpiv, :]
    return M
