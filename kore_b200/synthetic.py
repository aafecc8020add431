"""Seeded synthetic pencils with the structure of Kore's hydrodynamic matrices.

The throughput sizes of BASELINE.json (E = 1e-7 .. 1e-9) cannot be assembled on
the GPU box (the reference assembler is not there, and a 250 MB A.npz is not
shipped), so bench.py builds a synthetic pencil of the identical STRUCTURE
(SURVEY.md 8d / App. A): sections [u | v] of nb blocks of N1 radial
coefficients each (section-major ordering, as assemble.py:442-458), which in
l-major (chain) order alternate v_l, u_{l+1}, v_{l+2}, ...; every diagonal
block is banded (u: +-6, v: +-4) below `chop` dense boundary rows (u: 4, v: 2);
the l+-1 couplings are banded (u<-v: +-7, v<-u: +-3) and absent from the
boundary rows; B is real, block diagonal, banded inside the diagonal band and
has all-zero boundary rows (so B is singular, like Kore's).

Values are random (numpy default_rng(seed)); A = B (i Omega + Delta) + couplings,
with Omega spread over the inertial band (-2, 2) and damping Delta <= 0 scaled
so that the density of eigenvalues around the shift sigma = 1j does not depend
on the size -- the Krylov iteration then needs a comparable number of operator
applications at every size.  This is host-side input generation only.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as ss

U_BAND, V_BAND, UV_BAND, VU_BAND = 6, 4, 7, 3
U_CHOP, V_CHOP = 4, 2


def _banded_block(rng, b, w, chop, scale, dtype=np.complex128):
    """COO triplets of a b x b block: rows >= chop carry offsets -w..w."""
    rows = np.arange(chop, b)
    offs = np.arange(-w, w + 1)
    r = np.repeat(rows, offs.size)
    c = r + np.tile(offs, rows.size)
    ok = (c >= 0) & (c < b)
    r, c = r[ok], c[ok]
    if dtype == np.complex128:
        v = (rng.standard_normal(r.size) + 1j * rng.standard_normal(r.size)) * scale
    else:
        v = rng.standard_normal(r.size) * scale
    return r, c, v


def synthetic_pencil(P, b, seed=20260101):
    """Returns (A csr complex128, B csr float64, perm, nodeptr) for a chain of P
    nodes of b unknowns; A, B are in Kore's section-major ordering and
    (perm, nodeptr) is the l-major chain that kb_set_chain expects."""
    rng = np.random.default_rng(seed)
    n = P * b
    nodeptr = np.arange(0, n + 1, b, dtype=np.int64)
    dens = 5000.0                      # eigenvalues per unit area of the complex plane
    D = max(0.05, n / (4.0 * dens))    # damping range

    R, Cc, VA = [], [], []
    RB, CB, VB = [], [], []
    for p in range(P):
        is_u = (p % 2 == 1)
        w = U_BAND if is_u else V_BAND
        chop = U_CHOP if is_u else V_CHOP
        o = p * b
        # ---- B diagonal block: real, banded, dominant diagonal, zero boundary rows
        r, c, v = _banded_block(rng, b, w, chop, 0.08, np.float64)
        dmask = r == c
        v[dmask] = 1.0 + 0.2 * rng.random(dmask.sum())
        Bblk = ss.csr_matrix((v, (r, c)), shape=(b, b))
        # ---- A diagonal block: B (i omega + delta) + small banded viscous-like part
        omega = rng.uniform(-2.0, 2.0, b)
        delta = -D * rng.random(b)
        Ablk = (Bblk @ ss.diags(1j * omega + delta)).tocoo()
        r2, c2, v2 = _banded_block(rng, b, w, chop, 0.02)
        # boundary rows: dense, Chebyshev-like growth, well-conditioned leading block
        k = np.arange(b, dtype=np.float64)
        br = np.repeat(np.arange(chop), b)
        bc = np.tile(np.arange(b), chop)
        grow = np.tile(1.0 + (k / b) ** 2 * 3.0, chop)
        bv = (rng.standard_normal(br.size) + 1j * rng.standard_normal(br.size)) * 0.3 * grow
        lead = bc < chop
        bv[lead & (br == bc)] += 2.0
        R += [Ablk.row + o, r2 + o, br + o]
        Cc += [Ablk.col + o, c2 + o, bc + o]
        VA += [Ablk.data, v2, bv]
        Bc = Bblk.tocoo()
        RB.append(Bc.row + o)
        CB.append(Bc.col + o)
        VB.append(Bc.data)
        # ---- l +- 1 couplings (Coriolis-like), absent from boundary rows
        for q in (p - 1, p + 1):
            if q < 0 or q >= P:
                continue
            wc = UV_BAND if is_u else VU_BAND
            r3, c3, v3 = _banded_block(rng, b, wc, chop, 0.15)
            R.append(r3 + o)
            Cc.append(c3 + q * b)
            VA.append(v3)
    A = ss.csr_matrix((np.concatenate(VA), (np.concatenate(R), np.concatenate(Cc))), shape=(n, n))
    B = ss.csr_matrix((np.concatenate(VB), (np.concatenate(RB), np.concatenate(CB))), shape=(n, n))
    # Kore divides both by ||B||_F (assemble.py:582-584)
    bn = np.sqrt((B.data ** 2).sum())
    A = (A / bn).tocsr()
    B = (B / bn).tocsr()

    # ---- chain order -> section-major [u | v] (u = odd nodes, v = even nodes)
    u_nodes = np.arange(1, P, 2)
    v_nodes = np.arange(0, P, 2)
    order = np.concatenate([u_nodes, v_nodes])            # section-major list of chain nodes
    sec_of_chain = np.empty(P, dtype=np.int64)
    sec_of_chain[order] = np.arange(P)
    # perm[k] = original (section-major) index of chain position k
    perm = (sec_of_chain[:, None] * b + np.arange(b)[None, :]).reshape(-1).astype(np.int64)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    # A_orig[perm[i], perm[j]] = A_chain[i, j]
    A = A[inv][:, inv].tocsr()
    B = B[inv][:, inv].tocsr()
    A.sort_indices()
    B.sort_indices()
    return A, B, perm, nodeptr


def start_vector(n, seed=1):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)
