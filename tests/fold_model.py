"""NumPy statement of the folded chain sweep (kore_b200/csrc/kb_sweep2.cu).

Block-tridiagonal T with diagonal blocks D_p, sub-diagonal couplings L_p = T[p, p-1] and
super-diagonal couplings U_p = T[p, p+1]; the factorisation keeps the explicit inverses
M_p = S_p^-1 of the two-sided ("burn at both ends") Schur blocks.  The sweep never forms
y = M t or x explicitly along the chain: it runs on the right-hand sides of the dense products,

    forward    t_{p'}  = r_{p'}  - F_p t_p          F_p  = C_{p',p} M_p   (p' = node after p)
    middle     u_m     = r_m - FL_{m-1} t_{m-1} - FU_{m+1} t_{m+1}
    backward   u_{p''} = t_{p''} - F'_p u_p         F'_p = C_{p'',p} M_p  (p'' = node before p)
    solution   x_p     = M_p u_p                    (every node, no dependency)

with FL_p = L_{p+1} M_p and FU_p = U_{p-1} M_p formed once per factorisation.  This file is test
infrastructure (the checker of tests/test_fold_model.py); the product path is the CUDA kernel.
"""
import numpy as np


def two_sided_factor(D, L, U, mid):
    """M_p = S_p^-1 for the downward chain 0..mid-1, the upward chain P-1..mid+1 and the middle."""
    P = len(D)
    M = [None] * P
    for p in range(mid):
        S = D[p] if p == 0 else D[p] - L[p] @ M[p - 1] @ U[p - 1]
        M[p] = np.linalg.inv(S)
    for p in range(P - 1, mid, -1):
        S = D[p] if p == P - 1 else D[p] - U[p] @ M[p + 1] @ L[p + 1]
        M[p] = np.linalg.inv(S)
    S = D[mid].copy()
    if mid > 0:
        S = S - L[mid] @ M[mid - 1] @ U[mid - 1]
    if mid < P - 1:
        S = S - U[mid] @ M[mid + 1] @ L[mid + 1]
    M[mid] = np.linalg.inv(S)
    return M


def fold(M, L, U):
    """FL_p = L_{p+1} M_p (rows of node p+1), FU_p = U_{p-1} M_p (rows of node p-1)."""
    P = len(M)
    FL = [L[p + 1] @ M[p] if p + 1 < P else None for p in range(P)]
    FU = [U[p - 1] @ M[p] if p > 0 else None for p in range(P)]
    return FL, FU


def folded_sweep(M, FL, FU, r, mid):
    """x = T^-1 r by the schedules kbi_fold_prepare builds for the two groups."""
    P = len(M)
    t = [None] * P
    u = [None] * P
    # group 0: forward down to the middle
    t[0] = r[0]
    for p in range(mid):
        t[p + 1] = r[p + 1] - FL[p] @ t[p]          # for p + 1 == mid this is group 0's part of u_mid
    part0 = t[mid]
    # group 1: forward up to the middle
    part1 = 0
    if mid < P - 1:
        t[P - 1] = r[P - 1]
        for p in range(P - 1, mid + 1, -1):
            t[p - 1] = r[p - 1] - FU[p] @ t[p]
        part1 = -FU[mid + 1] @ t[mid + 1]             # published without a base
    u[mid] = part0 + part1
    # backward, both directions from the middle
    for p in range(mid, 0, -1):
        u[p - 1] = t[p - 1] - FU[p] @ u[p]
    for p in range(mid, P - 1):
        u[p + 1] = t[p + 1] - FL[p] @ u[p]
    return [M[p] @ u[p] for p in range(P)]
