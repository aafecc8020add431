"""The folded sweep algebra of kb_sweep2.cu (NumPy statement) solves the block-tridiagonal system."""
import numpy as np
import pytest

from fold_model import fold, folded_sweep, two_sided_factor


def random_chain(sizes, rng):
    P = len(sizes)
    D = [rng.standard_normal((b, b)) + 1j * rng.standard_normal((b, b)) + 4 * np.eye(b) for b in sizes]
    L = [None] + [0.3 * (rng.standard_normal((sizes[p], sizes[p - 1])) + 1j * rng.standard_normal((sizes[p], sizes[p - 1])))
                  for p in range(1, P)]
    U = [0.3 * (rng.standard_normal((sizes[p], sizes[p + 1])) + 1j * rng.standard_normal((sizes[p], sizes[p + 1])))
         for p in range(P - 1)] + [None]
    return D, L, U


def dense(D, L, U):
    off = np.concatenate([[0], np.cumsum([d.shape[0] for d in D])])
    T = np.zeros((off[-1], off[-1]), dtype=complex)
    for p in range(len(D)):
        T[off[p]:off[p + 1], off[p]:off[p + 1]] = D[p]
        if p > 0:
            T[off[p]:off[p + 1], off[p - 1]:off[p]] = L[p]
        if p + 1 < len(D):
            T[off[p]:off[p + 1], off[p + 1]:off[p + 2]] = U[p]
    return T, off


@pytest.mark.parametrize("sizes,mid", [
    ([5], 0),                                # a single node: x = M r
    ([4, 6], 1), ([4, 6, 3], 2),             # one-sided (mid = P - 1)
    ([5, 5, 5, 5], 2), ([7, 3, 7, 3, 7], 2),  # two-sided, uniform and alternating node sizes (thermal)
    ([6] * 9, 4), ([3, 8, 2, 9, 4, 7, 5], 3),
])
def test_folded_sweep_solves(sizes, mid):
    rng = np.random.default_rng(len(sizes) * 10 + mid)
    D, L, U = random_chain(sizes, rng)
    T, off = dense(D, L, U)
    rfull = rng.standard_normal(off[-1]) + 1j * rng.standard_normal(off[-1])
    r = [rfull[off[p]:off[p + 1]] for p in range(len(sizes))]
    M = two_sided_factor(D, L, U, mid)
    FL, FU = fold(M, L, U)
    x = np.concatenate(folded_sweep(M, FL, FU, r, mid))
    assert np.linalg.norm(T @ x - rfull) <= 1e-12 * np.linalg.norm(rfull) * np.linalg.cond(T)
