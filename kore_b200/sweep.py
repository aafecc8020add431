"""Sweep drivers over independent shifts (SURVEY.md 8e throughput mode, 8f rank 2).

Kore's forced problems are one shifted solve per forcing frequency,
``A_forced(omega) = ||B|| (A - i omega B)``, which the
reference runs as an outer shell loop that re-assembles and re-factors for every
frequency (/root/reference/tools/subramp.sh:59-141).  Here the pencil is ingested
once per GPU and only the numeric factorisation is repeated; the frequencies are
dealt round-robin to the ranks (one process per GPU) with no data-path
collective.  Eigenvalue sweeps over several targets work the same way.
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib


def deal(items, rank, world):
    """Round-robin share of `items` for this rank (the only 'partitioning' the sweep needs)."""
    return list(items[rank::world])


def forced_sweep(A, B, rhs, omegas, perm, nodeptr, device=0, rank=0, world=1, scale=1.0):
    """Solve (A - i omega B) x = rhs / scale for this rank's share of `omegas`.

    Returns (my_omegas, X) with one solution per column, and the per-frequency
    (factor_ms, solve_ms) measured with CUDA events."""
    mine = deal(list(omegas), rank, world)
    n = A.shape[0]
    X = np.empty((n, len(mine)), dtype=np.complex128, order="F")
    times = []
    with _lib.Solver(device) as s:
        s.set_pencil(A, B)
        s.set_chain(perm, nodeptr)
        b = np.asarray(rhs, dtype=np.complex128).ravel() / scale
        for k, om in enumerate(mine):
            s.factor(1j * om)
            X[:, k] = s.solve(b)
            st = s.stats()
            times.append((st["factor_ms"], st["solve_ms"]))
    return np.asarray(mine), X, np.asarray(times)


def eigen_sweep(A, B, targets, nev, perm, nodeptr, which="TM", device=0, rank=0, world=1, **kw):
    """nev eigenpairs around each of this rank's targets (e.g. mode tracking over Ek or m)."""
    out = []
    with _lib.Solver(device) as s:
        s.set_pencil(A, B)
        s.set_chain(perm, nodeptr)
        for tau in deal(list(targets), rank, world):
            s.factor(tau)
            lam, X, info = s.eigs(nev, which=which, target=tau, **kw)
            out.append((tau, lam, X, info))
    return out
