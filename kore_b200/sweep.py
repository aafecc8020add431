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


def forced_sweep(A, B, rhs, omegas, perm, nodeptr, device=0, rank=0, world=1, scale=1.0, handles=1):
    """Solve (A - i omega B) x = rhs / scale for this rank's share of `omegas`.

    `handles` > 1: that many library handles on this GPU, each on its own stream and driven by
    its own host thread (the C-ABI calls release the GIL), the rank's frequencies dealt among
    them.  A small pencil (a 1 600-unknown libration problem is ten CTAs of factorisation) leaves
    the GPU almost empty and its chain latency-bound; independent frequencies fill it.

    Returns (my_omegas, X) with one solution per column, and the per-frequency
    (factor_ms, solve_ms) measured with CUDA events."""
    mine = deal(list(omegas), rank, world)
    n = A.shape[0]
    X = np.empty((n, len(mine)), dtype=np.complex128, order="F")
    times = np.zeros((len(mine), 2))
    b = np.asarray(rhs, dtype=np.complex128).ravel() / scale
    handles = max(1, min(int(handles), len(mine)))

    def work(slot):
        with _lib.Solver(device) as s:
            s.set_pencil(A, B)
            s.set_chain(perm, nodeptr)
            for k in range(slot, len(mine), handles):
                s.factor(1j * mine[k])
                X[:, k] = s.solve(b)
                st = s.stats()
                times[k] = (st["factor_ms"], st["solve_ms"])

    if handles == 1:
        work(0)
    else:
        import threading
        errs = []

        def guarded(slot):
            try:
                work(slot)
            except BaseException as e:  # noqa: BLE001 -- re-raised in the caller's thread
                errs.append(e)

        threads = [threading.Thread(target=guarded, args=(i,)) for i in range(handles)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
    return np.asarray(mine), X, times


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


_generated = {}


def _operators_of(operators, q):
    """(radial operators of the parameter set q, a key that tells two sets of operators apart): the dict the
    caller holds, or operators generated from the parameters (one `radial.OperatorCache` per process when the
    caller passes None)."""
    if isinstance(operators, dict):
        return operators, None
    from . import radial as _radial
    cache = operators if operators is not None else _generated.setdefault("cache", _radial.OperatorCache())
    return cache.get(q), cache.key(q)


def parameter_sweep(pp, operators, cases, nev, which="TM", device=0, rank=0, world=1, want_vectors=False,
                    solver_factory=None, **kw):
    """Eigenpairs of a family of pencils that differ in their physical parameters -- azimuthal wavenumber m,
    symmetry, Rayleigh number (buoyancy factor), Ekman-number factors, boundary conditions -- on the same radial
    truncation: `cases` is a list of dicts of `assembly.PhysicsParams` fields to override (plus the optional key
    ``"tau"``: the shift / target of that case; default 0), dealt round-robin to the ranks.  Every pencil is
    ASSEMBLED ON THE GPU (kore_b200.assembly) from the radial operators (`operators`: a dict as
    `assembly.load_operators` returns, or None / a `radial.OperatorCache` to generate them from the parameters,
    once per truncation), which do not depend on m, symm or the dimensionless numbers, so a survey over m that the reference runs as one assemble.py + solve.py job per value
    (tools/subramp.sh-style loops; SURVEY.md 8e: "different azimuthal numbers m are different matrices") is here
    one handle per GPU and, per case, a new assembly program, the layout, one factorisation and one eigensolve.

    Returns [(case, eigenvalues, eigenvectors or None, info)] for this rank's cases."""
    from . import assembly as _asm
    from . import chain as _chain
    make = solver_factory or _lib.Solver
    out = []
    with make(device) as s:
        bnorm_of = {}
        for case in deal(list(cases), rank, world):
            fields = {k: v for k, v in case.items() if k != "tau"}
            q = _asm.PhysicsParams.from_dict({**pp.__dict__, **fields})
            q.check_supported()
            tau = complex(case.get("tau", 0.0))
            ops, okey = _operators_of(operators, q)
            # B (and its norm) depends on the degrees present, the equations that are on and the radial operators only
            key = (q.m, q.lmax, q.symm, q.thermal, q.heating, q.magnetic, okey)
            res = _asm.assemble(s, q, ops, bnorm=bnorm_of.get(key))
            bnorm_of[key] = res["bnorm"]
            perm, nodeptr = _chain.chain_from_params(q.N1, q.m, q.lmax, q.symm, -1, q.hydro, q.magnetic, q.thermal, q.compositional)
            s.set_chain(perm, nodeptr)
            s.factor(tau)
            lam, X, info = s.eigs(nev, which=which, target=tau, want_vectors=want_vectors, **kw)
            out.append((case, lam, X, info))
    return out


def track_mode(pp, operators, cases, tau0, nev=3, which="TM", device=0, solver_factory=None, **kw):
    """Follow ONE eigenmode through a sequence of parameter sets (the reference's `track_target = 1` mechanism:
    spin_doctor.py:323-331 writes the eigenvalue closest to the current target to the file `track_target`,
    parameters.py:316-321 reads it back as the target of the next run of the shell loop, tools/subramp.sh:59-141).
    `cases` as in `parameter_sweep` (without "tau"); the target of the first case is `tau0`, the target of every
    later case is the eigenvalue tracked in the previous one; of the `nev` pairs computed around the target the one
    closest to it is the tracked mode.  Every pencil is assembled on the GPU; a ramp in Ek, Ra or the boundary
    conditions reuses the same radial operators throughout (the reference re-runs submatrices.py and
    assemble.py at every step of the ramp).  With ``operators=None`` (or a `radial.OperatorCache`) the operators are
    generated from the parameters (kore_b200/radial.py), so a case may change the truncation too -- the usual ramp
    in Ekman number with ``N, lmax = radial.resolution_rule(Ek, m)`` at every step.

    Returns [(case, tracked eigenvalue, all eigenvalues, info)]."""
    from . import assembly as _asm
    from . import chain as _chain
    make = solver_factory or _lib.Solver
    out = []
    tau = complex(tau0)
    with make(device) as s:
        bnorm_of = {}
        for case in cases:
            q = _asm.PhysicsParams.from_dict({**pp.__dict__, **case})
            q.check_supported()
            ops, okey = _operators_of(operators, q)
            key = (q.m, q.lmax, q.symm, q.thermal, q.heating, q.magnetic, okey)
            res = _asm.assemble(s, q, ops, bnorm=bnorm_of.get(key))
            bnorm_of[key] = res["bnorm"]
            perm, nodeptr = _chain.chain_from_params(q.N1, q.m, q.lmax, q.symm, -1, q.hydro, q.magnetic, q.thermal, q.compositional)
            s.set_chain(perm, nodeptr)
            s.factor(tau)
            lam, _, info = s.eigs(nev, which=which, target=tau, want_vectors=False, **kw)
            if len(lam) == 0:
                raise RuntimeError("no converged eigenpair at %r (target %r)" % (case, tau))
            tracked = lam[int(np.argmin(np.abs(lam - tau)))]
            out.append((case, tracked, lam, info))
            tau = complex(tracked)
    return out
