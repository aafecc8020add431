"""Host logic of kore_b200.sweep.parameter_sweep (CPU): the cases are dealt to the ranks, every case gets its own
assembly program / chain / shift, and B's norm is reused between cases that share B.  The solver is a stand-in
that evaluates the programs with the NumPy model of the assembly kernels and solves with the CPU oracle."""
import json
import os

import numpy as np

import assembly_model as am
import kore_oracle as ko
from conftest import GOLDEN
from kore_b200 import assembly as asm, sweep


class OracleSolver:
    log = []

    def __init__(self, device=0):
        self.M = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def assemble(self, progA, progB=None):
        OracleSolver.log.append(("assemble", progA is not None, progB is not None))
        self.M = {}
        if progA is not None:
            self.M["A"] = am.evaluate(progA)
        if progB is not None:
            self.M["B"] = am.evaluate(progB)
        self.n = (progA if progA is not None else progB).n

    def get_assembled(self, which="A"):
        M = self.M[which]
        return M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data

    def set_chain(self, perm, nodeptr):
        assert len(perm) == self.n and nodeptr[-1] == self.n
        OracleSolver.log.append(("chain", len(nodeptr) - 1))

    def factor(self, tau):
        self.tau = tau

    def eigs(self, nev, which="TM", target=None, want_vectors=False, **kw):
        lam, X, info = ko.eigs(self.M["A"], self.M["B"], self.tau, nev, which)
        return lam, (X if want_vectors else None), dict(nconv=len(lam))


def test_parameter_sweep_over_m_and_symmetry():
    d = os.path.join(GOLDEN, "m0_small")
    pp = asm.PhysicsParams.from_dict(json.load(open(os.path.join(d, "asm_params.json"))))
    ops = asm.load_operators_npz(os.path.join(d, "operators.npz"))
    # the m0_small fixture itself (m = 0, symm = 1, lmax = 39) and two other members of the family
    cases = [{"m": 0, "symm": 1, "lmax": 39, "tau": 1j}, {"m": 2, "symm": 1, "lmax": 41, "tau": 1j},
             {"m": 2, "symm": 1, "lmax": 41, "tau": 0.5j}, {"m": 1, "symm": -1, "lmax": 40, "tau": 1j}]
    OracleSolver.log = []
    out = sweep.parameter_sweep(pp, ops, cases, nev=3, solver_factory=OracleSolver)
    assert [c for c, *_ in out] == cases
    # the norm of B is assembled once per distinct (m, lmax, symm): cases 1 and 2 share it
    b_only = [e for e in OracleSolver.log if e == ("assemble", False, True)]
    assert len(b_only) == 3
    # case 0 is the committed fixture: same eigenvalues as the oracle on the reference-assembled matrices
    A = ko.load_csr(os.path.join(d, "A.npz"))
    B = ko.load_csr(os.path.join(d, "B.npz"))
    lam_ref, _, _ = ko.eigs(A, B, 1j, 3, "TM")
    assert np.max(np.abs(np.sort_complex(out[0][1]) - np.sort_complex(lam_ref))) < 1e-12
    # round-robin over two ranks
    r0 = sweep.parameter_sweep(pp, ops, cases, nev=3, rank=0, world=2, solver_factory=OracleSolver)
    r1 = sweep.parameter_sweep(pp, ops, cases, nev=3, rank=1, world=2, solver_factory=OracleSolver)
    assert [c for c, *_ in r0] == cases[0::2] and [c for c, *_ in r1] == cases[1::2]
    for (c, lam, _, _), full in zip(r0 + r1, [out[0], out[2], out[1], out[3]]):
        assert np.allclose(np.sort_complex(lam), np.sort_complex(full[1]), rtol=0, atol=1e-13)


def test_track_mode_follows_the_spinover_mode_down_in_ekman_number():
    # the spin-over mode of the spin-over fixture's physics at a small truncation, Ek from 1e-3 down to 4e-4: the
    # same radial operators at every step, only the viscous factor changes; the tracked eigenvalue is the mode
    # closest to the previous one and moves smoothly (damping ~ Ek^(1/2))
    d = os.path.join(GOLDEN, "asm_mixed_bc")  # N = 24 operators (hydro)
    ops = asm.load_operators_npz(os.path.join(d, "operators.npz"))
    base = asm.PhysicsParams.from_dict(json.load(open(os.path.join(GOLDEN, "spinover", "asm_params.json"))))
    pp = asm.PhysicsParams.from_dict({**base.__dict__, "N": 24, "lmax": 24})
    eks = [1e-3, 8e-4, 6e-4, 4e-4]
    cases = [{"Ek": e, "ViscosD": e} for e in eks]
    lam_spin = complex(*json.load(open(os.path.join(GOLDEN, "spinover", "meta.json")))["reference_golden"]["eig"])
    out = sweep.track_mode(pp, ops, cases, lam_spin, nev=3, solver_factory=OracleSolver)
    tracked = np.array([t for _, t, _, _ in out])
    # first step: the N = 24 truncation still has the N = 68 fixture's spin-over mode to 1e-5
    assert abs(tracked[0] - lam_spin) < 1e-5 * abs(lam_spin)
    # the damping falls with Ek, smoothly; the frequency stays near 1
    assert np.all(np.diff(tracked.real) > 0) and np.all(tracked.real < 0)
    assert np.all(np.abs(np.diff(tracked)) < 0.03) and np.all(np.abs(tracked.imag - 1.0) < 0.02)
    ratio = tracked.real[-1] / tracked.real[0]
    assert 0.55 < ratio < 0.75  # (4e-4 / 1e-3)^(1/2) = 0.63


def test_track_mode_with_kores_resolution_rule_from_the_parameters_alone():
    # no operator files: every step takes N, lmax from Kore's own rule (parameters.py:6-16, 296-301) and its radial
    # operators from kore_b200/radial.py.  Ek = 1e-3 gives N = 68, lmax = 64 -- the spin-over fixture itself, so the
    # first tracked eigenvalue is the reference's golden value (tests/spinover/reference.eig, rtol 1e-8 there)
    from kore_b200 import radial
    base = asm.PhysicsParams.from_dict(json.load(open(os.path.join(GOLDEN, "spinover", "asm_params.json"))))
    assert radial.resolution_rule(1e-3, 1) == (68, 64) and radial.resolution_rule(1e-8, 1) == (676, 672)
    assert radial.resolution_rule(1e-7, 1)[0] == 428 and radial.resolution_rule(1e-9, 1)[0] == 1072
    cases = []
    for ek in (1e-3, 7e-4):
        N, lmax = radial.resolution_rule(ek, base.m)
        cases.append({"Ek": ek, "ViscosD": ek, "N": N, "lmax": lmax})
    assert (cases[1]["N"], cases[1]["lmax"]) == (72, 72)
    golden = complex(*json.load(open(os.path.join(GOLDEN, "spinover", "meta.json")))["reference_golden"]["eig"])
    cache = radial.OperatorCache()
    out = sweep.track_mode(base, cache, cases, golden, nev=3, solver_factory=OracleSolver)
    assert len(cache.store) == 2                              # one set of operators per truncation
    assert abs(out[0][1] - golden) <= 1e-8 * abs(golden)
    assert out[1][1].real > out[0][1].real and abs(out[1][1] - out[0][1]) < 0.03
