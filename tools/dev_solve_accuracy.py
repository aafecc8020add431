#!/usr/bin/env python3
"""Dev: relative size of the first refinement correction of the chain solve (= forward error of the
unrefined solve) on the fixtures, the synthetic benchmark pencil and the device-assembled E = 1e-8 pencil."""
import json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case
from kore_b200 import assembly as asm, chain, lib, synthetic


def probe(s, B, tau, tag, nev=5):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(s.n) + 1j * rng.standard_normal(s.n)
    rhs = B @ v
    s.set_option(lib.OPT_REFINE, 0); x0 = s.solve(rhs)
    s.set_option(lib.OPT_REFINE, 1); x1 = s.solve(rhs)
    s.set_option(lib.OPT_REFINE, 2); x2 = s.solve(rhs)
    s.set_option(lib.OPT_REFINE, 1)
    lam, X, info = s.eigs(nev, which="TM", target=tau, ncv=0, tol=1e-12, maxit=100, true_residual=True)
    print("%-18s n=%7d  |x1-x0|/|x1| = %.2e   |x2-x1|/|x2| = %.2e | eigs auto: probe %.2e solves/applies %d/%d eigs_ms %.1f max resid %.2e"
          % (tag, s.n, np.linalg.norm(x1 - x0) / np.linalg.norm(x1), np.linalg.norm(x2 - x1) / np.linalg.norm(x2),
             info["refine_resid"], info["solve_calls"], info["op_applies"], info["eigs_ms"], float(np.max(info["resid"]))), flush=True)


for name in ["spinover", "magnetic_small", "forced_small_eig", "m0_small", "dormy", "jones"]:
    c = load_case(name)
    with lib.Solver(0) as s:
        s.set_pencil(c.A, c.B); s.set_chain(c.perm, c.nodeptr); s.factor(c.tau)
        probe(s, c.B, c.tau, name)
d = os.path.join(ROOT, "tests", "golden", "asm_E1e-8")
pj = json.load(open(os.path.join(d, "asm_params.json")))
pp = asm.PhysicsParams.from_dict(pj)
with lib.Solver(0) as s:
    asm.assemble(s, pp, asm.load_operators_npz(os.path.join(d, "operators.npz")), bnorm=pj["Bnorm"])
    import scipy.sparse as sp
    ip, ix, v = s.get_assembled("B")
    B = sp.csr_matrix((v, ix, ip), shape=(s.n, s.n))
    perm, nodeptr = chain.chain_from_params(pp.N1, pp.m, pp.lmax, pp.symm, -1, 1, 0, 0, 0)
    s.set_chain(perm, nodeptr); s.factor(1j)
    probe(s, B, 1j, "kore_E1e-8_N600", 10)
A, B, perm, nodeptr = synthetic.synthetic_pencil(600, 600)
with lib.Solver(0) as s:
    s.set_pencil(A, B); s.set_chain(perm, nodeptr); s.factor(1j)
    probe(s, B, 1j, "synthetic_P600", 10)
