#!/usr/bin/env python3
"""Fixtures of the diagnostics tests (tests/test_diagnostics.py): what the UNMODIFIED reference
post-processing (bin/utils4pp.py: expand_reshape_sol + diagnose, the calls spin_doctor.py:119-147
makes) computes for an eigenvector of a reference-assembled pencil.

Needs /root/reference (build container only).  Per case: the eigenpair of largest real part among the
oracle's (oracle/kore_oracle.py on the case's A.npz / B.npz), and the per-degree integrals diagnose
returns (flow: kinetic energy, kinetic dissipation, internal dissipation, Lorentz, buoyancy, compositional
power; thermal: energy, dissipation, advection).  Stored as tests/golden/<case>/diagnostics.npz.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from make_golden import CASES  # noqa: E402
from make_asm_fixtures import NEW_CASES, APPEND  # noqa: E402

DIAG_CASES = ["spinover", "dormy", "jones", "asm_compositional", "magnetic_small"]

WORKER = r'''
import sys, json
import numpy as np
sys.path.insert(0, "bin"); sys.path.insert(0, %(oracle)r)
import warnings; warnings.simplefilter("ignore")
import parameters as par, utils as ut, utils4pp as upp
import kore_oracle as ko
A, B = ko.load_csr("A.npz"), ko.load_csr("B.npz")
lam, X, info = ko.eigs(A, B, par.tau, par.nev, par.which_eigenpairs)
i = int(np.argmax(lam.real))
x = X[:, i]
n = ut.n
u2 = upp.expand_reshape_sol(x[:2 * n], par.symm)
b2 = upp.expand_reshape_sol(x[2 * n:4 * n], ut.bsymm) if par.magnetic else 0
t0 = (2 + 2 * par.magnetic) * n
t2 = upp.expand_reshape_sol(x[t0:t0 + n], par.symm) if par.thermal else 0
c0 = (2 + 2 * par.magnetic + par.thermal) * n
c2 = upp.expand_reshape_sol(x[c0:c0 + n], par.symm) if par.compositional else 0
udgn, bdgn, tdgn, cdgn = upp.diagnose(u2, b2, t2, c2, par.ricb, ut.rcmb, 4)
extra = dict(comp=np.asarray(cdgn, dtype=float)) if par.compositional else {}
if par.magnetic:
    extra.update(magnetic=np.asarray(bdgn, dtype=float), bsymm=np.array([ut.bsymm]))
np.savez_compressed("diagnostics.npz", x=x, lam=np.array([lam[i]]), flow=np.asarray(udgn, dtype=float),
                    thermal=np.asarray(tdgn, dtype=float) if par.thermal else np.zeros((0, 3)), **extra)
print("ok", lam[i], np.sum(udgn, 0))
'''


def main():
    only = sys.argv[1:]
    for name in DIAG_CASES:
        if only and name not in only:
            continue
        params, ov = CASES[name] if name in CASES else NEW_CASES[name]
        if name in APPEND:
            ov = ["--append-params", APPEND[name]] + list(ov)
        out = "/tmp/diagfix_" + name
        shutil.rmtree(out, ignore_errors=True)
        log = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "make_case.py"),
                                       "--params", params, "--out", out, "--keep"] + ov).decode()
        work = [ln.split("scratch kept at ")[1].strip() for ln in log.splitlines() if "scratch kept at" in ln][0]
        r = subprocess.run([sys.executable, "-c", WORKER % {"oracle": os.path.join(ROOT, "oracle")}], cwd=work,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(name, r.stdout.strip().splitlines()[-1])
        if r.returncode != 0:
            print(r.stdout)
            raise SystemExit(1)
        shutil.copy(os.path.join(work, "diagnostics.npz"), os.path.join(HERE, name, "diagnostics.npz"))
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
