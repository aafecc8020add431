#!/usr/bin/env python3
"""CPU oracle (SciPy SuperLU + ARPACK, oracle/kore_oracle.py) on the REFERENCE-ASSEMBLED Kore pencil at
the benchmark's size (hydro, m = 1, symm = -1, Ek = 1e-8, N = lmax = 600, n = 360 000; A.npz / B.npz
from tools/make_case.py, i.e. the unmodified bin/assemble.py): nev = 10 pairs nearest tau = 1j.
Needs the directory make_case.py wrote (258 MB, not committed); writes
tests/golden/asm_E1e-8/oracle_eigs.json.  Run once in the build container (tens of minutes)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kore_oracle as ko  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/tmp/asm_big600"
A, B = ko.load_csr(os.path.join(src, "A.npz")), ko.load_csr(os.path.join(src, "B.npz"))
tau = 1j
t0 = time.perf_counter()
op = ko.ShiftInvert(A, B, tau)
t_factor = time.perf_counter() - t0
t0 = time.perf_counter()
lam, X, info = ko.eigs(A, B, tau, 10, "TM", ncv=25, tol=1e-12, op=op)
t_eigs = time.perf_counter() - t0
res = ko.residuals(A, B, lam, X)
out = {"tau": [tau.real, tau.imag], "nev": 10, "ncv": 25, "tol": 1e-12, "which": "TM",
       "eigenvalues": [[z.real, z.imag] for z in lam], "residuals": [float(r) for r in res],
       "napply": int(info["napply"]), "factor_s": t_factor, "eigs_s": t_eigs, "host_cpus": os.cpu_count(),
       "note": "SciPy splu (COLAMD, serial SuperLU) + ARPACK on the reference-assembled pencil"}
with open(os.path.join(ROOT, "tests", "golden", "asm_E1e-8", "oracle_eigs.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
