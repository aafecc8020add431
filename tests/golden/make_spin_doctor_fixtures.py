#!/usr/bin/env python3
"""Fixtures of the post-processing file contract (tests/test_diagnostics.py): what the UNMODIFIED
bin/spin_doctor.py writes -- flow.dat, thermal.dat, compositional.dat -- from the files bin/solve.py leaves
behind (eigenvalues0.dat, real_/imag_*.field, timing.dat), for eigenvectors of reference-assembled pencils.

Needs /root/reference (build container only).  spin_doctor.py reads `par.OmgTau`, which every shipped
parameter file leaves commented out (parameters.py:268-271): it is appended here (OmgTau = 1), as the
double-diffusive assembly fixture already does.  Per case the oracle's eigenpairs (oracle/kore_oracle.py on the
case's A.npz / B.npz) are written as the solve.py of the reference writes them (solve.py:275-306) and
spin_doctor.py is run on them in the scratch run directory.  Stored as tests/golden/<case>/spin_doctor.npz:
X (solutions), lam, the parameters spin_doctor.py used (a JSON string) and the rows of the three files.

Cases: the double-diffusive pencil (m = 3: no torque), the spin-over pencil (m = 1 antisymmetric: equatorial
viscous torque on the mantle), the m = 0 symmetric pencil (axial viscous torques on mantle and inner core).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from make_golden import CASES  # noqa: E402
from make_asm_fixtures import NEW_CASES, APPEND  # noqa: E402

# spin_doctor.py:290 reads `ut.heating`, which utils.py defines for thermal runs only: a hydrodynamic run dies there
# before any file is written, so the two torque cases carry the heat equation as well (thermal=1; stored under
# their own names with the pencil's solutions, no A.npz / B.npz needed by the tests)
SD_CASES = {"asm_compositional": ("asm_compositional", APPEND["asm_compositional"], []),
            "sd_spinover_thermal": ("spinover", "OmgTau = 1", ["thermal=1", "N=32", "lmax=32"]),
            "sd_m0_thermal": ("asm_thermal_m0", "OmgTau = 1\ntau = 0.002 - 0.01j", [])}

WORKER = r'''
import sys, os
import numpy as np
sys.path.insert(0, "bin"); sys.path.insert(0, %(oracle)r)
import warnings; warnings.simplefilter("ignore")
import parameters as par, utils as ut
import kore_oracle as ko
A, B = ko.load_csr("A.npz"), ko.load_csr("B.npz")
lam, X, info = ko.eigs(A, B, par.tau, par.nev, par.which_eigenpairs)
n = ut.n
np.savetxt("eigenvalues0.dat", np.c_[lam.real, lam.imag])
np.savetxt("timing.dat", [1.0])
o = 0
for name, w in (("flow", 2 * par.hydro), ("magnetic", 2 * par.magnetic), ("temperature", par.thermal),
                ("composition", par.compositional)):
    if w:
        np.savetxt("real_%%s.field" %% name, X[o:o + w * n].real)
        np.savetxt("imag_%%s.field" %% name, X[o:o + w * n].imag)
        o += w * n
import json
pars = dict(N=int(par.N), lmax=int(par.lmax), m=int(par.m), symm=int(par.symm), ricb=float(par.ricb), n=int(n),
            thermal=int(par.thermal), compositional=int(par.compositional), heating=str(par.heating),
            comp_background=str(getattr(par, "comp_background", "differential")), Ek=float(par.Ek),
            OmgTau=float(par.OmgTau), BV2=float(par.BV2), BV2_comp=float(par.BV2_comp), Etherm=float(par.Etherm),
            Ecomp=float(par.Ecomp))
np.savez("sd_inputs.npz", X=X, lam=lam, params=json.dumps(pars))
'''


def main():
    import numpy as np
    only = sys.argv[1:]
    for name, (base, append, more) in SD_CASES.items():
        if only and name not in only:
            continue
        params, ov = CASES[base] if base in CASES else NEW_CASES[base]
        ov = list(ov) + more
        out = "/tmp/sdfix_" + name
        shutil.rmtree(out, ignore_errors=True)
        log = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "make_case.py"), "--params", params,
                                       "--out", out, "--keep", "--append-params", append] + list(ov)).decode()
        work = [ln.split("scratch kept at ")[1].strip() for ln in log.splitlines() if "scratch kept at" in ln][0]
        for cmd in ([sys.executable, "-c", WORKER % {"oracle": os.path.join(ROOT, "oracle")}],
                    [sys.executable, "bin/spin_doctor.py", "4"]):
            r = subprocess.run(cmd, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                print(r.stdout)
                raise SystemExit(1)
        print(r.stdout)
        z = np.load(os.path.join(work, "sd_inputs.npz"))
        files = {}
        for fn in ("flow", "thermal", "compositional"):
            p = os.path.join(work, fn + ".dat")
            if os.path.exists(p):
                files[fn + "_dat"] = np.loadtxt(p, ndmin=2)
        os.makedirs(os.path.join(HERE, name), exist_ok=True)
        np.savez_compressed(os.path.join(HERE, name, "spin_doctor.npz"), X=z["X"], lam=z["lam"], params=z["params"], **files)
        print(name, {k: v.shape for k, v in files.items()})
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
