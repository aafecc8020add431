#!/usr/bin/env python3
"""Build-container check of INTEGRATION.md C': bin/spin_doctor.py with its ONE call of utils4pp.diagnose
(spin_doctor.py:149) replaced by kore_b200.diagnostics writes the same flow.dat / thermal.dat / compositional.dat
as the unmodified script.  Needs /root/reference; the kernel is stood in for by its NumPy model
(tests/diag_model.py behind Solver.diagnose, as in tests/test_diagnostics.py) because this container has no GPU --
what is checked is the edit (argument order, array shapes, the script's own sums downstream of the call), the
kernel itself is compared with utils4pp.diagnose by the GPU tests.
Usage: tools/check_spin_doctor_dropin.py   (writes a log to stdout)"""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_spin_doctor_fixtures import SD_CASES, WORKER  # noqa: E402
from make_golden import CASES  # noqa: E402
from make_asm_fixtures import NEW_CASES  # noqa: E402

CALL = "[ udgn, bdgn, tdgn, cdgn ] = upp.diagnose( u_sol2, b_sol2, t_sol2, c_sol2, par.ricb, ut.rcmb, int(ncpus) )"
EDIT = '''x = rflow + 1j*iflow
        if par.thermal:       x = np.r_[x, rthm + 1j*ithm]
        if par.compositional: x = np.r_[x, rcmp + 1j*icmp]
        geom = (par.N, par.lmax, par.m, par.symm, par.ricb)
        bdgn = 0
        if par.compositional:
            udgn, tdgn, cdgn, _ = dg.diagnose_double_diffusive(kb_solver, x, *geom, thermal=par.thermal, heating=par.heating,
                                                               comp_background=par.comp_background)
            udgn, tdgn, cdgn = udgn[0], tdgn[0], cdgn[0]
        else:
            udgn, tdgn, _ = dg.diagnose(kb_solver, x, *geom, thermal=par.thermal, heating=par.heating)
            udgn, tdgn, cdgn = udgn[0], tdgn[0], 0'''
HEAD = '''import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from kore_b200 import diagnostics as dg
from test_diagnostics import ModelSolver
kb_solver = ModelSolver()
''' % (ROOT, os.path.join(ROOT, "tests"))


def main():
    bad = 0
    for name, (base, append, more) in SD_CASES.items():
        params, ov = CASES[base] if base in CASES else NEW_CASES[base]
        out = "/tmp/sddrop_" + name
        shutil.rmtree(out, ignore_errors=True)
        log = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "make_case.py"), "--params", params,
                                       "--out", out, "--keep", "--append-params", append] + list(ov) + more).decode()
        work = [ln.split("scratch kept at ")[1].strip() for ln in log.splitlines() if "scratch kept at" in ln][0]
        src = open(os.path.join(work, "bin", "spin_doctor.py")).read()
        assert src.count(CALL) == 1
        open(os.path.join(work, "bin", "spin_doctor_kb.py"), "w").write(
            src.replace("import utils4pp as upp\n", "import utils4pp as upp\n" + HEAD).replace(CALL, EDIT))
        subprocess.check_call([sys.executable, "-c", WORKER % {"oracle": os.path.join(ROOT, "oracle")}], cwd=work)
        files = {}
        for script in ("spin_doctor.py", "spin_doctor_kb.py"):
            for fn in ("flow.dat", "thermal.dat", "compositional.dat", "params.dat", "eigenvalues.dat"):
                if os.path.exists(os.path.join(work, fn)):
                    os.remove(os.path.join(work, fn))
            subprocess.run([sys.executable, "bin/" + script, "4"], cwd=work, check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
            files[script] = {fn: np.loadtxt(os.path.join(work, fn), ndmin=2) for fn in ("flow.dat", "thermal.dat", "compositional.dat")
                             if os.path.exists(os.path.join(work, fn))}
        for fn, ref in files["spin_doctor.py"].items():
            got = files["spin_doctor_kb.py"][fn]
            scale = np.maximum(np.max(np.abs(ref), axis=0), 1e-300)
            err = float(np.max(np.abs(got - ref) / scale))
            ok = got.shape == ref.shape and err < 1e-9
            bad += not ok
            print("%-22s %-18s %s rows x %d columns, max |edited - unmodified| / column max = %.2e  %s"
                  % (name, fn, ref.shape[0], ref.shape[1], err, "ok" if ok else "MISMATCH"))
        shutil.rmtree(work, ignore_errors=True)
    return bad


if __name__ == "__main__":
    sys.exit(main())
