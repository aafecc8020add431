#!/usr/bin/env python3
"""Randomised parity sweep of the assembly programs against the UNMODIFIED reference assembler (build
container only: needs /root/reference).  Every trial draws a parameter set (m, symmetry, boundary conditions,
thermal on/off with its heating mode and boundary conditions, inner core or full sphere, Ekman number, forcing
mode, truncation), runs the reference stages through tools/make_case.py --asm and compares what
kore_b200.assembly + the NumPy model of the kernels (tests/assembly_model.py) produce with A.npz / B.npz:
pattern and values, bit for bit.  With a third argument `magnetic` the trials are magnetic runs (any degree-1
background field, shell or full sphere, insulating boundaries, with or without the heat equation) and the bar is the rounding-level one of
tests/test_zz_assembly_extensions.py: B bit for bit, every block of A within 1e-13 of its largest entry.
`anelastic` instead draws density-stratified runs (bit for bit again; one in three with a viscosity profile,
whose nested sums put its two viscous blocks at the rounding-level bar).  In every trial the radial operators are also
generated from the parameters alone (kore_b200/radial.py, with the trial's radProfs tables) and compared with the ones
the reference's submatrices.py wrote: label set and every entry, bit for bit.
Usage: tools/fuzz_assembly.py [ntrials] [seed] [magnetic | anelastic]."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import assembly_model as am  # noqa: E402
from kore_b200 import assembly as asm  # noqa: E402
from kore_b200 import radial  # noqa: E402


APPEND = {}  # overrides of a trial -> source appended to its parameters.py (make_case.py --append-params)
PROFILES = {}  # overrides of a trial -> source appended to its radial_profiles.py (make_case.py --profiles)


def load(fn):
    z = np.load(fn)
    M = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    M.sort_indices()
    return M


def same(M, R):
    return np.array_equal(M.indptr, R.indptr) and np.array_equal(M.indices, R.indices) and np.array_equal(M.data, R.data)


def block_relative_error(A, A_ref, N1):
    D, R = (A - A_ref).tocoo(), A_ref.tocoo()
    nbr = A.shape[0] // N1
    mx = np.zeros((nbr, nbr))
    np.maximum.at(mx, (R.row // N1, R.col // N1), np.abs(R.data))
    er = np.zeros((nbr, nbr))
    np.maximum.at(er, (D.row // N1, D.col // N1), np.abs(D.data))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(mx > 0, er / np.where(mx > 0, mx, 1.0), np.where(er > 0, np.inf, 0.0))


def draw_magnetic(rng):
    thermal = int(rng.integers(0, 2))
    full = int(rng.integers(0, 4) == 0)
    params = "tests/dormy2004/params.dormy04" if thermal else "tests/spinover/params.spinover"
    m = int(rng.integers(0, 5))
    nl = 2 * int(rng.integers(4, 9))
    fields = ["axial", "G21 dipole", "Luo_S1", "FDM"] + ([] if full else ["dipole"])
    ov = ["magnetic=1", "B0='%s'" % rng.choice(fields), "m=%d" % m, "symm=%d" % rng.choice([-1, 1]),
          "N=%d" % (rng.choice([16, 20, 24]) * (2 if full else 1)), "lmax=%d" % (nl + m - 1), "Ek=%g" % (10.0 ** rng.uniform(-5, -2)),
          "ricb=%s" % ("0" if full else "%.3f" % rng.uniform(0.2, 0.7)), "bci=%d" % rng.integers(0, 2), "bco=%d" % rng.integers(0, 2),
          "Lambda=%g" % (10.0 ** rng.uniform(-2, 0.5)), "Pm=%g" % (10.0 ** rng.uniform(-4, -1)), "forcing=0"]
    if thermal:
        ov += ["heating='%s'" % rng.choice(["internal"] if full else ["differential", "internal"]), "Ra_gap=%g" % (10.0 ** rng.uniform(4, 7))]
    if rng.integers(0, 3) == 0:  # thin conducting layer below the mantle
        ov += ["mantle='TWA'", "c_cmb=%.3f" % rng.uniform(0, 0.5), "c1_cmb=%.3f" % rng.uniform(0, 0.5), "mu=%.2f" % rng.uniform(0.5, 2)]
    if not full and rng.integers(0, 3) == 0:  # ... and on top of the inner core
        ov += ["innercore='TWA'", "c_icb=%.3f" % rng.uniform(0, 0.5), "c1_icb=%.3f" % rng.uniform(0, 0.5)]
    if not full and rng.integers(0, 4) == 0:  # libration forcing of a magnetic run (boundary flow, no-slip)
        mm = int(rng.choice([0, 2]))
        ov = [o for o in ov if not o.startswith(("m=", "symm=", "lmax=", "bci=", "bco=", "forcing="))]
        ov += ["m=%d" % mm, "symm=1", "lmax=%d" % (nl + mm - 1), "bci=1", "bco=1", "forcing=7",
               "forcing_frequency=%.3f" % rng.uniform(-1.5, 1.5), "forcing_amplitude_icb=%.2f" % rng.uniform(0, 1)]
    if not full and rng.integers(0, 3) == 0:  # density-stratified background
        params = "tests/dormy2004/params.dormy04"
        ov = [o.replace("'dipole'", "'axial'") for o in ov if not o.startswith("heating=")]  # the reference fails with a dipole
        ov += ["anelastic=1", "thermal=%d" % thermal, "Nrho=%.2f" % rng.uniform(0.5, 4.0), "polind=%.2f" % rng.uniform(1.0, 3.0)]
    return params, ov


def draw_anelastic(rng):
    thermal = int(rng.integers(0, 2))
    m = int(rng.integers(0, 5))
    nl = 2 * int(rng.integers(4, 9))
    ov = ["anelastic=1", "thermal=%d" % thermal, "m=%d" % m, "symm=%d" % rng.choice([-1, 1]), "N=%d" % rng.choice([20, 24, 28]),
          "lmax=%d" % (nl + m - 1), "Ek=%g" % (10.0 ** rng.uniform(-5, -2)), "ricb=%.3f" % rng.uniform(0.2, 0.7),
          "bci=%d" % rng.integers(0, 2), "bco=%d" % rng.integers(0, 2), "Nrho=%.2f" % rng.uniform(0.5, 4.0),
          "polind=%.2f" % rng.uniform(1.0, 3.0), "forcing=0"]
    if thermal:
        ov += ["bci_thermal=%d" % rng.integers(0, 2), "bco_thermal=%d" % rng.integers(0, 2), "Ra_gap=%g" % (10.0 ** rng.uniform(4, 7))]
    if rng.integers(0, 3) == 0:  # a viscosity profile of the run's own (the shipped one is identically zero)
        ov.append("variable_viscosity=1")
        PROFILES[tuple(ov)] = "def viscosity(r):\n    return %.2f + %.2f*r**2" % (rng.uniform(0.5, 2), rng.uniform(-0.4, 0.8))
    return "tests/dormy2004/params.dormy04", ov


def draw(rng):
    thermal = int(rng.integers(0, 2))
    full = int(rng.integers(0, 4) == 0)
    params = "tests/dormy2004/params.dormy04" if thermal else "tests/spinover/params.spinover"
    m = int(rng.integers(0, 6))
    symm = int(rng.choice([-1, 1]))
    N = int(rng.choice([16, 20, 24, 28])) * (2 if full else 1)
    nl = 2 * int(rng.integers(4, 10))
    lmax = nl + m - 1
    ov = ["m=%d" % m, "symm=%d" % symm, "N=%d" % N, "lmax=%d" % lmax, "Ek=%g" % (10.0 ** rng.uniform(-6, -2)),
          "bco=%d" % rng.integers(0, 2)]
    if full:
        ov.append("ricb=0")
        if rng.integers(0, 4) == 0:  # inviscid full sphere
            ov = [o for o in ov if not o.startswith("Ek=")] + ["Ek=0"]
    else:
        ov += ["ricb=%.3f" % rng.uniform(0.1, 0.8), "bci=%d" % rng.integers(0, 2)]
    if thermal:
        ov += ["heating='%s'" % rng.choice(["differential", "internal", "two zone", "user defined"] if not full else ["internal", "two zone"]),
               "bco_thermal=%d" % rng.integers(0, 2), "Ra_gap=%g" % (10.0 ** rng.uniform(4, 7))]
        if not full:
            ov.append("bci_thermal=%d" % rng.integers(0, 2))
        if rng.integers(0, 6) == 0:
            ov.append("ThermaD=0")
    append_src = None
    if rng.integers(0, 5) == 0:  # composition equation (needs the OmgTau the shipped params files leave commented out)
        ov += ["compositional=1", "comp_background='%s'" % rng.choice(["internal"] if full else ["internal", "differential"]),
               "bco_compositional=%d" % rng.integers(0, 2)] + ([] if full else ["bci_compositional=%d" % rng.integers(0, 2)])
        append_src = "OmgTau = %.2f\nSchmidt = %.2f\nBV2_comp = -%.3g * Ek**2 / Schmidt" % (
            rng.uniform(0.5, 2), rng.uniform(0.1, 5), 10.0 ** rng.uniform(4, 7))
    forcing = int(rng.choice([0, 0, 0, 10, 7]))
    if forcing == 10 and symm == 1 and m > 0:
        ov += ["forcing=10", "forcing_frequency=%.3f" % rng.uniform(-1.5, 1.5)]
    elif forcing == 7 and not full:
        ov = [o for o in ov if not o.startswith(("m=", "symm=", "lmax=", "bci=", "bco="))]
        m = int(rng.choice([0, 2]))
        ov += ["m=%d" % m, "symm=1", "lmax=%d" % (nl + m - 1), "bci=1", "bco=1", "forcing=7",
               "forcing_frequency=%.3f" % rng.uniform(-1.5, 1.5), "forcing_amplitude_icb=%.2f" % rng.uniform(0, 1)]
    if append_src:
        APPEND[tuple(ov)] = append_src
    return params, ov


def main():
    ntrials = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    magnetic = len(sys.argv) > 3 and sys.argv[3] == "magnetic"
    anelastic = len(sys.argv) > 3 and sys.argv[3] == "anelastic"
    bad = 0
    for t in range(ntrials):
        params, ov = draw_magnetic(rng) if magnetic else (draw_anelastic(rng) if anelastic else draw(rng))
        out = "/tmp/asmfuzz_%d_%d" % (os.getpid(), t)
        shutil.rmtree(out, ignore_errors=True)
        prof = PROFILES.get(tuple(ov))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_case.py"), "--params", params, "--out", out,
                            "--asm"] + (["--profiles", prof] if prof else [])
                           + (["--append-params", APPEND[tuple(ov)]] if tuple(ov) in APPEND else []) + ov,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 or not os.path.exists(os.path.join(out, "A.npz")):
            print("trial %d: the reference itself failed on %s (skipped)" % (t, " ".join(ov)))
            continue
        pp = asm.PhysicsParams.from_dict(json.load(open(os.path.join(out, "asm_params.json"))))
        ops = asm.load_operators_npz(os.path.join(out, "operators.npz"))
        pA = asm.build_program_A(pp, ops)
        ok = True
        rpf = os.path.join(out, "radprofs.npz")
        mine = radial.radial_operators(pp, radprofs=dict(np.load(rpf)) if os.path.exists(rpf) else None)
        ops_ok = sorted(mine) == sorted(ops) and all(np.array_equal(mine[k].toarray(), ops[k].toarray()) for k in ops)
        ok &= ops_ok
        if pp.forcing == 0:
            pB = asm.build_program_B(pp, ops)
            bn = asm.frobenius_norm(am.evaluate(pB).data)
            ok &= same(am.evaluate(pB.with_final_scale(1. / bn)), load(os.path.join(out, "B.npz")))
            pA = pA.with_final_scale(1. / bn)
        else:
            z = np.load(os.path.join(out, "B_forced.npz"))
            ref = np.asarray(sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"])).todense()).ravel()
            ok &= np.array_equal(asm.forcing_vector(pp), ref)
        note = "bit-identical"
        if magnetic or pp.variable_viscosity:
            rel = block_relative_error(am.evaluate(pA), load(os.path.join(out, "A.npz")), pp.N1).max()
            ok &= rel <= 1e-13
            note = "%s bit-identical, A within %.1e of the block maxima" % ("B" if pp.forcing == 0 else "forcing vector", rel)
        else:
            ok &= same(am.evaluate(pA), load(os.path.join(out, "A.npz")))
        bad += not ok
        note += "; %d radial operators %s" % (len(ops), "bit-identical" if ops_ok else "DIFFER")
        print("trial %d: %s  %s" % (t, note if ok else "MISMATCH " + note, " ".join(ov)), flush=True)
        shutil.rmtree(out, ignore_errors=True)
    print("%d mismatches in %d trials" % (bad, ntrials))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
