#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

Needs /root/reference (only available in the build container): runs the
UNMODIFIED reference assembler through tools/make_case.py, stores the matrices
compressed, and stores what the CPU oracle (oracle/kore_oracle.py, SciPy
SuperLU + ARPACK on the explicit shift-invert operator) computes on them.
The GPU box has no /root/reference; tests read only what this script wrote.

Cases (SURVEY.md 8d "parity inputs"):
  spinover        tests/spinover/params.spinover            (C1, golden reference.eig)
  dormy           tests/dormy2004/params.dormy04            (C2, golden reference.dormy04)
  jones           tests/jones2000/params.jones at Ra_c      (golden reference.jones)
  magnetic_small  default params + magnetic dipole, N=40    (C4 structure, reduced size)
  forced_small    default params + forcing=7, m=2, N=40     (C5 structure, reduced size)
  forced_small_eig  same physics, forcing=0 (A_eig, B_eig for the omega sweep identity)
  m0_small        m=0, symm=1, N=40                         (ll = 1..lmax+1 layout)

dormy additionally gets A1.npz = dA/dRa_gap (the buoyancy entries, 84 656 nonzeros) from two more
runs of the reference assembler at Ra_gap = 1.6e6 and 1.7e6, for the critical-Rayleigh search
(kore_b200/rac.py; find_Rac.py re-assembles at every trial Ra, A is affine in Ra_gap).
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import kore_oracle as ko  # noqa: E402

CASES = {
    "spinover": ("tests/spinover/params.spinover", []),
    "dormy": ("tests/dormy2004/params.dormy04", []),
    "jones": ("tests/jones2000/params.jones", ["Ra_gap=4669860.0"]),
    "magnetic_small": ("tests/spinover/params.spinover", ["magnetic=1", "B0='dipole'", "N=40"]),
    "forced_small": ("tests/spinover/params.spinover",
                     ["forcing=7", "m=2", "symm=1", "forcing_amplitude_icb=1.0", "N=40"]),
    "forced_small_eig": ("tests/spinover/params.spinover", ["m=2", "symm=1", "N=40"]),
    "m0_small": ("tests/spinover/params.spinover", ["m=0", "symm=1", "N=40"]),
}

# the reference's own goldens (facts quoted from its test data, not code):
REFERENCE_GOLDENS = {
    "spinover": {"source": "tests/spinover/reference.eig:1",
                 "eig": [-1.019947097016500742e-01, 1.005733690103750355e+00], "rtol": 1e-8},
    "dormy": {"source": "tests/dormy2004/reference.dormy04:1",
              "Ek": 2.000e-05, "ricb": 0.35, "Ra_c": 1.65404e+06, "m": 9, "omega_c": -1.10162e-02},
    "jones": {"source": "tests/jones2000/reference.jones:1",
              "Ek": 6.325e-05, "ricb": 0.00, "Ra_c": 4.66986e+06, "m": 9, "omega_c": -1.93444e-02},
}


# Ra_gap of the committed A.npz (tests/dormy2004/params.dormy04:177)
RA_GAP_REF = {"dormy": 1654042.168683}


def recompress(src, dst):
    z = np.load(src)
    np.savez_compressed(dst, **{k: z[k] for k in z.files})


def make_rayleigh_slope(name, params, out, Ra_ref):
    """A1 = (A(Ra_hi) - A(Ra_lo)) / (Ra_hi - Ra_lo), checked against the committed A(Ra_ref)."""
    lo, hi = 1.6e6, 1.7e6
    mats = {}
    for Ra in (lo, hi):
        tmp = "/tmp/golden_%s_ra%d" % (name, int(Ra))
        shutil.rmtree(tmp, ignore_errors=True)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_case.py"),
                               "--params", params, "--out", tmp, "Ra_gap=%r" % Ra])
        mats[Ra] = ko.load_csr(os.path.join(tmp, "A.npz"))
    A1 = (mats[hi] - mats[lo]) / (hi - lo)
    A1.eliminate_zeros()
    A1 = A1.tocsr()
    A1.sort_indices()
    Aref = ko.load_csr(os.path.join(out, "A.npz"))
    err = abs(mats[lo] + (Ra_ref - lo) * A1 - Aref).max() / abs(Aref).max()
    assert err < 1e-15, err
    np.savez_compressed(os.path.join(out, "A1.npz"), data=A1.data, indices=A1.indices,
                        indptr=A1.indptr, shape=np.array(A1.shape))
    print(name, "A1 nnz=%d affine check %.1e" % (A1.nnz, err))


def main():
    only = sys.argv[1:]
    for name, (params, ov) in CASES.items():
        if only and name not in only:
            continue
        out = os.path.join(HERE, name)
        tmp = "/tmp/golden_" + name
        shutil.rmtree(tmp, ignore_errors=True)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_case.py"),
                               "--params", params, "--out", tmp] + ov)
        os.makedirs(out, exist_ok=True)
        for fn in ("A.npz", "B.npz", "B_forced.npz"):
            if os.path.exists(os.path.join(tmp, fn)):
                recompress(os.path.join(tmp, fn), os.path.join(out, fn))
        meta = json.load(open(os.path.join(tmp, "meta.json")))
        meta["make_case_overrides"] = ov
        meta["params_file"] = params
        if name in REFERENCE_GOLDENS:
            meta["reference_golden"] = REFERENCE_GOLDENS[name]
        if name in RA_GAP_REF:
            meta["Ra_gap"] = RA_GAP_REF[name]
        json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1, sort_keys=True)

        A = ko.load_csr(os.path.join(out, "A.npz"))
        n = A.shape[0]
        rng = np.random.default_rng(20260101)
        store = {}
        if meta["forcing"] == 0:
            B = ko.load_csr(os.path.join(out, "B.npz"))
            tau = complex(meta["rtau"], meta["itau"])
            lam, X, info = ko.eigs(A, B, tau, meta["nev"], meta["which_eigenpairs"])
            store["eig"] = lam
            store["eig_resid"] = ko.residuals(A, B, lam, X)
            store["eig_napply"] = np.array(info["napply"])
            # one shifted solve on a seeded right-hand side in range(B)
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            op = ko.ShiftInvert(A, B, tau)
            store["solve_rhs"] = B @ v
            store["solve_x"] = op.solve(store["solve_rhs"])
        else:
            b = ko.load_csr(os.path.join(out, "B_forced.npz"))
            store["forced_x"] = ko.forced_solve(A, b)
        np.savez_compressed(os.path.join(out, "oracle.npz"), **store)
        if name == "dormy":
            make_rayleigh_slope(name, params, out, RA_GAP_REF[name])
        print(name, "n=%d nnz=%d" % (n, A.nnz), {k: (v.shape if hasattr(v, "shape") else v) for k, v in store.items()})


if __name__ == "__main__":
    main()
