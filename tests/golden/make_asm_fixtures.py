#!/usr/bin/env python3
"""Fixtures of the device-side assembly tests (tests/test_assembly.py).

Needs /root/reference (build container only).  For every case the UNMODIFIED reference stages
run through tools/make_case.py --asm; stored per case: the radial operators submatrices.py wrote
(operators.npz), the physics parameters (asm_params.json) and, for the cases that are not
already golden fixtures, the matrices assemble.py produced (A.npz, B.npz).  The existing cases
(spinover, dormy, jones, forced_small, m0_small) keep the A.npz / B.npz make_golden.py stored:
same reference run, same bits (checked below).
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
from make_golden import CASES, recompress  # noqa: E402

NEW_CASES = {
    # stress-free boundaries on both spheres
    "asm_stressfree": ("tests/spinover/params.spinover", ["bci=0", "bco=0", "N=40"]),
    # thermal, stress-free, flux condition at the inner boundary
    "asm_thermal_flux": ("tests/dormy2004/params.dormy04",
                         ["bci=0", "bco=0", "bci_thermal=1", "N=40", "lmax=40"]),
    # internal heating
    "asm_internal": ("tests/dormy2004/params.dormy04", ["heating='internal'", "N=40", "lmax=40", "bco_thermal=1"]),
    # corners of the parameter space at a small truncation (N = 24 / 48)
    "asm_m0_antisym": ("tests/spinover/params.spinover", ["m=0", "symm=-1", "N=24"]),
    "asm_thermal_m0": ("tests/dormy2004/params.dormy04", ["m=0", "symm=1", "N=24", "lmax=23"]),
    "asm_fullsphere_antisym": ("tests/jones2000/params.jones", ["N=48", "lmax=32", "symm=-1", "Ra_gap=1e6"]),
    "asm_mixed_bc": ("tests/spinover/params.spinover", ["bci=0", "bco=1", "N=24", "m=3", "symm=1"]),
    "asm_forced_thermal": ("tests/dormy2004/params.dormy04",
                           ["forcing=7", "m=2", "symm=1", "N=24", "lmax=25", "forcing_frequency=0.37"]),
    "asm_fullsphere_stressfree": ("tests/jones2000/params.jones",
                                  ["N=48", "lmax=32", "bco=0", "heating='differential'", "Ra_gap=1e6"]),
    # a background temperature gradient of the run's own (radial_profiles.twozone, Vidal & Schaeffer 2015): full sphere
    "asm_twozone": ("tests/jones2000/params.jones", ["heating='two zone'", "N=48", "lmax=16", "m=1", "symm=-1"]),
    # no thermal diffusion: heat equation of first order in the C^(0) basis, no thermal boundary rows
    "asm_no_thermal_diffusion": ("tests/dormy2004/params.dormy04", ["ThermaD=0", "N=24", "lmax=20", "m=3"]),
    # inviscid full sphere (Ek = 0): C^(2) / C^(1) bases for the momentum equations, one no-penetration row, none
    # for the toroidal scalar
    "asm_inviscid": ("tests/spinover/params.spinover", ["Ek=0", "ricb=0", "N=48", "lmax=16", "m=1", "symm=-1"]),
    # double-diffusive: heat and composition equations (APPEND below sets the OmgTau every shipped params file
    # leaves commented out, and a compositional Rayleigh number)
    "asm_compositional": ("tests/dormy2004/params.dormy04",
                          ["compositional=1", "N=24", "lmax=20", "m=3", "bci_compositional=0", "comp_background='differential'"]),
    # the other forcing modes that work in the reference: radial boundary-flow forcing
    "asm_forcing9": ("tests/spinover/params.spinover",
                     ["forcing=9", "m=2", "symm=1", "N=24", "forcing_amplitude_icb=0.7", "forcing_amplitude_cmb=1.3",
                      "forcing_frequency=0.61"]),
    "asm_forcing9_stressfree": ("tests/spinover/params.spinover",
                                ["forcing=9", "m=2", "symm=1", "N=24", "bco=0", "forcing_amplitude_icb=0.7",
                                 "forcing_frequency=0.61"]),
    "asm_forcing10": ("tests/spinover/params.spinover", ["forcing=10", "m=3", "symm=1", "N=24", "forcing_frequency=-0.45"]),
    # magnetic runs (axial and dipole background field, insulating boundaries): rounding-level, not bit-level, parity
    "asm_magnetic_axial": ("tests/spinover/params.spinover", ["magnetic=1", "N=24", "m=2", "symm=1"]),
    "asm_magnetic_dipole_thermal": ("tests/dormy2004/params.dormy04",
                                    ["magnetic=1", "B0='dipole'", "N=24", "lmax=24", "m=3", "forcing=0"]),
    # the other degree-1 background fields: the axial field's block structure, other radial operators
    "asm_magnetic_g21": ("tests/spinover/params.spinover", ["magnetic=1", "B0='G21 dipole'", "N=24", "lmax=17", "m=2", "symm=1"]),
    "asm_magnetic_luo_s1": ("tests/spinover/params.spinover", ["magnetic=1", "B0='Luo_S1'", "N=24", "lmax=18", "m=1", "symm=-1"]),
    "asm_magnetic_fdm": ("tests/spinover/params.spinover", ["magnetic=1", "B0='FDM'", "N=24", "lmax=17", "m=0", "symm=1"]),
    # magnetic full sphere (no inner boundary rows, parity-reduced radial basis), with the heat equation
    "asm_magnetic_fullsphere": ("tests/jones2000/params.jones", ["magnetic=1", "B0='G21 dipole'", "N=48", "lmax=17", "m=2", "symm=-1"]),
    # thin conducting layers on both boundaries (Roberts, Glatzmaier & Clune 2010), dipole field, axisymmetric
    "asm_magnetic_thinwall": ("tests/spinover/params.spinover",
                              ["magnetic=1", "innercore='TWA'", "mantle='TWA'", "B0='dipole'", "N=24", "lmax=17", "m=0", "symm=1",
                               "c_cmb=0.1", "c1_cmb=0.05", "c_icb=0.2", "c1_icb=0.07", "mu=0.8"]),
    # radially varying conductivity (the user's radial_profiles.conductivity; PROFILES below): the eta' terms of the
    # toroidal diffusion, which vanish in every other magnetic fixture
    "asm_magnetic_conductivity": ("tests/dormy2004/params.dormy04",
                                  ["magnetic=1", "B0='axial'", "N=24", "lmax=18", "m=1", "symm=-1", "forcing=0"]),
    # anelastic with a viscosity profile (PROFILES below; the shipped radial_profiles.viscosity is identically zero)
    "asm_anelastic_viscosity": ("tests/dormy2004/params.dormy04",
                                ["anelastic=1", "variable_viscosity=1", "N=24", "lmax=22", "m=3", "bci=0", "bco=1"]),
    # libration-forced magnetic run (dipole field)
    "asm_magnetic_forced": ("tests/spinover/params.spinover",
                            ["magnetic=1", "B0='dipole'", "forcing=7", "m=2", "symm=1", "N=24", "lmax=15",
                             "forcing_amplitude_icb=0.7", "forcing_frequency=0.61"]),
    # quadrupolar background field in a full sphere: radial operators only (tests/test_radial.py); the assembly
    # programs do not cover its l +- 2 couplings, so no matrices are stored
    "ops_magnetic_luo_s2": ("tests/spinover/params.spinover", ["magnetic=1", "B0='Luo_S2'", "N=48", "lmax=18", "m=3", "symm=1", "ricb=0"]),
    # anelastic and magnetic: the density enters the field equations, d ln(rho)/dr the toroidal induction (with a
    # dipole the reference's own viscous terms refer to operators it never generates)
    "asm_anelastic_magnetic": ("tests/dormy2004/params.dormy04",
                               ["anelastic=1", "magnetic=1", "B0='Luo_S1'", "N=24", "lmax=20", "m=3", "forcing=0", "Nrho=2.5"]),
    # anelastic (polytropic background of the params file): wide profile operators; bit for bit again
    "asm_anelastic": ("tests/dormy2004/params.dormy04", ["anelastic=1", "N=24", "lmax=24", "m=3"]),
    "asm_anelastic_stressfree": ("tests/dormy2004/params.dormy04",
                                 ["anelastic=1", "N=24", "lmax=24", "m=3", "bci=0", "bco=0", "bco_thermal=1"]),
    "asm_anelastic_hydro": ("tests/dormy2004/params.dormy04", ["anelastic=1", "thermal=0", "N=24", "lmax=23", "m=0", "symm=-1"]),
}
PROFILES = {"asm_magnetic_conductivity": "def conductivity(r):\n    return 1 + 0.5*r**2",
            "asm_anelastic_viscosity": "def viscosity(r):\n    return 1 + 0.3*r**2"}
APPEND = {"asm_compositional": "OmgTau = 1\nSchmidt = 0.3\nBV2_comp = -3.7e6 * Ek**2 / Schmidt"}
EXISTING = ["spinover", "dormy", "jones", "forced_small", "m0_small", "magnetic_small"]


def same_npz(a, b):
    za, zb = np.load(a), np.load(b)
    return all(np.array_equal(za[k], zb[k]) for k in ("data", "indices", "indptr"))


def main():
    only = sys.argv[1:]
    todo = {k: CASES[k] for k in EXISTING}
    todo.update(NEW_CASES)
    for name, (params, ov) in todo.items():
        if only and name not in only:
            continue
        out = os.path.join(HERE, name)
        tmp = "/tmp/asmfix_" + name
        shutil.rmtree(tmp, ignore_errors=True)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_case.py"),
                               "--params", params, "--out", tmp, "--asm"]
                              + (["--profiles", PROFILES[name]] if name in PROFILES else [])
                              + (["--append-params", APPEND[name]] if name in APPEND else []) + ov)
        os.makedirs(out, exist_ok=True)
        for fn in ("operators.npz", "asm_params.json"):
            shutil.copy(os.path.join(tmp, fn), os.path.join(out, fn))
        if os.path.exists(os.path.join(tmp, "radprofs.npz")):
            shutil.copy(os.path.join(tmp, "radprofs.npz"), os.path.join(out, "radprofs.npz"))
        for fn in ("A.npz", "B.npz", "B_forced.npz"):
            src = os.path.join(tmp, fn)
            if not os.path.exists(src) or name.startswith("ops_"):
                continue
            if name in NEW_CASES:
                recompress(src, os.path.join(out, fn))
            else:
                assert same_npz(src, os.path.join(out, fn)), (name, fn)
        if name in NEW_CASES:
            meta = json.load(open(os.path.join(tmp, "meta.json")))
            meta["make_case_overrides"] = ov
            meta["params_file"] = params
            json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1, sort_keys=True)
        # ||B||_F as THIS machine's BLAS computed it inside the reference run (a threaded dot product: its
        # last bit depends on the core count): found as the norm that makes the model reproduce B.npz
        pj = json.load(open(os.path.join(out, "asm_params.json")))
        if os.path.exists(os.path.join(out, "B.npz")) and not name.startswith("ops_"):
            sys.path.insert(0, ROOT)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import assembly_model as am
            from kore_b200 import assembly as asm
            pp = asm.PhysicsParams.from_dict(pj)
            ops = asm.load_operators_npz(os.path.join(out, "operators.npz"))
            pB = asm.build_program_B(pp, ops)
            bnorm = asm.frobenius_norm(am.evaluate(pB).data)
            z = np.load(os.path.join(out, "B.npz"))
            assert np.array_equal(am.evaluate(pB.with_final_scale(1. / bnorm)).data, z["data"]), name
            pj["Bnorm"] = bnorm
            json.dump(pj, open(os.path.join(out, "asm_params.json"), "w"), indent=1, sort_keys=True)
        print(name, "ok")


if __name__ == "__main__":
    main()
