"""Radial operators (kore_b200/radial.py) against the operators the UNMODIFIED
reference wrote (bin/submatrices.py through tools/make_case.py --asm; tests/golden/*/operators.npz).

The reference's band entries are numpy.dot products, so their last bit belongs to the BLAS of the
machine the fixtures were made on (the build container).  The bar: every entry within 16 ulp of its row's
scale (a band entry's rounding passes through up to four basis changes), the
sparsity pattern identical; and bit for bit when this machine's BLAS sums as the container's does
(which the test finds out on one fixture and reports in its skip message otherwise).
"""
import glob
import json
import os
import types

import numpy as np
import pytest

from kore_b200 import assembly as asm
from kore_b200 import radial

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(HERE, "golden", "*", "operators.npz")))
LARGE = {"asm_E1e-7", "asm_E1e-8", "asm_E1e-8_kore_rule"}


def load_case(name):
    d = os.path.join(HERE, "golden", name)
    pp = asm.PhysicsParams.from_dict(json.load(open(os.path.join(d, "asm_params.json"))))
    ref = asm.load_operators_npz(os.path.join(d, "operators.npz"))
    rp = os.path.join(d, "radprofs.npz")
    return pp, ref, (dict(np.load(rp)) if os.path.exists(rp) else None)


def ulp_distance(a, b):
    """largest difference in units of the last place of the largest entry of its row (a band entry
    is a short sum of products that may cancel: its rounding error scales with the terms, not with
    the sum)"""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    scale = np.spacing(np.maximum(np.abs(a).max(axis=1), np.abs(b).max(axis=1)))[:, None]
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - b) / scale))


def test_fixtures_cover_the_parameter_space():
    assert len(CASES) >= 41
    kinds = set()
    for name in CASES:
        pp, _, _ = load_case(name)
        kinds.add((pp.ricb > 0, bool(pp.thermal), bool(pp.magnetic), bool(pp.anelastic)))
    # shell / full sphere, hydro / thermal, magnetic (axial and dipole), anelastic
    assert {(True, False, False, False), (False, True, False, False), (True, True, False, False),
            (True, False, True, False), (True, True, True, False), (True, True, False, True)} <= kinds


@pytest.mark.parametrize("name", CASES)
def test_operators_equal_the_reference(name):
    pp, ref, radprofs = load_case(name)
    ops = radial.radial_operators(pp, radprofs=radprofs)
    assert sorted(ops) == sorted(ref)                     # the same ``.mtx`` files, no more, no fewer
    exact = True
    for lab in ref:
        a, b = ref[lab].toarray(), ops[lab].toarray()
        assert a.shape == b.shape == (pp.N1, pp.N1), lab
        assert np.array_equal(a != 0, b != 0), lab
        if not np.array_equal(a, b):
            exact = False
            assert ulp_distance(a, b) <= 16.0, lab
    if not exact:
        pytest.skip("within 16 ulp of the row scale; not bit for bit: this machine's BLAS sums numpy.dot in another order than the "
                    "machine the fixtures were made on")


@pytest.mark.parametrize("name", ["spinover", "jones", "asm_magnetic_dipole_thermal", "asm_anelastic", "asm_E1e-7"])
def test_ordered_summation_is_rounding_away(name):
    """dot='ordered' (machine independent) differs from the BLAS dot products by rounding only"""
    pp, ref, radprofs = load_case(name)
    ops = radial.radial_operators(pp, radprofs=radprofs, dot="ordered", dense=True)
    for lab in ref:
        a = ref[lab].toarray()
        assert np.array_equal(a != 0, ops[lab] != 0), lab
        assert ulp_distance(a, ops[lab]) <= 16.0, lab


def test_generated_operators_assemble_the_reference_pencil():
    """parameters -> operators -> assembly program -> A, B: the reference's matrices from its
    parameter file alone (the NumPy model of the device kernel stands in for the GPU here)"""
    import assembly_model as am
    for name in ("spinover", "asm_thermal_flux", "asm_fullsphere_stressfree"):
        d = os.path.join(HERE, "golden", name)
        pj = json.load(open(os.path.join(d, "asm_params.json")))
        pp = asm.PhysicsParams.from_dict(pj)
        ops = radial.radial_operators(pp)
        B = am.evaluate(asm.build_program_B(pp, ops).with_final_scale(1. / pj["Bnorm"]))
        A = am.evaluate(asm.build_program_A(pp, ops).with_final_scale(1. / pj["Bnorm"]))
        import scipy.sparse as sp
        for M, fn in ((A, "A.npz"), (B, "B.npz")):
            z = np.load(os.path.join(d, fn))
            if np.array_equal(M.indptr, z["indptr"]) and np.array_equal(M.indices, z["indices"]) and np.array_equal(M.data, z["data"]):
                continue                                      # bit for bit (the machine the fixtures were made on)
            R = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=M.shape)
            assert abs(M - R).max() <= 1e-13 * np.abs(z["data"]).max(), (name, fn)


def test_multiplication_matrix_multiplies():
    """the multiplication matrix of r^p in the C^(lamb) basis does multiply: applied to the
    C^(lamb) coefficients of a polynomial q it gives those of r^p q (a property, no fixture)"""
    N, ricb, rcmb = 40, 0.35, 1.0
    bases = radial.GegenbauerBases(N)
    r = radial._nodes(N, ricb, rcmb)
    rng = np.random.default_rng(5)
    q = np.zeros(N)
    q[:12] = rng.standard_normal(12)                      # Chebyshev coefficients of q, degree 11
    import numpy.polynomial.chebyshev as ch
    x = np.cos(np.pi * (np.arange(N) + 0.5) / N)
    for p in (1, 3):
        prod = radial._dct_coefficients(r ** p * ch.chebval(x, q), N, 0.0)
        for lamb in (0, 1, 2, 4):
            rp = radial.chebco(p, N, radial.TOL, ricb, rcmb)
            to = (lambda v: v) if lamb == 0 else bases.S[lamb].apply
            M = radial.multiplication(to(rp), lamb, 0)
            assert np.allclose(M @ to(q), to(prod), rtol=0, atol=1e-12 * np.abs(to(prod)).max()), (p, lamb)


def test_derivative_and_basis_change():
    """D^lamb of a Chebyshev series, brought back to values on the nodes, is the derivative"""
    import numpy.polynomial.chebyshev as ch
    N, ricb, rcmb = 24, 0.35, 1.0
    rng = np.random.default_rng(7)
    q = np.zeros(N)
    q[:8] = rng.standard_normal(8)
    dq = ch.chebder(q) * 2 / (rcmb - ricb)                # d/dr
    dq = np.concatenate([dq, np.zeros(N - dq.size)])
    S0 = radial.basis_change(0, N)
    D1 = np.zeros((N, N))
    d = radial.derivative_diagonal(1, N, ricb, rcmb)
    D1[np.arange(N - 1), np.arange(1, N)] = d
    assert np.allclose(D1 @ q, S0 @ dq, atol=1e-13)


def test_profile_tables_and_uniform_conductivity():
    """compute_profiles.py's table of a uniform conductivity is what the magnetic fixtures hold"""
    pp, _, radprofs = load_case("asm_magnetic_axial")
    rap = types.SimpleNamespace(magnetic_diffusivity=lambda r: 1. / np.ones_like(r))
    t = radial.profile_tables(pp, rap)
    assert list(t) == ["cd_eta"] and np.array_equal(t["cd_eta"], radprofs["cd_eta"])
    # a run with its own conductivity profile (tools/make_case.py --profiles; radial_profiles.py:252-262)
    pp, _, radprofs = load_case("asm_magnetic_conductivity")
    rap = types.SimpleNamespace(magnetic_diffusivity=lambda r: 1. / (1 + 0.5 * r ** 2))
    assert np.array_equal(radial.profile_tables(pp, rap)["cd_eta"], radprofs["cd_eta"])
    assert np.count_nonzero(radprofs["cd_eta"][:, 1]) > 5
    # the run's own background temperature gradient (heating = 'two zone'; radial_profiles.py:6-23 is the user's file)
    pp, _, radprofs = load_case("asm_twozone")

    def twozone(r, args):
        rc, h, sym = args
        out = np.zeros_like(r)
        for i, x in enumerate(r):
            out[i] = (1 if x >= 0 else sym) * (1 + np.tanh(2 * (abs(x) - rc) / h)) / 2
        return out
    pp.args = [0.7, 0.1, -1]                              # parameters.py:203-206
    t = radial.profile_tables(pp, types.SimpleNamespace(twozone=twozone))
    assert list(t) == ["cd_ent"] and t["cd_ent"].shape == radprofs["cd_ent"].shape
    assert np.max(np.abs(t["cd_ent"] - radprofs["cd_ent"])) <= 1e-15
    # a derivative column differentiates: d/dr of r^3 = 3 r^2
    tab = radial.profile_table(lambda r: r ** 3, 2, 24, 0.35, 1.0)
    assert np.allclose(tab[:, 1], 3 * radial.chebco(2, 24, 0.0, 0.35, 1.0), atol=1e-9)
    assert np.allclose(tab[:, 2], 6 * radial.chebco(1, 24, 0.0, 0.35, 1.0), atol=1e-9)


def test_unsupported_runs_are_refused():
    pp, _, _ = load_case("asm_anelastic")
    with pytest.raises(ValueError):
        radial.radial_operators(pp)                       # anelastic without the run's profiles
    pp, _, _ = load_case("dormy")
    pp.heating = "something else"
    with pytest.raises(NotImplementedError):
        radial.radial_operators(pp)
    pp, _, _ = load_case("asm_magnetic_axial")
    pp.B0, pp.B0_l = "FDM", 2
    with pytest.raises(NotImplementedError):
        radial.radial_operators(pp)


def test_mtx_files_are_read_back_by_the_assembly(tmp_path):
    pp, ref, _ = load_case("asm_mixed_bc")
    radial.write_mtx(str(tmp_path), radial.radial_operators(pp))
    back = asm.load_operators(str(tmp_path))
    assert sorted(back) == sorted(ref)
    for lab in ref:
        assert np.array_equal(back[lab].toarray(), ref[lab].toarray())


def test_command_line_twin_of_submatrices(tmp_path):
    """``python -m kore_b200.radial ncpus`` in a run directory writes the reference's ``.mtx`` files"""
    import subprocess
    import sys
    d = os.path.join(HERE, "golden", "dormy")
    pj = json.load(open(os.path.join(d, "asm_params.json")))
    os.makedirs(tmp_path / "bin")
    with open(tmp_path / "bin" / "parameters.py", "w") as f:
        for k, v in pj.items():
            f.write("%s = %r\n" % (k, v))
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "kore_b200.radial", "8"], cwd=tmp_path, env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "Submatrices generated and written to disk" in out.stdout
    ref = asm.load_operators_npz(os.path.join(d, "operators.npz"))
    back = asm.load_operators(str(tmp_path))
    assert sorted(back) == sorted(ref)
    for lab in ref:
        assert ulp_distance(back[lab].toarray(), ref[lab].toarray()) <= 16.0
