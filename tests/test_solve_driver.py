"""The bin/solve.py twin (kore_b200/solve.py): option parsing and file contract on CPU,
the full run through the GPU library under -m gpu."""
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, load_case


def write_run_dir(tmp_path, name):
    c = load_case(name)
    m = c.meta
    d = tmp_path / name
    (d / "bin").mkdir(parents=True)
    for fn in ("A.npz", "B.npz", "B_forced.npz"):
        src = os.path.join(GOLDEN, name, fn)
        if os.path.exists(src):
            shutil.copy(src, d / fn)
    keys = ["hydro", "magnetic", "thermal", "compositional", "m", "lmax", "N", "symm", "ricb", "forcing",
            "nev", "maxit", "tol", "rtau", "itau"]
    lines = ["%s = %r" % (k, m[k]) for k in keys]
    lines += ["which_eigenpairs = %r" % m["which_eigenpairs"], "B0 = %r" % m["B0"], "tau = rtau + itau*1j"]
    (d / "bin" / "parameters.py").write_text("\n".join(lines) + "\n")
    return c, d


def test_options_parser():
    from kore_b200.eps import Options
    o = Options("-st_type sinvert -st_pc_factor_mat_solver_type mumps -mat_mumps_icntl_14 1000 "
                "-eps_true_residual -eps_balance twoside -eps_error_relative ::ascii_info_detail "
                "-eps_nev 7 -eps_tol 1e-10 -eps_target 0.1+1.5i -nbl 12".split())
    assert o.getString("st_type") == "sinvert"
    assert o.getInt("mat_mumps_icntl_14") == 1000
    assert o.hasName("eps_true_residual") and o.hasName("eps_error_relative")
    assert o.getInt("eps_nev") == 7 and o.getReal("eps_tol") == 1e-10
    assert o.getScalar("eps_target") == 0.1 + 1.5j
    assert o.getInt("nbl") == 12 and o.getInt("missing", 3) == 3


@pytest.mark.parametrize("name", ["spinover", "dormy", "jones", "magnetic_small"])
def test_sizes_and_field_slices_follow_reference(name):
    import types
    from kore_b200 import solve as drv
    c = load_case(name)
    par = types.SimpleNamespace(**c.meta)
    N1, n, sizmat, symmB0 = drv.kore_sizes(par)
    assert (N1, n, sizmat, symmB0) == (c.meta["N1"], c.meta["n"], c.meta["sizmat"], c.meta["symmB0"])
    sl = drv.field_slices(par, n)
    assert sl[0] == ("flow", 0, 2 * n)
    assert sl[-1][2] == sizmat


@pytest.mark.gpu
def test_eigen_run_writes_reference_files(tmp_path, monkeypatch, lib):
    import sys
    c, d = write_run_dir(tmp_path, "spinover")
    monkeypatch.chdir(d)
    sys.modules.pop("parameters", None)
    from kore_b200 import solve as drv
    assert drv.main(["-st_type", "sinvert", "-eps_error_relative", "::ascii_info_detail"]) == 0
    sys.modules.pop("parameters", None)
    eig = np.loadtxt("eigenvalues0.dat").reshape(-1, 2)
    # tests/test_spinover.py:21-29
    best = eig[np.argmax(eig[:, 0])]
    np.testing.assert_allclose(best, c.meta["reference_golden"]["eig"], rtol=1e-8, atol=1e-20)
    n = c.meta["n"]
    ru, iu = np.loadtxt("real_flow.field"), np.loadtxt("imag_flow.field")
    assert ru.reshape(2 * n, -1).shape == (2 * n, eig.shape[0]) and iu.shape == ru.shape
    assert not os.path.exists("real_magnetic.field") and not os.path.exists("no_conv_solution")
    assert np.loadtxt("timing.dat").size == 1
    # the written vectors are eigenvectors of the pencil
    x = (ru + 1j * iu).reshape(2 * n, -1)[:, 0]
    lam = eig[0, 0] + 1j * eig[0, 1]
    bx = c.B @ x
    assert np.linalg.norm(c.A @ x - lam * bx) <= 1e-10 * abs(lam) * np.linalg.norm(bx)


@pytest.mark.gpu
def test_forced_run_writes_single_column(tmp_path, monkeypatch, lib):
    import sys
    c, d = write_run_dir(tmp_path, "forced_small")
    monkeypatch.chdir(d)
    sys.modules.pop("parameters", None)
    from kore_b200 import solve as drv
    assert drv.main(["-ksp_type", "preonly", "-pc_type", "lu"]) == 0
    sys.modules.pop("parameters", None)
    n = c.meta["n"]
    x = np.loadtxt("real_flow.field") + 1j * np.loadtxt("imag_flow.field")
    assert x.shape == (2 * n,)
    xo = c.oracle["forced_x"]
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
    assert not os.path.exists("eigenvalues0.dat")


def write_operator_dir(tmp_path, name):
    """A run directory as bin/submatrices.py leaves it: parameters.py and the *.mtx radial operators,
    no A.npz / B.npz (those would come from bin/assemble.py)."""
    import json
    import scipy.io as sio
    from kore_b200 import assembly as asm
    c = load_case(name)
    d = tmp_path / (name + "_ops")
    (d / "bin").mkdir(parents=True)
    pj = json.load(open(os.path.join(GOLDEN, name, "asm_params.json")))
    m = c.meta
    lines = ["%s = %r" % (k, v) for k, v in pj.items() if k != "rcmb"]
    lines += ["%s = %r" % (k, m[k]) for k in ("nev", "maxit", "tol", "rtau", "itau")]
    lines += ["which_eigenpairs = %r" % m["which_eigenpairs"], "B0 = %r" % m["B0"], "tau = rtau + itau*1j"]
    (d / "bin" / "parameters.py").write_text("\n".join(lines) + "\n")
    for lab, M in asm.load_operators_npz(os.path.join(GOLDEN, name, "operators.npz")).items():
        sio.mmwrite(str(d / (lab + ".mtx")), M)
    if os.path.exists(os.path.join(GOLDEN, name, "B_forced.npz")):
        shutil.copy(os.path.join(GOLDEN, name, "B_forced.npz"), d / "B_forced.npz")
    return c, d


def test_operator_dir_round_trips_the_operators(tmp_path):
    # MatrixMarket text keeps every digit: what the driver reads back is what the fixture holds
    from kore_b200 import assembly as asm
    c, d = write_operator_dir(tmp_path, "m0_small")
    ops = asm.load_operators(str(d))
    ref = asm.load_operators_npz(os.path.join(GOLDEN, "m0_small", "operators.npz"))
    assert sorted(ops) == sorted(ref)
    for k in ref:
        assert (ops[k] != ref[k]).nnz == 0


@pytest.mark.gpu
def test_eigen_run_from_radial_operators_equals_run_from_matrices(tmp_path, monkeypatch, lib):
    # device-side assembly produces the reference's matrices (up to the last bit of ||B||_F, which is this
    # host's BLAS): the two runs write the same eigenvalues and fields
    import sys
    from kore_b200 import solve as drv
    c, d1 = write_run_dir(tmp_path, "spinover")
    monkeypatch.chdir(d1)
    sys.modules.pop("parameters", None)
    assert drv.main(["-st_type", "sinvert"]) == 0
    sys.modules.pop("parameters", None)
    c, d2 = write_operator_dir(tmp_path, "spinover")
    monkeypatch.chdir(d2)
    assert not os.path.exists("A.npz")
    assert drv.main(["-st_type", "sinvert", "-kb_diagnose", "-kb_npz"]) == 0
    z = np.load("eigenpairs.npz")
    assert z["vectors"].shape == (c.n, len(z["eigenvalues"])) and z["fields"].tolist() == [["flow", "0", str(c.n)]]
    ev = np.loadtxt("eigenvalues0.dat").reshape(-1, 2)
    assert np.array_equal(z["eigenvalues"], ev[:, 0] + 1j * ev[:, 1])
    assert np.array_equal(z["vectors"][:, 0].real, np.loadtxt("real_flow.field").reshape(c.n, -1)[:, 0])
    sys.modules.pop("parameters", None)
    # -kb_diagnose: one row of energy / dissipation integrals and power-balance residuals per solution
    pb = np.loadtxt("power_balance.dat").reshape(-1, 8)
    eig = np.loadtxt("eigenvalues0.dat").reshape(-1, 2)
    assert pb.shape[0] == eig.shape[0]
    assert pb[np.argmax(eig[:, 0]), 7] < 1e-4 and np.all(pb[:, 0] > 0)
    e1, e2 = np.loadtxt(d1 / "eigenvalues0.dat"), np.loadtxt(d2 / "eigenvalues0.dat")
    assert e1.shape == e2.shape and np.max(np.abs(e1 - e2)) <= 1e-11
    for fn in ("real_flow.field", "imag_flow.field"):
        f1, f2 = np.loadtxt(d1 / fn), np.loadtxt(d2 / fn)
        assert f1.shape == f2.shape and np.max(np.abs(f1 - f2)) <= 1e-8 * np.max(np.abs(f1)), fn


@pytest.mark.gpu
def test_forced_run_from_radial_operators(tmp_path, monkeypatch, lib):
    import sys
    from kore_b200 import solve as drv
    c, d = write_operator_dir(tmp_path, "forced_small")
    monkeypatch.chdir(d)
    os.remove("B_forced.npz")  # the libration forcing vector is formed by the driver (assemble.py:278-329)
    sys.modules.pop("parameters", None)
    assert drv.main(["-kb_assemble"]) == 0
    sys.modules.pop("parameters", None)
    x = np.loadtxt("real_flow.field") + 1j * np.loadtxt("imag_flow.field")
    xo = c.oracle["forced_x"]
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
