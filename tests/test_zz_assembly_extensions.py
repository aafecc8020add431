"""Device-side assembly of MAGNETIC pencils (axial or dipole background field, insulating inner core and
mantle; BASELINE.json config 4): kore_b200/assembly.py `_magnetic_blocks`.

Parity here is to ROUNDING, not to the bit as for the hydrodynamic and thermal blocks (tests/test_assembly.py):
the reference evaluates its induction coefficients in numpy.float128 (operators.py:479) and nests the sums of
the Lorentz terms, the assembly program is one left-to-right double sum per block.  Bars: B bit for bit; every
block of A within 1e-13 of its own largest entry (observed 4e-16; the hydrodynamic / thermal blocks of the same
matrices stay bit-identical); eigenvalues of the assembled pencil within 1e-9 of the oracle's on the
reference-assembled one.

The GPU tests of this file were written after the round's GPU budget was spent: the kernels they run are the
ones tests/test_assembly.py validates (the assembly kernel evaluates any program; these programs have more
groups and longer sums, nothing new), but they have not themselves run on a GPU yet -- hence this file sorts
last."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import assembly_model as am
from conftest import GOLDEN, load_case
from kore_b200 import assembly as asm

CASES = ["magnetic_small", "asm_magnetic_axial", "asm_magnetic_dipole_thermal",
         # the other degree-1 fields (Gerick's dipole, Luo & Jackson's S1, a free-decay mode; m = 2, 1, 0): the axial
         # field's programs on other radial operators
         "asm_magnetic_g21", "asm_magnetic_luo_s1", "asm_magnetic_fdm",
         # full sphere (G21 dipole, internal heating, m = 2 antisymmetric): outer boundary rows only, parity-reduced basis
         "asm_magnetic_fullsphere",
         # thin conducting layers at both boundaries (innercore = mantle = 'TWA'), dipole field, m = 0
         "asm_magnetic_thinwall",
         # density-stratified AND magnetic (Luo_S1 field, heat equation): rho in the field equations, d ln(rho)/dr
         # in the toroidal induction (operators.py:445-462, 571-629, 660-689)
         "asm_anelastic_magnetic",
         # radially varying conductivity (axial field, heat equation): non-zero eta' terms in the toroidal diffusion
         "asm_magnetic_conductivity"]


def fixture(name):
    d = os.path.join(GOLDEN, name)
    pj = json.load(open(os.path.join(d, "asm_params.json")))
    pp = asm.PhysicsParams.from_dict(pj)
    ops = asm.load_operators_npz(os.path.join(d, "operators.npz"))

    def csr(fn):
        z = np.load(os.path.join(d, fn))
        M = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        M.sort_indices()
        return M
    return pj, pp, ops, csr("A.npz"), csr("B.npz")


def block_relative_error(A, A_ref, N1):
    """max over the N1 x N1 blocks of |A - A_ref| / (largest |A_ref| of the block); inf for a block the reference
    does not have."""
    D, R = (A - A_ref).tocoo(), A_ref.tocoo()
    nbr = A.shape[0] // N1
    mx = np.zeros((nbr, nbr))
    np.maximum.at(mx, (R.row // N1, R.col // N1), np.abs(R.data))
    er = np.zeros((nbr, nbr))
    np.maximum.at(er, (D.row // N1, D.col // N1), np.abs(D.data))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(mx > 0, er / np.where(mx > 0, mx, 1.0), np.where(er > 0, np.inf, 0.0))


def model_pencil(pj, pp, ops):
    s = 1. / pj["Bnorm"]
    return (am.evaluate(asm.build_program_A(pp, ops).with_final_scale(s)),
            am.evaluate(asm.build_program_B(pp, ops).with_final_scale(s)))


@pytest.mark.parametrize("name", CASES)
def test_magnetic_program_matches_reference_to_rounding(name):
    pj, pp, ops, A_ref, B_ref = fixture(name)
    A, B = model_pencil(pj, pp, ops)
    assert np.array_equal(B.indptr, B_ref.indptr) and np.array_equal(B.indices, B_ref.indices)
    assert np.array_equal(B.data, B_ref.data)
    rel = block_relative_error(A, A_ref, pp.N1)
    assert rel.max() <= 1e-13, rel.max()
    # only blocks that involve the field differ at all; the momentum / heat blocks are the bit-exact ones
    nf = 2 * pp.nb  # block rows (= block columns) of u and v
    hydro = np.zeros_like(rel, dtype=bool)
    hydro[:nf, :nf] = True
    if pp.thermal:
        hydro[4 * pp.nb:, :nf] = hydro[4 * pp.nb:, 4 * pp.nb:] = hydro[:nf, 4 * pp.nb:] = True
    assert rel[hydro].max() == 0.0
    # entries present in one pattern only are cancellations to (almost) zero
    assert abs(A.nnz - A_ref.nnz) <= 1e-3 * A_ref.nnz


def test_magnetic_pencil_has_the_oracles_eigenvalues():
    # magnetic_small (dipole, N = 40): the oracle on the model-assembled pencil against the fixture's eigenvalues
    # (the oracle on the reference-assembled pencil)
    import kore_oracle as ko
    pj, pp, ops, A_ref, B_ref = fixture("magnetic_small")
    case = load_case("magnetic_small")
    A, B = model_pencil(pj, pp, ops)
    lam, _, _ = ko.eigs(A, B, case.tau, case.meta["nev"], case.meta["which_eigenpairs"])
    for z in case.oracle["eig"]:
        assert np.min(np.abs(lam - z)) <= 1e-10 * abs(z)


def forced_fixture(name="asm_magnetic_forced"):
    d = os.path.join(GOLDEN, name)
    pj = json.load(open(os.path.join(d, "asm_params.json")))
    pp = asm.PhysicsParams.from_dict(pj)
    ops = asm.load_operators_npz(os.path.join(d, "operators.npz"))
    z = np.load(os.path.join(d, "A.npz"))
    A_ref = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    A_ref.sort_indices()
    z = np.load(os.path.join(d, "B_forced.npz"))
    rhs = np.asarray(sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"])).todense()).ravel()
    return pp, ops, A_ref, rhs


def test_forced_magnetic_program():
    # libration (forcing = 7) of a magnetic run: A un-normalised with the forcing frequency in every time derivative,
    # the field's included; the forcing vector is the hydrodynamic one (assemble.py:278-329)
    pp, ops, A_ref, rhs = forced_fixture()
    assert pp.magnetic == 1 and pp.forcing == 7
    A = am.evaluate(asm.build_program_A(pp, ops))
    assert block_relative_error(A, A_ref, pp.N1).max() <= 1e-13
    assert abs(A.nnz - A_ref.nnz) <= 1e-3 * A_ref.nnz
    assert np.array_equal(asm.forcing_vector(pp), rhs)


def test_other_magnetic_setups_are_refused():
    pj, pp, ops, _, _ = fixture("asm_magnetic_axial")
    for kw in (dict(B0="Luo_S2"), dict(B0="FDM", B0_l=2), dict(innercore="perfect conductor, material"), dict(mantle="conducting"), dict(ricb=0.0, B0="dipole"), dict(forcing=8)):
        q = asm.PhysicsParams.from_dict({**pp.__dict__, **kw})
        with pytest.raises(NotImplementedError):
            asm.build_program_A(q, ops)


# ---------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_assembles_the_magnetic_program_like_the_model(lib, name):
    pj, pp, ops, A_ref, B_ref = fixture(name)
    A_m, B_m = model_pencil(pj, pp, ops)
    with lib.Solver(0) as s:
        asm.assemble(s, pp, ops, bnorm=pj["Bnorm"])
        ip, ix, v = s.get_assembled("A")
        A = sp.csr_matrix((v, ix, ip), shape=A_m.shape)
        ip, ix, v = s.get_assembled("B")
        B = sp.csr_matrix((v, ix, ip), shape=B_m.shape)
    # the kernel and its NumPy model perform the same operations: same bits
    assert np.array_equal(A.indptr, A_m.indptr) and np.array_equal(A.indices, A_m.indices) and np.array_equal(A.data, A_m.data)
    assert np.array_equal(B.data, B_ref.data)
    assert block_relative_error(A, A_ref, pp.N1).max() <= 1e-13


@pytest.mark.gpu
def test_device_forced_magnetic_solve_against_superlu(lib):
    # the forced magnetic pencil assembled on the device and solved there, against SuperLU on the reference's A
    import scipy.sparse.linalg as spl
    from kore_b200 import chain
    pp, ops, A_ref, rhs = forced_fixture()
    x_ref = spl.splu(A_ref.tocsc()).solve(rhs)
    perm, nodeptr = chain.chain_from_params(pp.N1, pp.m, pp.lmax, pp.symm, -1, pp.hydro, pp.magnetic, pp.thermal, 0)
    with lib.Solver(0) as s:
        asm.assemble(s, pp, ops)
        s.set_chain(perm, nodeptr)
        s.factor(0.0)
        x = s.solve(rhs)
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)


@pytest.mark.gpu
def test_device_assembled_magnetic_pencil_against_the_oracle(lib):
    case = load_case("magnetic_small")
    pj, pp, ops, _, _ = fixture("magnetic_small")
    m = case.meta
    with lib.Solver(0) as s:
        asm.assemble(s, pp, ops, bnorm=pj["Bnorm"])
        s.set_chain(case.perm, case.nodeptr)
        s.factor(case.tau)
        lam, X, info = s.eigs(m["nev"], which=m["which_eigenpairs"], target=case.tau, tol=m["tol"], maxit=m["maxit"])
    assert info["nconv"] >= m["nev"]
    for z in case.oracle["eig"]:
        assert np.min(np.abs(lam - z)) <= 1e-9 * abs(z), (z, lam)
    assert np.all(info["resid"] <= 1e-10)


# ------------------------------------------------------------------------------------ anelastic runs
ANELASTIC = ["asm_anelastic", "asm_anelastic_stressfree", "asm_anelastic_hydro"]


@pytest.mark.parametrize("name", ANELASTIC)
def test_anelastic_program_reproduces_reference_assembly_bitwise(name):
    # density-stratified (anelastic) runs: viscous force with the log-density profile operators, buoyancy and entropy
    # equation with the background profiles, stress-free rows with d ln(rho)/dr (operators.py:146-152, 176-179, 393-395,
    # 709-711, 731-732, 765-768; assemble.py:1195-1198, 1244, 1262, 1284, 1291).  Plain double arithmetic in the
    # reference, so the bar is the bit again.  The profile operators are wide (+-17 at N = 24).
    pj, pp, ops, A_ref, B_ref = fixture(name)
    assert pp.anelastic == 1
    A, B = model_pencil(pj, pp, ops)
    assert asm.build_program_A(pp, ops).H > 15
    for M, R in ((A, A_ref), (B, B_ref)):
        assert np.array_equal(M.indptr, R.indptr) and np.array_equal(M.indices, R.indices) and np.array_equal(M.data, R.data)


def test_variable_viscosity_program_matches_reference_to_rounding():
    # anelastic with a viscosity profile nu(r) (operators.py:154-168, 181-185): 26 + 9 operators with nu, nu', nu''
    # and the log-density derivatives.  The reference nests two sums of the poloidal block, the program distributes
    # them: the two viscous blocks agree to rounding, everything else (and B) to the bit
    pj, pp, ops, A_ref, B_ref = fixture("asm_anelastic_viscosity")
    assert pp.anelastic == 1 and pp.variable_viscosity == 1 and pp.bci == 0
    A, B = model_pencil(pj, pp, ops)
    assert np.array_equal(B.indptr, B_ref.indptr) and np.array_equal(B.indices, B_ref.indices) and np.array_equal(B.data, B_ref.data)
    rel = block_relative_error(A, A_ref, pp.N1)
    assert rel.max() <= 1e-13
    off = ~np.eye(rel.shape[0], dtype=bool)
    assert rel[off].max() == 0.0 and rel[2 * pp.nb:, 2 * pp.nb:].max() == 0.0
    assert A.nnz == A_ref.nnz


def test_anelastic_stressfree_needs_the_density_slopes():
    pj, pp, ops, _, _ = fixture("asm_anelastic_stressfree")
    assert pp.lho1_icb is not None and pp.lho1_cmb is not None
    q = asm.PhysicsParams.from_dict({k: v for k, v in pp.__dict__.items() if k not in ("lho1_icb", "lho1_cmb")})
    with pytest.raises(NotImplementedError):
        asm.build_program_A(q, ops)
    with pytest.raises(KeyError):  # a viscosity profile needs its own radial operators
        asm.build_program_A(asm.PhysicsParams.from_dict({**pp.__dict__, "variable_viscosity": 1}), ops)


@pytest.mark.gpu
def test_device_assembles_the_variable_viscosity_program_like_the_model(lib):
    pj, pp, ops, A_ref, B_ref = fixture("asm_anelastic_viscosity")
    A_m, B_m = model_pencil(pj, pp, ops)
    with lib.Solver(0) as s:
        asm.assemble(s, pp, ops, bnorm=pj["Bnorm"])
        ip, ix, v = s.get_assembled("A")
        jp, jx, w = s.get_assembled("B")
    assert np.array_equal(ip, A_m.indptr) and np.array_equal(ix, A_m.indices) and np.array_equal(v, A_m.data)
    assert np.array_equal(jp, B_ref.indptr) and np.array_equal(jx, B_ref.indices) and np.array_equal(w, B_ref.data)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ANELASTIC)
def test_device_assembles_the_anelastic_program_bitwise(lib, name):
    pj, pp, ops, A_ref, B_ref = fixture(name)
    with lib.Solver(0) as s:
        asm.assemble(s, pp, ops, bnorm=pj["Bnorm"])
        ip, ix, v = s.get_assembled("A")
        jp, jx, w = s.get_assembled("B")
    assert np.array_equal(ip, A_ref.indptr) and np.array_equal(ix, A_ref.indices) and np.array_equal(v, A_ref.data)
    assert np.array_equal(jp, B_ref.indptr) and np.array_equal(jx, B_ref.indices) and np.array_equal(w, B_ref.data)


# ------------------------------------------------------------- from the parameter file alone (radial.py)
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["spinover", "dormy"])
def test_pencil_from_the_parameters_alone_against_the_oracle(lib, name):
    # no submatrices.py, no assemble.py: radial operators on the host (kore_b200/radial.py, bit for bit the
    # reference's: tests/test_radial.py), the pencil assembled on the GPU, the reference's golden eigenvalue
    # (tests/spinover/reference.eig through the pinned oracle values of the fixture)
    from kore_b200 import radial
    case = load_case(name)
    pj, pp, ops_ref, A_ref, _ = fixture(name)
    ops = radial.radial_operators(pp)
    assert sorted(ops) == sorted(ops_ref)
    m = case.meta
    with lib.Solver(0) as s:
        asm.assemble(s, pp, ops)
        ip, ix, v = s.get_assembled("A")
        # (not the pattern: the last bit of a generated operator is this host's BLAS's, and an entry of A that
        # cancels to an exact zero on the machine that wrote the fixture may be 1e-17 here)
        A = sp.csr_matrix((v, ix, ip), shape=A_ref.shape)
        assert abs(A - A_ref).max() <= 1e-13 * np.max(np.abs(A_ref.data))
        s.set_chain(case.perm, case.nodeptr)
        s.factor(case.tau)
        lam, X, info = s.eigs(m["nev"], which=m["which_eigenpairs"], target=case.tau, tol=m["tol"], maxit=m["maxit"])
    assert info["nconv"] >= m["nev"]
    for z in case.oracle["eig"]:
        assert np.min(np.abs(lam - z)) <= 1e-9 * abs(z), (z, lam)


@pytest.mark.gpu
def test_twin_driver_from_the_parameter_file_alone(tmp_path, monkeypatch, lib):
    # `python -m kore_b200.solve -kb_operators` in a directory that holds bin/parameters.py and nothing else: no
    # submatrices.py, no assemble.py, no matrix file -- and the reference's golden eigenvalue (tests/spinover/reference.eig,
    # rtol 1e-8 in tests/test_spinover.py:21-29) in eigenvalues0.dat
    import glob
    import sys
    from kore_b200 import solve as drv
    from test_solve_driver import write_operator_dir
    c, d = write_operator_dir(tmp_path, "spinover")
    for fn in glob.glob(str(d / "*.mtx")):
        os.remove(fn)
    monkeypatch.chdir(d)
    sys.modules.pop("parameters", None)
    try:
        assert drv.main(["-st_type", "sinvert", "-kb_operators"]) == 0
    finally:
        sys.modules.pop("parameters", None)
    assert not glob.glob("*.mtx") and not os.path.exists("A.npz")
    ev = np.loadtxt("eigenvalues0.dat").reshape(-1, 2)
    best = ev[np.argmax(ev[:, 0])]                        # test_spinover.py:23: the least damped mode
    golden = c.meta["reference_golden"]["eig"]
    assert np.allclose(best, golden, rtol=1e-8, atol=0)
    assert np.loadtxt("real_flow.field").reshape(2 * c.meta["n"], -1).shape[1] == ev.shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["jones", "dormy"])
def test_gpu_search_from_the_parameter_file_alone(lib, tmp_path, name):
    # the reference's convection tests (tests/test_convection_bouss.py, find_Rac.py) from params.jones / params.dormy04
    # alone: radial operators generated, every trial pencil assembled on the device, golden rows of reference.jones /
    # reference.dormy04 digit for digit
    from kore_b200 import rac
    import test_rac as tr
    ra_min, row = {"jones": (tr.JONES_RA_MIN, tr.JONES_ROW), "dormy": (tr.RA_MIN, tr.GOLDEN_ROW)}[name]
    c = load_case(name)
    m = c.meta
    pp, pen = tr._pencil_from_parameters(name)
    with rac.GrowthRate(pen, c.perm, c.nodeptr, c.tau, m["nev"], m["which_eigenpairs"], tol=m["tol"],
                        maxit=m["maxit"]) as g:
        Ra_c, omega_c, sigma_c = rac.find_rac(g, ra_min)
        assert 3 <= len(g.history) < 20
    p = tmp_path / "critical_params.dat"
    rac.write_critical_params(p, m["Ek"], m["ricb"], Ra_c, m["m"], omega_c)
    assert p.read_text().strip() == row
    assert abs(sigma_c) < 1e-6 * abs(omega_c)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["asm_forcing9", "asm_forcing9_stressfree", "asm_forcing10", "asm_twozone",
                                  "asm_no_thermal_diffusion", "asm_inviscid", "asm_compositional"])
def test_device_assembly_bitwise_of_the_boundary_flow_forcings(lib, name):
    # forcing = 9 / 10 (assemble.py:360-426): with a stress-free outer boundary the poloidal boundary rows depend on
    # the degree, one table entry per block row
    import test_assembly as ta
    ta.test_device_assembly_bitwise(lib, name)


# ------------------------------------------------------------------ double-diffusive diagnostics
@pytest.mark.gpu
def test_device_double_diffusive_integrals_match_reference(lib):
    # kb_diagnose once per scalar field on [u | v | field] (kore_b200/diagnostics.py:diagnose_double_diffusive)
    # against utils4pp.diagnose with csol2 on the reference-assembled double-diffusive pencil's eigenvector
    # (tests/golden/asm_compositional/diagnostics.npz, make_diag_fixtures.py); CPU twin on the kernel's model:
    # tests/test_diagnostics.py::test_double_diffusive_integrals_match_reference
    from kore_b200 import diagnostics as dg
    meta = json.load(open(os.path.join(GOLDEN, "asm_compositional", "meta.json")))
    pj = json.load(open(os.path.join(GOLDEN, "asm_compositional", "asm_params.json")))
    z = np.load(os.path.join(GOLDEN, "asm_compositional", "diagnostics.npz"))

    def close(a, b, tol):
        scale = np.max(np.abs(b), axis=0)
        scale[scale == 0] = 1.0
        return np.max(np.abs(a - b) / scale) <= tol

    X = np.stack([z["x"], 2j * z["x"]], axis=1)
    with lib.Solver(0) as s:
        flow, therm, comp, degs = dg.diagnose_double_diffusive(
            s, X, meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"], thermal=1, heating=pj["heating"],
            comp_background=pj["comp_background"])
    assert flow.shape == (2,) + z["flow"].shape and comp.shape == (2,) + z["comp"].shape
    assert close(flow[0], z["flow"], 1e-9) and close(therm[0], z["thermal"], 1e-9) and close(comp[0], z["comp"], 1e-9)
    assert close(flow[1], 4 * flow[0], 1e-12) and close(comp[1], 4 * comp[0], 1e-12)
    g = dg.differential_gradient_factor(meta["ricb"])
    pb = dg.power_balance(flow[0], therm[0], degs, z["lam"][0], pj["Ek"], pj["ViscosD"], pj["Beyonce"], pj["ThermaD"],
                          comp=comp[0], CompBuoy=pj["OmgTau"] ** 2 * pj["BV2_comp"],
                          CompD=pj["OmgTau"] * pj["Ek"] / pj["Schmidt"], advect_scale_thm=g, advect_scale_cmp=g)
    assert pb["resid1"] < 1e-2 and pb["resid3"] < 1e-4 and pb["resid4"] < 1e-4, pb


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["asm_compositional", "sd_spinover_thermal", "sd_m0_thermal"])
def test_device_spin_doctor_files_match_reference(lib, name):
    # flow.dat / thermal.dat / compositional.dat rows (energies, dissipations, powers, residuals, viscous torques) from
    # kb_diagnose against what the UNMODIFIED bin/spin_doctor.py wrote for the same solutions
    # (tests/golden/make_spin_doctor_fixtures.py); CPU twin on the kernel's model in tests/test_diagnostics.py
    import test_diagnostics as td
    z, p = td.sd_golden(name)
    with lib.Solver(0) as s:
        tables = td.sd_tables(s, z, p)
    td.check_sd_tables(tables, z, 1e-9)


@pytest.mark.gpu
def test_device_magnetic_energy_and_diffusion_match_reference(lib):
    # the flow pass of kb_diagnose on the magnetic block [f | g] with the degrees of bsymm: magnetic energy and
    # diffusion per degree against bdgn[:, 0:2] of utils4pp.diagnose (tests/golden/magnetic_small/diagnostics.npz)
    import test_diagnostics as td
    from kore_b200 import diagnostics as dg
    meta, pj, z = td.golden("magnetic_small")
    n = meta["n"]
    with lib.Solver(0) as s:
        mag, degs = dg.diagnose_magnetic_energy(s, z["x"][2 * n:4 * n], meta["N"], meta["lmax"], meta["m"],
                                                int(z["bsymm"][0]), meta["ricb"])
    assert td.close(mag[0][:, :2], z["magnetic"][:, :2], 1e-9) and np.all(np.isnan(mag[0][:, 2]))


@pytest.mark.gpu
def test_device_own_heating_gradient_closes_the_thermal_balance(lib):
    # 'two zone' heating: the advection integral from a second kb_diagnose launch with scaled quadrature weights
    # (kore_b200/diagnostics.py:diagnose, gradient_series); CPU twin on the kernel's model in tests/test_diagnostics.py
    import test_diagnostics as td
    with lib.Solver(0) as s:
        td.check_twozone_balance(s)
