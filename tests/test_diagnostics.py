"""Power-balance diagnostics: kore_b200/diagnostics.py + csrc/kb_diag.cu against
what the UNMODIFIED reference post-processing computes (tests/golden/*/diagnostics.npz, written by
tests/golden/make_diag_fixtures.py with bin/utils4pp.py: expand_reshape_sol + diagnose) for an
eigenvector of the spin-over, dormy (thermal, differential heating) and jones (full sphere, internal
heating) pencils.  Floating point: the kernel sums the same integrands with a different series
evaluation, so the bar is relative 1e-9 of the largest degree (observed 1e-13 ... 1e-15)."""
import json
import os

import numpy as np
import pytest

import diag_model as dm
from conftest import GOLDEN, load_case
from kore_b200 import chain, diagnostics as dg

CASES = [("spinover", "differential"), ("dormy", "differential"), ("jones", "internal")]


def golden(name):
    meta = json.load(open(os.path.join(GOLDEN, name, "meta.json")))
    pj = json.load(open(os.path.join(GOLDEN, name, "asm_params.json")))
    z = np.load(os.path.join(GOLDEN, name, "diagnostics.npz"))
    return meta, pj, z


def close(a, b, tol):
    scale = np.max(np.abs(b), axis=0)
    scale[scale == 0] = 1.0
    return np.max(np.abs(a - b) / scale) <= tol


@pytest.mark.parametrize("name,heating", CASES)
def test_model_reproduces_reference_integrals(name, heating):
    meta, pj, z = golden(name)
    flow, therm = dm.diagnose(z["x"], meta, heating)
    assert z["flow"].shape == flow.shape == (meta["lmax"] - meta["m"] + 1, 6)
    assert close(flow, z["flow"], 1e-12)
    if meta["thermal"]:
        assert close(therm, z["thermal"], 1e-12)


def test_quadrature_nodes():
    nodes = dg.quadrature_nodes(64, 0.35)
    x0, rk, w = nodes
    assert np.allclose(rk, 0.35 + 0.325 * (x0 + 1)) and rk.min() > 0.35 and rk.max() < 1.0
    # Chebyshev-Gauss: int_ricb^rcmb r^2 dr to spectral accuracy
    assert abs(np.sum(w * rk ** 2) - (1 - 0.35 ** 3) / 3) < 1e-4
    x0, rk, w = dg.quadrature_nodes(64, 0.0)
    assert np.allclose(x0, rk) and rk.min() > 0 and abs(np.sum(w) - 1.0) < 1e-3


@pytest.mark.parametrize("name,heating", CASES)
def test_power_balance_of_the_reference_integrals(name, heating):
    # 2 sigma KE = viscous dissipation - buoyancy power holds for an eigenvector: the check the reference's
    # users make (spin_doctor.py:227-242), here on the reference's own integrals
    meta, pj, z = golden(name)
    pb = dg.power_balance(z["flow"], z["thermal"], chain.ell(meta["m"], meta["lmax"], meta["symm"]), z["lam"][0],
                          pj["Ek"], pj["ViscosD"], pj["Beyonce"], pj["ThermaD"])
    assert pb["resid1"] < 1e-4 and pb["resid0"] < 1e-2
    assert abs(pb["KP"] + pb["KT"] - pb["KE"]) <= 1e-12 * pb["KE"]
    if name == "jones":
        assert pb["resid3"] < 1e-9


class ModelSolver:
    """Stand-in for lib.Solver.diagnose on CPU: the NumPy model of the kernel (tests/diag_model.py) behind the
    same call, so that the HOST logic of kore_b200/diagnostics.py runs in the CPU suite."""

    def diagnose(self, p, nodes, X):
        meta = dict(N=p.N, N1=p.N1, n=p.N1 * p.nb, ricb=p.ricb, m=p.m, lmax=p.lmax, symm=p.symm, thermal=p.thermal)
        out = [dm.diagnose(X[:, k], meta, "differential" if p.heating == 0 else "internal") for k in range(X.shape[1])]
        return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


def test_double_diffusive_integrals_match_reference():
    # [u | v | h | c]: utils4pp.diagnose with csol2 (thermal_worker with the 'compositional' flag, the second
    # buoyancy column of flow_worker) on the eigenvector of the reference-assembled double-diffusive pencil
    meta, pj, z = golden("asm_compositional")
    assert meta["sizmat"] == 4 * meta["n"] and z["comp"].shape == z["thermal"].shape
    flow, therm, comp, degs = dg.diagnose_double_diffusive(
        ModelSolver(), z["x"], meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"], thermal=1,
        heating=pj["heating"], comp_background=pj["comp_background"])
    assert flow.shape == (1,) + z["flow"].shape and np.all(z["flow"][:, 5][np.isin(degs[2], degs[0])] != 0)
    assert close(flow[0], z["flow"], 1e-12)
    assert close(therm[0], z["thermal"], 1e-12) and close(comp[0], z["comp"], 1e-12)
    # without the heat equation: [u | v | c]
    n = meta["n"]
    x3 = np.concatenate([z["x"][:2 * n], z["x"][3 * n:]])
    f3, t3, c3, _ = dg.diagnose_double_diffusive(ModelSolver(), x3, meta["N"], meta["lmax"], meta["m"], meta["symm"],
                                                 meta["ricb"], thermal=0, comp_background=pj["comp_background"])
    assert close(c3[0], z["comp"], 1e-12) and not np.any(t3) and not np.any(f3[0][:, 4])
    assert close(f3[0][:, [0, 1, 2, 5]], z["flow"][:, [0, 1, 2, 5]], 1e-12)
    with pytest.raises(ValueError):
        dg.diagnose_double_diffusive(ModelSolver(), x3, meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"])


def test_double_diffusive_power_balance_closes():
    # momentum: both buoyancy powers enter; heat and composition: 2 sigma E = D + W once the advection carries the
    # ricb / gap of the 'differential' gradient that operators.py:736, 802 has and utils4pp.py:409, 419 leaves out
    meta, pj, z = golden("asm_compositional")
    degs = chain.ell(meta["m"], meta["lmax"], meta["symm"])
    kw = dict(comp=z["comp"], CompBuoy=pj["OmgTau"] ** 2 * pj["BV2_comp"], CompD=pj["OmgTau"] * pj["Ek"] / pj["Schmidt"])
    pb = dg.power_balance(z["flow"], z["thermal"], degs, z["lam"][0], pj["Ek"], pj["ViscosD"], pj["Beyonce"],
                          pj["ThermaD"], **kw)
    assert pb["resid1"] < 1e-2 and pb["Wcmp"] != 0 and pb["Wthm"] != 0
    # ... and the momentum balance does NOT close without the compositional power
    assert dg.power_balance(z["flow"], z["thermal"], degs, z["lam"][0], pj["Ek"], pj["ViscosD"], pj["Beyonce"],
                            pj["ThermaD"])["resid1"] > 0.5
    assert pb["resid3"] > 0.1 and pb["resid4"] > 0.1      # the reference's own numbers (no gradient constant)
    g = dg.differential_gradient_factor(meta["ricb"])
    pb = dg.power_balance(z["flow"], z["thermal"], degs, z["lam"][0], pj["Ek"], pj["ViscosD"], pj["Beyonce"],
                          pj["ThermaD"], advect_scale_thm=g, advect_scale_cmp=g, **kw)
    assert pb["resid3"] < 1e-4 and pb["resid4"] < 1e-4, (pb["resid3"], pb["resid4"])  # 1.5e-6, 1.4e-5 at N = 24


def test_driver_writes_the_double_diffusive_power_balance(tmp_path, monkeypatch):
    # solve.py -kb_diagnose on a run with the composition equation: one row per solution, 17 columns
    import types
    from kore_b200 import solve
    meta, pj, z = golden("asm_compositional")
    par = types.SimpleNamespace(**{k: v for k, v in pj.items() if not isinstance(v, (list, dict))})
    monkeypatch.chdir(tmp_path)
    solve.write_power_balance(par, ModelSolver(), np.stack([z["x"], -z["x"]], axis=1), [z["lam"][0]] * 2)
    rows = np.loadtxt("power_balance.dat")
    assert rows.shape == (2, 17) and np.array_equal(rows[0], rows[1])
    KE, Wthm, resid1, TE, Wcmp, CE, Dcmp = rows[0][[0, 5, 7, 8, 12, 13, 14]]
    assert np.isclose(KE, z["flow"][:, 0].sum(), rtol=1e-12) and np.isclose(TE, z["thermal"][:, 0].sum(), rtol=1e-12)
    assert np.isclose(CE, z["comp"][:, 0].sum(), rtol=1e-12) and resid1 < 1e-2
    assert np.isclose(Wcmp, pj["BV2_comp"] * z["flow"][:, 5].sum(), rtol=1e-12)
    assert np.isclose(Dcmp, pj["Ek"] / pj["Schmidt"] * z["comp"][:, 1].sum(), rtol=1e-12)
    assert rows[0][11] < 1e-4 and rows[0][16] < 1e-4   # resid3, resid4: the driver's own file closes the balances
    # ... and spin_doctor.py's own files, appended to on a second call as the reference does
    zs, p = sd_golden("asm_compositional")
    par.BV2, par.Etherm, par.Ecomp = p["BV2"], p["Etherm"], p["Ecomp"]
    for f in ("flow.dat", "thermal.dat", "compositional.dat", "eigenvalues.dat"):
        os.remove(f)
    for _ in range(2):
        solve.write_power_balance(par, ModelSolver(), zs["X"], zs["lam"])
    nsol = len(zs["lam"])
    tables = {k: np.loadtxt(k + ".dat") for k in ("flow", "thermal", "compositional")}
    assert all(t.shape[0] == 2 * nsol and np.array_equal(t[:nsol], t[nsol:]) for t in tables.values())
    check_sd_tables({k: t[:nsol] for k, t in tables.items()}, zs, 1e-11)
    ev = np.loadtxt("eigenvalues.dat")
    assert ev.shape == (2 * nsol, 2) and np.array_equal(ev[:nsol, 0] + 1j * ev[:nsol, 1], zs["lam"])


SD_CASES = ["asm_compositional", "sd_spinover_thermal", "sd_m0_thermal"]


def sd_golden(name):
    z = np.load(os.path.join(GOLDEN, name, "spin_doctor.npz"))
    return z, json.loads(str(z["params"]))


def sd_tables(solver, z, p):
    geom = (p["N"], p["lmax"], p["m"], p["symm"], p["ricb"])
    comp = None
    if p["compositional"]:
        flow, therm, comp, degs = dg.diagnose_double_diffusive(solver, z["X"], *geom, thermal=p["thermal"],
                                                               heating=p["heating"], comp_background=p["comp_background"])
    else:
        flow, therm, degs = dg.diagnose(solver, z["X"], *geom, thermal=p["thermal"], heating=p["heating"])
    vt, vi = dg.viscous_torques(z["X"], *geom, p["Ek"])
    return dg.spin_doctor_tables(flow, therm, comp, degs, z["lam"], p["Ek"], p["OmgTau"], p["BV2"], p["BV2_comp"],
                                 p["Etherm"], p["Ecomp"], vt, vi)


def check_sd_tables(tables, z, tol):
    # flow.dat / thermal.dat / compositional.dat as the UNMODIFIED bin/spin_doctor.py wrote them for the same
    # solutions (tests/golden/make_spin_doctor_fixtures.py); residual columns are differences of nearly equal
    # terms, so their bar is absolute
    for k, cols_abs in (("flow", [8, 9]), ("thermal", [3]), ("compositional", [])):
        if k + "_dat" not in z:
            assert k not in tables
            continue
        ref, got = z[k + "_dat"], tables[k]
        assert got.shape == ref.shape
        scale = np.maximum(np.max(np.abs(ref), axis=0), 1e-300)
        err = np.abs(got - ref) / scale
        rel = [c for c in range(ref.shape[1]) if c not in cols_abs]
        assert err[:, rel].max() <= tol, (k, err.max(axis=0))
        if cols_abs:
            assert np.max(np.abs(got[:, cols_abs] - ref[:, cols_abs])) <= max(tol, 1e-9), k


@pytest.mark.parametrize("name", SD_CASES)
def test_spin_doctor_files_match_reference(name):
    z, p = sd_golden(name)
    check_sd_tables(sd_tables(ModelSolver(), z, p), z, 1e-11)


def test_viscous_torques_of_the_reference_cases():
    # m = 1 antisymmetric: equatorial torque on the mantle only; m = 0 symmetric: axial torques on mantle and inner
    # core; any other m: none (utils.py:1288-1399 through spin_doctor.py:164-166)
    z, p = sd_golden("sd_spinover_thermal")
    assert np.all(np.abs(z["flow_dat"][:, 10:12]).sum(axis=1) > 0) and not np.any(z["flow_dat"][:, 12:])
    z, p = sd_golden("sd_m0_thermal")
    assert np.all(np.abs(z["flow_dat"][:, 10:]).min(axis=1) > 0)
    vt, vi = dg.viscous_torques(z["X"], p["N"], p["lmax"], p["m"], p["symm"], p["ricb"], p["Ek"])
    assert np.allclose(vt, z["flow_dat"][:, 10] + 1j * z["flow_dat"][:, 11], rtol=1e-11, atol=0)
    assert np.allclose(vi, z["flow_dat"][:, 12] + 1j * z["flow_dat"][:, 13], rtol=1e-11, atol=0)
    z, p = sd_golden("asm_compositional")
    vt, vi = dg.viscous_torques(z["X"], p["N"], p["lmax"], p["m"], p["symm"], p["ricb"], p["Ek"])
    assert not np.any(vt) and not np.any(vi) and not np.any(z["flow_dat"][:, 10:])


def test_magnetic_energy_and_diffusion_match_reference():
    # bdgn[:, 0:2] of utils4pp.diagnose on the eigenvector of the reference-assembled dipole-field pencil
    # ([u | v | f | g], bsymm = symm * symmB0 = +1), and the flow columns of the same solution
    meta, pj, z = golden("magnetic_small")
    n, bsymm = meta["n"], int(z["bsymm"][0])
    assert bsymm == meta["symm"] * meta["symmB0"] and z["magnetic"].shape == (meta["lmax"] - meta["m"] + 1, 3)
    geom = (meta["N"], meta["lmax"], meta["m"])
    mag, degs = dg.diagnose_magnetic_energy(ModelSolver(), z["x"][2 * n:4 * n], *geom, bsymm, meta["ricb"])
    assert close(mag[0][:, :2], z["magnetic"][:, :2], 1e-12) and np.all(np.isnan(mag[0][:, 2]))
    assert np.all(z["magnetic"][:, 0] > 0)
    flow, _, _ = dg.diagnose(ModelSolver(), z["x"][:2 * n], *geom, meta["symm"], meta["ricb"])
    assert close(flow[0][:, :3], z["flow"][:, :3], 1e-12)


def twozone_case():
    import kore_oracle as ko
    d = os.path.join(GOLDEN, "asm_twozone")
    meta, pj = json.load(open(os.path.join(d, "meta.json"))), json.load(open(os.path.join(d, "asm_params.json")))
    lam, X, info = ko.eigs(ko.load_csr(os.path.join(d, "A.npz")), ko.load_csr(os.path.join(d, "B.npz")),
                           complex(meta["rtau"], meta["itau"]), meta["nev"], meta["which_eigenpairs"])
    return meta, pj, lam, X, np.load(os.path.join(d, "radprofs.npz"))["cd_ent"][:, 0]


def check_twozone_balance(solver):
    # 'two zone' heating (a background gradient of the run's own): utils4pp.py:410-411 cannot run in the reference
    # (ut.twozone does not exist), so the check is the physics: on eigenvectors of the reference-assembled pencil
    # 2 sigma TE = ThermaD Dthm + Wadv closes with the sign of operators.py:738, and does not with the other one
    meta, pj, lam, X, cd_ent = twozone_case()
    assert pj["heating"] == "two zone"
    geom = (meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"])
    with pytest.raises(ValueError):
        dg.diagnose(solver, X, *geom, thermal=1, heating="two zone")
    flow, therm, degs = dg.diagnose(solver, X, *geom, thermal=1, heating="two zone", gradient_series=cd_ent)
    ref, _, _ = dg.diagnose(solver, X, *geom, thermal=1, heating="internal")
    assert np.array_equal(flow, ref)      # only the advection column depends on the gradient
    for i in range(len(lam)):
        r = [dg.power_balance(flow[i], therm[i], degs, lam[i], pj["Ek"], pj["ViscosD"], pj["Beyonce"], pj["ThermaD"],
                              advect_scale_thm=sc)["resid3"] for sc in (-1.0, 1.0)]
        assert r[0] < 1.5e-2 and r[1] > 2 * r[0], (i, r)   # N = 48, lmax = 16: resid1 itself is 2-5e-2 here


def test_own_heating_gradient_closes_the_thermal_balance():
    check_twozone_balance(ModelSolver())


def test_driver_takes_the_gradient_from_the_runs_profile_tables(tmp_path, monkeypatch):
    # solve.py -kb_diagnose under 'two zone' heating: cd_ent from the run's radProfs.mat (compute_profiles.py) through
    # radial.run_profiles, resid3 of power_balance.dat closes
    import types
    import scipy.io as sio
    from kore_b200 import solve
    meta, pj, lam, X, cd_ent = twozone_case()
    monkeypatch.chdir(tmp_path)
    sio.savemat("radProfs.mat", {"cd_ent": cd_ent.reshape(-1, 1)})
    solve.write_power_balance(types.SimpleNamespace(**pj), ModelSolver(), X, lam)
    rows = np.loadtxt("power_balance.dat")
    assert rows.shape == (len(lam), 12) and rows[:, 11].max() < 1.5e-2 and not os.path.exists("flow.dat")


@pytest.mark.parametrize("ricb", [0.35, 0.0])
def test_viscous_torque_is_the_boundary_value_of_the_series(ricb):
    # independent of the reference: (8 pi / 3) R^2 (R T'(R) - T(R)) of the degree-1 toroidal scalar, with T and T'
    # from numpy's Chebyshev class on the solution's radial domain; full sphere: odd polynomials only (m = 0,
    # symmetric flow: toroidal parity (m + s) % 2 = 1)
    from numpy.polynomial import chebyshev as C
    rng = np.random.default_rng(5)
    N, lmax, m, symm, Ek = 24, 8, 0, 1, 1e-3
    N1 = N if ricb > 0 else N // 2
    n = N1 * ((lmax - m + 1) // 2)
    X = rng.standard_normal((2 * n, 2)) + 1j * rng.standard_normal((2 * n, 2))
    vt, vi = dg.viscous_torques(X, N, lmax, m, symm, ricb, Ek)
    for k in range(2):
        c = X[n:n + N1, k]
        if ricb == 0:
            full = np.zeros(N, dtype=complex)
            full[1::2] = c
            c = full
        T = C.Chebyshev(c, domain=[ricb if ricb > 0 else -1.0, 1.0])
        for got, R in ((vt[k], 1.0), (vi[k], ricb)):
            want = Ek * (8 * np.pi / 3) * R ** 2 * (R * T.deriv()(R) - T(R)) if R > 0 else 0.0
            assert abs(got - want) <= 1e-11 * max(1.0, abs(want)), (ricb, R, got, want)


# ---------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,heating", CASES)
def test_device_integrals_match_reference(lib, name, heating):
    meta, pj, z = golden(name)
    with lib.Solver(0) as s:
        flow, therm, degs = dg.diagnose(s, z["x"], meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"],
                                        thermal=meta["thermal"], heating=heating)
    assert flow.shape == (1,) + z["flow"].shape
    assert close(flow[0], z["flow"], 1e-9)
    if meta["thermal"]:
        assert close(therm[0], z["thermal"], 1e-9)
    fm, tm = dm.diagnose(z["x"], meta, heating)
    assert close(flow[0], fm, 1e-11)


@pytest.mark.gpu
def test_device_integrals_of_several_solutions(lib):
    # the integrals are quadratic in the solution; every column is processed independently
    meta, pj, z = golden("dormy")
    X = np.stack([z["x"], 2.0 * z["x"], (1 + 1j) * z["x"]], axis=1)
    with lib.Solver(0) as s:
        flow, therm, _ = dg.diagnose(s, X, meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"],
                                     thermal=1, heating="differential")
    assert close(flow[1], 4.0 * flow[0], 1e-13) and close(flow[2], 2.0 * flow[0], 1e-13)
    assert close(therm[1], 4.0 * therm[0], 1e-13) and close(therm[2], 2.0 * therm[0], 1e-13)


@pytest.mark.gpu
def test_eigs_then_diagnose_closes_the_power_balance(lib):
    # end to end on the GPU: spin-over eigenpairs from kb_eigs straight into kb_diagnose
    case = load_case("spinover")
    m = case.meta
    pj = json.load(open(os.path.join(GOLDEN, "spinover", "asm_params.json")))
    with lib.Solver(0) as s:
        s.set_pencil(case.A, case.B)
        s.set_chain(case.perm, case.nodeptr)
        s.factor(case.tau)
        lam, X, info = s.eigs(m["nev"], which=m["which_eigenpairs"], target=case.tau, tol=m["tol"], maxit=m["maxit"])
        flow, therm, degs = dg.diagnose(s, X, m["N"], m["lmax"], m["m"], m["symm"], m["ricb"])
    res = [dg.power_balance(flow[k], None, degs, lam[k], pj["Ek"], pj["ViscosD"])["resid1"] for k in range(len(lam))]
    assert res[int(np.argmax(lam.real))] < 1e-4, res   # the spin-over mode itself (1.2e-6 on the oracle's vector)
    assert max(res) < 1e-2, res                        # its neighbours are less well resolved at N = 68


@pytest.mark.gpu
def test_diagnose_rejects_inconsistent_sizes(lib):
    meta, pj, z = golden("spinover")
    with lib.Solver(0) as s:
        with pytest.raises(ValueError):
            dg.diagnose(s, z["x"][:-1], meta["N"], meta["lmax"], meta["m"], meta["symm"], meta["ricb"])
        p = lib.KbDiagParams(N=meta["N"], N1=meta["N"] // 2, nb=32, m=1, lmax=64, symm=-1, thermal=0, heating=0,
                             ricb=0.35, rcmb=1.0)
        with pytest.raises(lib.KoreB200Error):
            s.diagnose(p, dg.quadrature_nodes(meta["N"], 0.35), z["x"])
