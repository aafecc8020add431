"""Critical-Rayleigh search (BASELINE.json config 2, tests/dormy2004/find_Rac.py).

CPU: the search logic, the affine pencil A(Ra), and the oracle driven through the same search
reproduces the reference's golden row `tests/dormy2004/reference.dormy04:1` digit for digit.
GPU (-m gpu): the same search through the C ABI reproduces that row too."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_case

GOLDEN_ROW = "2.000e-05 0.35 1.65404e+06 9 -1.10162e-02"  # tests/dormy2004/reference.dormy04:1
RA_MIN = 1.6e6  # find_Rac.py:21


def dormy_pencil():
    import kore_oracle as ko
    from kore_b200 import rac
    c = load_case("dormy")
    A1 = ko.load_csr(os.path.join(GOLDEN, "dormy", "A1.npz"))
    pos, val = rac.slope_positions(c.A, A1)
    return c, rac.RayleighPencil(c.A, c.meta["Ra_gap"], pos, val, c.B), A1


def test_bracket_walks_up_and_down():
    from kore_b200 import rac
    calls = []

    def f(x, a):
        calls.append(x)
        return a * (x - 6.2185)

    x = rac.bracket_brentq(f, 6.2, args=(1.0,), tol=1e-9)
    assert abs(x - 6.2185) < 1e-8
    # first growth rate negative -> walked upwards in steps of dx = 0.01 before Brent
    assert calls[1] == pytest.approx(6.21) and calls[2] == pytest.approx(6.22)
    calls.clear()
    x = rac.bracket_brentq(f, 6.25, args=(1.0,), tol=1e-9)
    assert abs(x - 6.2185) < 1e-8 and calls[1] == pytest.approx(6.24)
    # a supplied x2 that brackets is used as it is; one that does not restarts from the nearer end
    calls.clear()
    x = rac.bracket_brentq(f, 6.0, x2=6.5, args=(1.0,), tol=1e-9)
    assert abs(x - 6.2185) < 1e-8 and calls[:2] == [6.0, 6.5]
    calls.clear()
    x = rac.bracket_brentq(f, 6.0, x2=6.1, args=(1.0,), tol=1e-9)
    assert abs(x - 6.2185) < 1e-8 and calls[2] == pytest.approx(6.11)
    with pytest.raises(RuntimeError):
        rac.bracket_brentq(lambda x: 1.0 + x * x, 0.0, max_walk=20)


def test_rayleigh_pencil_is_the_assembled_matrix():
    # A(Ra) rebuilt from (A_ref, A1) equals what scipy computes, on A's own pattern, for any Ra;
    # unsorted column order inside rows is handled
    from kore_b200 import rac
    c, pen, A1 = dormy_pencil()
    for Ra in (1.6e6, c.meta["Ra_gap"], 1.7e6):
        want = (c.A + (Ra - c.meta["Ra_gap"]) * A1).tocsr()
        got = pen.at(Ra)
        assert np.array_equal(got.indices, c.A.indices) and np.array_equal(got.indptr, c.A.indptr)
        assert abs(got - want).max() <= 1e-16 * abs(want).max()
    # affine_from_two recovers the same slope from two assemblies
    lo, hi = pen.at(1.6e6).copy(), pen.at(1.7e6).copy()
    pos, val = rac.affine_from_two(lo, 1.6e6, hi, 1.7e6)
    assert np.array_equal(np.sort(pos), np.sort(pen.pos))
    o1, o2 = np.argsort(pos), np.argsort(pen.pos)
    assert np.allclose(val[o1], pen.val[o2], rtol=1e-9, atol=0)
    with pytest.raises(ValueError):
        import scipy.sparse as sp
        rac.slope_positions(sp.identity(4, format="csr", dtype=complex), sp.csr_matrix(np.ones((4, 4))))


def test_critical_params_row_format(tmp_path):
    from kore_b200 import rac
    p = tmp_path / "critical_params.dat"
    rac.write_critical_params(p, 2e-5, 0.35, 1654042.2, 9, -0.01101623)
    assert p.read_text().strip() == GOLDEN_ROW


def test_oracle_search_reproduces_reference_row(tmp_path):
    """The oracle (SciPy SuperLU + ARPACK) behind the product's search logic lands on the
    reference's Ra_c and omega_c to every digit the golden holds."""
    import kore_oracle as ko
    from kore_b200 import rac
    c, pen, _ = dormy_pencil()
    m = c.meta

    class OracleGrowth:
        cache = {}

        def __call__(self, x):
            Ra = 10.0 ** x
            if Ra not in self.cache:
                lam, _, _ = ko.eigs(pen.at(Ra), c.B, c.tau, m["nev"], m["which_eigenpairs"])
                self.cache[Ra] = lam[np.argmax(lam.real)]
            return self.cache[Ra].real

    g = OracleGrowth()
    Ra_c, omega_c, sigma_c = rac.find_rac(g, RA_MIN)
    assert len(g.cache) < 20
    p = tmp_path / "critical_params.dat"
    rac.write_critical_params(p, m["Ek"], m["ricb"], Ra_c, m["m"], omega_c)
    assert p.read_text().strip() == GOLDEN_ROW
    assert abs(sigma_c) < 1e-6 * abs(omega_c)
    assert Ra_c == pytest.approx(m["Ra_gap"], rel=5e-6)  # params.dormy04:177 holds more digits


@pytest.mark.gpu
def test_gpu_search_reproduces_reference_row(lib, tmp_path):
    from kore_b200 import rac
    c, pen, _ = dormy_pencil()
    m = c.meta
    with rac.GrowthRate(pen, c.perm, c.nodeptr, c.tau, m["nev"], m["which_eigenpairs"], tol=m["tol"],
                        maxit=m["maxit"]) as g:
        Ra_c, omega_c, sigma_c = rac.find_rac(g, RA_MIN)
        hist = list(g.history)
    assert 3 <= len(hist) < 20
    p = tmp_path / "critical_params.dat"
    rac.write_critical_params(p, m["Ek"], m["ricb"], Ra_c, m["m"], omega_c)
    assert p.read_text().strip() == GOLDEN_ROW
    assert abs(sigma_c) < 1e-6 * abs(omega_c)
    assert Ra_c == pytest.approx(m["Ra_gap"], rel=5e-6)
    # the marginal eigenvalue at the golden Ra is the oracle's
    lam_o = c.oracle["eig"][np.argmax(c.oracle["eig"].real)]
    with rac.GrowthRate(pen, c.perm, c.nodeptr, c.tau, m["nev"], m["which_eigenpairs"], tol=m["tol"],
                        maxit=m["maxit"]) as g:
        g(np.log10(m["Ra_gap"]))
        lam = list(g.cache.values())[0]
    assert abs(lam - lam_o) <= 1e-9 * abs(lam_o)


def _dormy_assembly_inputs():
    import json
    import os
    from conftest import GOLDEN
    from kore_b200 import assembly as asm
    d = os.path.join(GOLDEN, "dormy")
    pj = json.load(open(os.path.join(d, "asm_params.json")))
    pp = asm.PhysicsParams.from_dict(pj)
    pp.Bnorm_fixture = pj["Bnorm"]  # ||B||_F of the reference run that wrote the fixture (pins the last bit)
    return pp, asm.load_operators_npz(os.path.join(d, "operators.npz"))


def test_buoyancy_factor_is_the_parameter_files():
    # params.dormy04:177-186 evaluated by the reference's own parameters.py (stored in asm_params.json)
    from kore_b200 import rac
    c = load_case("dormy")
    pp, _ = _dormy_assembly_inputs()
    assert rac.buoyancy_factor(c.meta["Ra_gap"], pp.Ek, pp.ricb, 1) == pp.Beyonce


@pytest.mark.gpu
def test_gpu_search_on_device_assembled_pencils(lib, tmp_path):
    # the same search with every trial matrix assembled on the GPU (no A.npz, no host update)
    from kore_b200 import rac
    c = load_case("dormy")
    m = c.meta
    pp, ops = _dormy_assembly_inputs()
    pen = rac.AssembledPencil(pp, ops, lambda Ra: rac.buoyancy_factor(Ra, pp.Ek, pp.ricb, 1), bnorm=pp.Bnorm_fixture)
    with rac.GrowthRate(pen, c.perm, c.nodeptr, c.tau, m["nev"], m["which_eigenpairs"], tol=m["tol"],
                        maxit=m["maxit"]) as g:
        Ra_c, omega_c, sigma_c = rac.find_rac(g, RA_MIN)
        assert 3 <= len(g.history) < 20
        # at the reference's own Ra_gap the assembled matrix is the fixture's, bit for bit
        pen.install(g.solver, m["Ra_gap"])
        ip, ix, v = g.solver.get_assembled("A")
    p = tmp_path / "critical_params.dat"
    rac.write_critical_params(p, m["Ek"], m["ricb"], Ra_c, m["m"], omega_c)
    assert p.read_text().strip() == GOLDEN_ROW
    assert Ra_c == pytest.approx(m["Ra_gap"], rel=5e-6)
    A = c.A.copy()
    A.sort_indices()
    assert np.array_equal(ip, A.indptr) and np.array_equal(ix, A.indices) and np.array_equal(v, A.data)


# ------------------------------------------------- the reference's third test, from the parameter file alone
JONES_ROW = "6.325e-05 0.00 4.66986e+06 9 -1.93444e-02"  # tests/jones2000/reference.jones:1
JONES_RA_MIN = 4.6e6  # tests/jones2000/find_Rac.py:21


def _pencil_from_parameters(name):
    """The physics parameters of the reference's params file as its parameters.py evaluates them -- and nothing
    else: the radial operators come from kore_b200/radial.py, every trial pencil from the assembly programs."""
    import json
    from kore_b200 import assembly as asm
    from kore_b200 import rac
    pp = asm.PhysicsParams.from_dict(json.load(open(os.path.join(GOLDEN, name, "asm_params.json"))))
    return pp, rac.AssembledPencil(pp, None, lambda Ra: rac.buoyancy_factor(Ra, pp.Ek, pp.ricb, 1))


def _jones_from_parameters():
    """Full sphere, internal heating (tests/test_convection_bouss.py:10-25)"""
    return _pencil_from_parameters("jones")


@pytest.mark.parametrize("name, ra_min, row", [("jones", JONES_RA_MIN, JONES_ROW), ("dormy", RA_MIN, GOLDEN_ROW)])
def test_search_from_the_parameter_file_alone(tmp_path, name, ra_min, row):
    """tests/jones2000/find_Rac.py and tests/dormy2004/find_Rac.py end to end on the CPU stand-ins (assembly programs
    evaluated by the NumPy model of the kernel, eigenvalues by the oracle): the golden rows of reference.jones and
    reference.dormy04, digit for digit."""
    import kore_oracle as ko
    from kore_b200 import rac
    from test_assembly import ModelSolver
    c = load_case(name)
    m = c.meta
    pp, pen = _pencil_from_parameters(name)
    s = ModelSolver()

    class Growth:
        def __init__(self):
            self.cache = {}

        def __call__(self, x):
            Ra = 10.0 ** x
            if Ra not in self.cache:
                pen.install(s, Ra)
                lam, _, _ = ko.eigs(s.M["A"], s.M["B"], c.tau, m["nev"], m["which_eigenpairs"])
                self.cache[Ra] = lam[np.argmax(lam.real)]
            return self.cache[Ra].real

    g = Growth()
    Ra_c, omega_c, sigma_c = rac.find_rac(g, ra_min)
    assert len(g.cache) < 20
    p = tmp_path / "critical_params.dat"
    rac.write_critical_params(p, m["Ek"], m["ricb"], Ra_c, m["m"], omega_c)
    assert p.read_text().strip() == row
    assert abs(sigma_c) < 1e-6 * abs(omega_c)
