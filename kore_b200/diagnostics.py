"""Power-balance diagnostics of a solution on the GPU.

The reference post-processes every solution in bin/spin_doctor.py:105-242: the eigenvector is cut
into per-degree Chebyshev series, `utils4pp.diagnose` (utils4pp.py:794-858) integrates kinetic
energy, viscous dissipation, buoyancy power, thermal energy / dissipation / advection degree by degree
in a multiprocessing pool, and the sums enter the power-balance residuals the physicists use as their
own correctness check (spin_doctor.py:227-242).  Here the integrals are one kernel launch for all
degrees and all solutions (`kb_diagnose`, csrc/kb_diag.cu); this module prepares the quadrature
nodes, calls it and forms the same sums and residuals.  Hydrodynamic, Boussinesq thermal and
double-diffusive (thermal + compositional) solutions; of magnetic solutions the energy and diffusion
of the induced field only (no Lorentz / induction powers).
"""
from __future__ import annotations

import numpy as np

from . import chain as _chain
from . import lib as _lib


def quadrature_nodes(N, ricb, rcmb=1.0, Ra=None, Rb=None):
    """(3, N): the Chebyshev-Gauss nodes x_k = cos((k + 1/2) pi / N) mapped to r in [Ra, Rb] and then
    into the Chebyshev domain of the solution ([ricb, rcmb], or [-rcmb, rcmb] without inner core), the
    radii, and the weights (pi / N) sqrt(1 - x_k^2) (Rb - Ra) / 2 (utils4pp.py:15-24, 54-64, 806-822)."""
    Ra = ricb if Ra is None else Ra
    Rb = rcmb if Rb is None else Rb
    k = np.arange(N)
    xk = np.cos((k + 0.5) * np.pi / N)
    rk = 0.5 * (Rb - Ra) * (xk + 1) + Ra
    r0 = ricb if ricb > 0 else -rcmb
    x0 = 2 * (rk - r0) / (rcmb - r0) - 1
    w = (np.pi / N) * np.sqrt(1 - xk ** 2) * (Rb - Ra) / 2
    return np.vstack([x0, rk, w])


def diagnose(solver, X, N, lmax, m, symm, ricb, thermal=0, heating="differential", rcmb=1.0, Ra=None, Rb=None,
             gradient_series=None):
    """Per-degree integrals of the solutions in the columns of X (Kore ordering [u | v | h]).

    Returns (flow, therm, degrees): flow[nsol, n_l, 6] with the columns of the reference's `udgn`
    (kinetic energy, kinetic dissipation, internal dissipation, 0, buoyancy power, 0), therm[nsol, nb, 3]
    (`tdgn`: thermal energy, dissipation, advection), degrees = (poloidal, toroidal, all).

    heating = 'two zone' / 'user defined': the advection integrand carries `fr = r * twozone(r, args)` (or
    `r * BVprof(r, args)`, utils4pp.py:410-413) instead of r^2 or 1/r; `gradient_series` are the Chebyshev
    coefficients of exactly that function on the solution's radial domain -- the `cd_ent` table the radial
    operators of such a run are built from (kore_b200/radial.py:profile_tables, submatrices.py:407) -- and the
    integral comes from a second launch of the same kernel with the quadrature weights scaled by fr / r^2 under
    'internal' heating (the weights enter every integrand as a plain factor).  The reference itself cannot run
    this branch (it calls `ut.twozone`, which lives in radial_profiles.py), so there is no golden; the sign is
    the one utils4pp.py writes down, and since the heat equation carries the gradient with a minus
    (operators.py:738) the thermal balance closes with `power_balance(advect_scale_thm=-1)`."""
    own_gradient = heating in ("two zone", "user defined")
    if own_gradient and thermal and gradient_series is None:
        raise ValueError("heating = %r needs gradient_series (the cd_ent profile table)" % (heating,))
    if heating not in ("differential", "internal") and not own_gradient:
        raise NotImplementedError("heating = %r" % (heating,))
    N1 = N if ricb > 0 else N // 2
    nb = (lmax - m + 1) // 2
    p = _lib.KbDiagParams(N=N, N1=N1, nb=nb, m=m, lmax=lmax, symm=symm, thermal=int(bool(thermal)),
                          heating=0 if heating == "differential" else 1, ricb=float(ricb), rcmb=float(rcmb))
    X = np.asarray(X, dtype=np.complex128)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    want = N1 * nb * (3 if thermal else 2)
    if X.shape[0] != want:
        raise ValueError("solutions have %d rows, the parameters imply %d (hydro%s only)"
                         % (X.shape[0], want, " + thermal" if thermal else ""))
    nodes = quadrature_nodes(N, ricb, rcmb, Ra, Rb)
    flow, therm = solver.diagnose(p, nodes, X)
    if own_gradient and thermal:
        fr = np.polynomial.chebyshev.chebval(nodes[0], np.asarray(gradient_series, dtype=float).ravel())
        scaled = nodes.copy()
        scaled[2] = nodes[2] * fr / nodes[1] ** 2
        therm[:, :, 2] = solver.diagnose(p, scaled, X)[1][:, :, 2]
    return flow, therm, _chain.ell(m, lmax, symm)


def diagnose_double_diffusive(solver, X, N, lmax, m, symm, ricb, thermal=1, heating="differential",
                              comp_background="differential", rcmb=1.0, Ra=None, Rb=None, gradient_series=None):
    """`diagnose` for runs with the composition equation (par.compositional = 1): the columns of X are
    [u | v | h | c] (or [u | v | c] without the heat equation, solve.py:262-273 / assemble.py:442-534
    section order).

    The reference integrates the compositional field with the thermal worker itself (utils4pp.py:846-850:
    `thermal_worker(..., csol2, ..., 'compositional')`; the buoyancy power of the composition is
    `buoyancy_power` with `clm0` in the place of `hlm0`, utils4pp.py:452-466), the background being
    `par.comp_background` instead of `par.heating` (utils4pp.py:414-419).  So the same kernel runs once
    per scalar field on [u | v | field]: no second code path on the device.

    Returns (flow[nsol, n_l, 6], therm[nsol, nb, 3], comp[nsol, nb, 3], degrees): `flow` with BOTH buoyancy
    columns of `udgn` (4: thermal, 5: compositional), `therm` = `tdgn`, `comp` = `cdgn`."""
    if comp_background not in ("differential", "internal"):
        raise NotImplementedError("comp_background = %r" % (comp_background,))
    N1 = N if ricb > 0 else N // 2
    nb = (lmax - m + 1) // 2
    n = N1 * nb
    X = np.asarray(X, dtype=np.complex128)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    want = n * (3 + int(bool(thermal)))
    if X.shape[0] != want:
        raise ValueError("solutions have %d rows, the parameters imply %d (hydro%s + compositional)"
                         % (X.shape[0], want, " + thermal" if thermal else ""))
    kw = dict(rcmb=rcmb, Ra=Ra, Rb=Rb)
    c0 = want - n
    flow_c, comp, degrees = diagnose(solver, np.vstack([X[:2 * n], X[c0:]]), N, lmax, m, symm, ricb, thermal=1,
                                     heating=comp_background, **kw)
    if thermal:
        flow, therm, _ = diagnose(solver, X[:3 * n], N, lmax, m, symm, ricb, thermal=1, heating=heating,
                                  gradient_series=gradient_series, **kw)
    else:
        flow, therm = flow_c.copy(), np.zeros_like(comp)
        flow[:, :, 4] = 0.0
    flow[:, :, 5] = flow_c[:, :, 4]
    return flow, therm, comp, degrees


def diagnose_magnetic_energy(solver, Xb, N, lmax, m, bsymm, ricb, rcmb=1.0, Ra=None, Rb=None):
    """Magnetic energy and magnetic diffusion of the induced field, degree by degree: columns 0 and 1 of the
    `bdgn` of utils4pp.diagnose (magnetic_worker, utils4pp.py:491-533).  Their integrands are the kinetic ones
    (`energy_pol/tor`, `diffus_pol/tor`) applied to the poloidal / toroidal scalars [f | g] of b with the
    degrees of `bsymm = symm * symmB0` (utils.py:56), so it is the flow pass of `kb_diagnose` on the magnetic
    block `Xb = X[2n:4n]` of a solution.  Column 2 (the induction power) is NaN: it needs the coupling of u
    with the background field (induction4pp, utils4pp.py:678-775, written for a quadrupolar field only -- on
    the degree-1 fields the assembly covers the reference itself returns 0 there), as does the Lorentz power
    of the momentum balance.  Returns (mag[nsol, n_l, 3], degrees of b)."""
    flow, _, degrees = diagnose(solver, Xb, N, lmax, m, bsymm, ricb, thermal=0, rcmb=rcmb, Ra=Ra, Rb=Rb)
    mag = np.full(flow.shape[:2] + (3,), np.nan)
    mag[:, :, :2] = flow[:, :, :2]
    return mag, degrees


def differential_gradient_factor(ricb, rcmb=1.0):
    """ricb / (rcmb - ricb): the constant of the 'differential' background gradient that the advection
    operators carry (operators.py:736, 802) and utils4pp.thermal_advect leaves out (utils4pp.py:409, 419);
    `power_balance(advect_scale_thm=..., advect_scale_cmp=...)` takes it when the balance is to close."""
    return ricb / (rcmb - ricb)


def power_balance(flow, therm, degrees, lam, Ek, ViscosD=None, Beyonce=0.0, ThermaD=0.0, pss=0.0, comp=None,
                  CompBuoy=0.0, CompD=0.0, advect_scale_thm=1.0, advect_scale_cmp=1.0):
    """The sums and residuals of spin_doctor.py:148-242 for ONE solution: `flow` [n_l, 6], `therm`
    [nb, 3], `lam` its eigenvalue (growth rate = real part; 0 for a forced solution).

    spin_doctor.py scales the dissipation and power terms with `par.OmgTau`, which current
    parameters.py files no longer define; with the time scale those files do define
    (`Gaspard = 1`, viscous factor `ViscosD`, buoyancy factor `Beyonce`, thermal factor `ThermaD`,
    parameters.py:273-277) the balances read
        resid0: Dint + Dkin - pss = 0            (internal vs kinetic dissipation)
        resid1: 2 sigma KE - ViscosD Dkin + Beyonce Wthm = 0
        resid3: 2 sigma TE - ThermaD Dthm - Wadv = 0
    each divided by its largest term.  With a compositional field (`comp` = `cdgn` [nb, 3];
    spin_doctor.py:155, 183-186): `CompBuoy` = OmgTau^2 BV2_comp (operators.py:423) enters resid1 with
    column 5 of `flow`, `CompD` = OmgTau Ek / Schmidt (operators.py:824) scales the dissipation, and
        resid4: 2 sigma CE - CompD Dcmp - Wadv_cmp = 0
    is the compositional twin of resid3 (the reference prints CE, Dcmp, Wadv_cmp and forms no residual
    of them).  `advect_scale_*` multiply the advection integrals before the residuals (default 1: what
    the reference computes; see `differential_gradient_factor`)."""
    lp, lt, ll = degrees
    ll = np.asarray(ll)
    ViscosD = Ek if ViscosD is None else ViscosD
    sigma = complex(lam).real
    KE, Dkin0, Dint0, _, Wthm0, Wcmp0 = flow.sum(axis=0)
    out = {"KE": KE, "KP": flow[np.isin(ll, lp), 0].sum(), "KT": flow[np.isin(ll, lt), 0].sum(),
           "Dkin": ViscosD * Dkin0, "Dint": ViscosD * Dint0, "Wthm": Beyonce * Wthm0, "Wcmp": CompBuoy * Wcmp0}
    out["resid0"] = abs(Dint0 + Dkin0 - pss) / max(abs(Dint0), abs(Dkin0), abs(pss))
    out["resid1"] = abs(2 * sigma * KE - out["Dkin"] + out["Wthm"] + out["Wcmp"]) / max(
        abs(2 * sigma * KE), abs(out["Dkin"]), abs(out["Wthm"]), abs(out["Wcmp"]))
    if therm is not None and therm.size and np.any(therm):
        TE, Dthm0, Wadv = therm.sum(axis=0)
        Wadv = advect_scale_thm * Wadv
        out.update(TE=TE, Dthm=ThermaD * Dthm0, Wadv_thm=Wadv)
        out["resid3"] = abs(2 * sigma * TE - out["Dthm"] - Wadv) / max(abs(2 * sigma * TE), abs(out["Dthm"]), abs(Wadv))
    if comp is not None and comp.size and np.any(comp):
        CE, Dcmp0, Wadv = comp.sum(axis=0)
        Wadv = advect_scale_cmp * Wadv
        out.update(CE=CE, Dcmp=CompD * Dcmp0, Wadv_cmp=Wadv)
        out["resid4"] = abs(2 * sigma * CE - out["Dcmp"] - Wadv) / max(abs(2 * sigma * CE), abs(out["Dcmp"]), abs(Wadv))
    return out


def viscous_torques(X, N, lmax, m, symm, ricb, Ek, rcmb=1.0):
    """(vtorq[nsol], vtorq_icb[nsol]): the viscous torques on the mantle and on the inner core as
    spin_doctor.py:164-166 forms them, `Ek * gamma . u` with the row vectors of utils.py:1288-1399 for
    SPHERICAL boundaries (the only call the reference makes is `gamma_visc(0, 0, 0)`).

    Only the degree-1 toroidal scalar T carries a torque through a sphere of radius R:
    (8 pi / 3) R^2 (R T'(R) - T(R)), axial for m = 0 (equatorially symmetric flow; mantle and inner core),
    equatorial with a further sqrt(2) for m = 1 (antisymmetric flow; mantle only in the reference).  Boundary
    values of the Chebyshev series: T_k(+-1) = (+-1)^k, dT_k/dr(+-1) = (+-1)^(k+1) k^2 * 2 / (rcmb - r0)
    (utils.py:1261-1285)."""
    X = np.asarray(X, dtype=np.complex128)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    zero = np.zeros(X.shape[1], dtype=complex)
    _, lt, _ = _chain.ell(m, lmax, symm)
    if not ((m == 0 and symm == 1) or (m == 1 and len(lt) and lt[0] == 1)):
        return zero, zero.copy()
    N1 = N if ricb > 0 else N // 2
    n = N1 * ((lmax - m + 1) // 2)
    # Chebyshev orders held by the degree-1 toroidal block (every other one in a full sphere, utils4pp.py:67-107)
    s = (symm + 1) // 2
    k = np.arange(N, dtype=float) if ricb > 0 else (m + s) % 2 + 2.0 * np.arange(N1)
    r0 = ricb if ricb > 0 else -rcmb
    scale = 2.0 / (rcmb - r0)
    T1 = X[n:n + N1]  # first toroidal degree is l = 1 in both cases

    def gamma(x, R):
        return (8 * np.pi / 3) * R ** 2 * (R * x ** (k + 1) * k ** 2 * scale - x ** k)

    vt = Ek * (np.sqrt(2.0) if m == 1 else 1.0) * (gamma(1.0, 1.0) @ T1)
    vi = Ek * (gamma(-1.0, ricb) @ T1) if (m == 0 and ricb > 0) else zero
    return vt, vi


def spin_doctor_tables(flow, therm, comp, degrees, lam, Ek, OmgTau, BV2=0.0, BV2_comp=0.0, Etherm=0.0, Ecomp=0.0,
                       vtorq=None, vtorq_icb=None, pss=0.0):
    """The rows spin_doctor.py:370-396 appends to flow.dat, thermal.dat and compositional.dat, one per
    solution, from the per-degree integrals of `diagnose` / `diagnose_double_diffusive` -- with the
    reference's OWN scalings (`par.OmgTau`, `par.BV2`, `par.Etherm`, `par.Ecomp`, spin_doctor.py:158-186) and
    its own residuals (:227-242: resid1 without the compositional power), unlike `power_balance`.
    flow.dat: KE KP KT Dkin Dint Wlor Wthm Wcmp resid0 resid1 Re/Im vtorq Re/Im vtorq_icb (Wlor = 0: no
    magnetic runs here); thermal.dat: TE Wadv_thm Dthm resid3; compositional.dat: CE Wadv_cmp Dcmp."""
    lp, lt, ll = degrees
    ll = np.asarray(ll)
    nsol = flow.shape[0]
    lam = np.asarray(lam, dtype=complex).reshape(-1)
    vtorq = np.zeros(nsol, dtype=complex) if vtorq is None else vtorq
    vtorq_icb = np.zeros(nsol, dtype=complex) if vtorq_icb is None else vtorq_icb
    out = {"flow": np.zeros((nsol, 14))}
    if therm is not None:
        out["thermal"] = np.zeros((nsol, 4))
    if comp is not None:
        out["compositional"] = np.zeros((nsol, 3))
    for i in range(nsol):
        sigma = lam[i].real
        KE, Dkin0, Dint0, _, Wthm0, Wcmp0 = flow[i].sum(axis=0)
        KP, KT = flow[i][np.isin(ll, lp), 0].sum(), flow[i][np.isin(ll, lt), 0].sum()
        Dkin, Dint = OmgTau * Ek * Dkin0, OmgTau * Ek * Dint0
        Wlor, Wthm, Wcmp = 0.0, OmgTau ** 2 * BV2 * Wthm0, OmgTau ** 2 * BV2_comp * Wcmp0
        resid0 = abs(Dint0 + Dkin0 - pss) / max(abs(Dint0), abs(Dkin0), abs(pss)) if Ek != 0 else np.nan
        resid1 = abs(2 * sigma * KE - Dkin - Wlor + Wthm) / max(abs(2 * sigma * KE), abs(Dkin), abs(Wlor), abs(Wthm))
        out["flow"][i] = [KE, KP, KT, Dkin, Dint, Wlor, Wthm, Wcmp, resid0, resid1, vtorq[i].real, vtorq[i].imag,
                          vtorq_icb[i].real, vtorq_icb[i].imag]
        if therm is not None:
            TE, Dthm0, Wadv = therm[i].sum(axis=0)
            Dthm = Dthm0 * Etherm
            out["thermal"][i] = [TE, Wadv, Dthm, abs(2 * sigma * TE - Dthm - Wadv) / max(abs(2 * sigma * TE), abs(Dthm), abs(Wadv))]
        if comp is not None:
            CE, Dcmp0, Wadv = comp[i].sum(axis=0)
            out["compositional"][i] = [CE, Wadv, Dcmp0 * Ecomp]
    return out
