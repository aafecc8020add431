#!/usr/bin/env python3
"""Drop-in for Kore's bin/solve.py on a B200.

Same inputs, same outputs, same parameter semantics:

    cd <run dir with A.npz and B.npz | B_forced.npz, and bin/parameters.py>
    python -m kore_b200.solve $opts          # instead of: mpiexec -n k ./bin/solve.py $opts

Reads  : A.npz, B.npz (eigen problem) or B_forced.npz (forced problem)   solve.py:40, 65, 209
         bin/parameters.py  (forcing, nev, tol, maxit, which_eigenpairs, tau, hydro, magnetic,
                             thermal, compositional, m, lmax, N, symm, ricb, B0)
         argv: PETSc-style options (-st_type sinvert, -eps_*; solver-package options are
               accepted and ignored, see kore_b200/eps.py)
Writes : eigenvalues0.dat, real_/imag_{flow,magnetic,temperature,composition}.field,
         timing.dat (appended), no_conv_solution (when nothing converged)   solve.py:194-199, 275-311
The EPS / ST / KSP block (solve.py:91-123, 220-227) is replaced by libkoreb200 through the
facades in kore_b200/eps.py; everything else keeps the reference's file formats.

With `-kb_assemble` (or when there is no A.npz but the `*.mtx` radial operators of
bin/submatrices.py are present) the run skips bin/assemble.py as well: the pencil is assembled on
the GPU from the radial operators (kore_b200/assembly.py; hydrodynamic, Boussinesq thermal and
degree-1 magnetic and anelastic set-ups) -- the same matrices, bit for bit, without A.npz / B.npz ever being written or read.
When the `*.mtx` files are absent too (or with `-kb_operators`), bin/submatrices.py is skipped as
well: the radial operators come from parameters.py alone (kore_b200/radial.py).
`-kb_diagnose` adds power_balance.dat and spin_doctor.py's flow.dat / thermal.dat / compositional.dat
(kore_b200/diagnostics.py), `-kb_npz` adds eigenpairs.npz
(eigenvalues, the complex solution block and the row ranges of the fields, binary).
The forced right-hand side comes from B_forced.npz when it exists, else (forcing = 7, libration) it
is formed here.
"""
from __future__ import annotations

import os
import sys
from timeit import default_timer as timer

import numpy as np
import scipy.sparse as ss


def load_csr(filename):
    """bin/utils.py:1253-1256 (same .npz keys)."""
    z = np.load(filename)
    return ss.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))


def import_parameters(rundir):
    """`import parameters as par` exactly as the reference resolves it: bin/ of the run dir."""
    for d in (os.path.join(rundir, "bin"), rundir):
        if os.path.exists(os.path.join(d, "parameters.py")):
            sys.path.insert(0, d)
            break
    import parameters as par  # noqa: E402
    return par


def run_utils(rundir):
    """The run's own bin/utils.py (rcmb; for anelastic runs with stress-free boundaries the density slopes are
    evaluated through it and the run's radial_profiles.py, as assemble.py:1195-1198 does), or None when the
    directory holds a bare parameters.py."""
    for d in (os.path.join(rundir, "bin"), rundir):
        if os.path.exists(os.path.join(d, "parameters.py")):
            if not os.path.exists(os.path.join(d, "utils.py")):
                return None
            import importlib.util
            spec = importlib.util.spec_from_file_location("utils", os.path.join(d, "utils.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules.setdefault("utils", mod)     # radial_profiles.py / bc_variables.py do `import utils`
            spec.loader.exec_module(mod)
            return mod
    return None


def kore_sizes(par):
    """N1, n, sizmat, symmB0 as bin/utils.py:26-57 derives them from parameters."""
    N, ricb = int(par.N), par.ricb
    N1 = int(N / 2) * int(1 + np.sign(ricb)) + int((N % 2) * np.sign(ricb))
    n = int(N1 * (par.lmax - par.m + 1) / 2)
    sizmat = 2 * n * par.hydro + 2 * n * par.magnetic + n * par.thermal + n * par.compositional
    B0 = getattr(par, "B0", "axial")
    if B0 in ("axial", "dipole", "G21 dipole", "Luo_S1"):
        symmB0 = -1
    elif B0 == "Luo_S2":
        symmB0 = 1
    elif B0 == "FDM":
        symmB0 = int((-1) ** par.B0_l)
    else:
        symmB0 = -1
    return N1, n, sizmat, symmB0


def field_slices(par, n):
    """Row ranges of each field in the solution vector (solve.py:163-190, 243-261)."""
    out = []
    if par.hydro == 1:
        out.append(("flow", 0, 2 * n))
    if par.magnetic == 1:
        o = 2 * n * par.hydro
        out.append(("magnetic", o, o + 2 * n))
    if par.thermal == 1:
        o = 2 * n * par.hydro + 2 * n * par.magnetic
        out.append(("temperature", o, o + n))
    if par.compositional == 1:
        o = 2 * n * par.hydro + 2 * n * par.magnetic + n * par.thermal
        out.append(("composition", o, o + n))
    return out


def write_fields(par, n, vec):
    """One solution per column, np.savetxt default format (solve.py:283-306), written by the
    library's threaded formatter (kb_savetxt: the same bytes, ~20x faster than np.savetxt, which
    at the E = 1e-8 size would otherwise cost several times the solve itself)."""
    from . import lib as _lib
    for name, a, b in field_slices(par, n):
        _lib.savetxt("real_%s.field" % name, vec[a:b, :], "real")
        _lib.savetxt("imag_%s.field" % name, vec[a:b, :], "imag")


def write_power_balance(par, solver, vec, lam):
    """`-kb_diagnose`: the energy / dissipation / power integrals spin_doctor.py:119-242 derives from
    the field files, computed on the GPU straight from the solutions (kore_b200/diagnostics.py) and
    written to power_balance.dat, one row per solution:
    KE KP KT Dkin Dint Wthm resid0 resid1 [TE Dthm Wadv_thm resid3] [Wcmp CE Dcmp Wadv_cmp resid4]
    (the last group for runs with the composition equation, spin_doctor.py:155, 183-186)."""
    from . import diagnostics as dg
    if par.magnetic or not par.hydro or getattr(par, "anelastic", 0):
        print("-kb_diagnose covers hydrodynamic, Boussinesq thermal and double-diffusive runs only; skipped")
        return
    geom = (int(par.N), par.lmax, par.m, par.symm, par.ricb)
    heating = getattr(par, "heating", "differential")
    grad = {}
    if par.thermal and heating in ("two zone", "user defined"):
        # the run's own background gradient: the cd_ent table its radial operators are built from
        from . import assembly as _assembly, radial as _radial
        pp = _assembly.PhysicsParams.from_modules(par, run_utils(os.getcwd()))
        grad = dict(gradient_series=np.asarray(_radial.run_profiles(pp)["cd_ent"]).ravel())
    comp, extra = None, {}
    if par.compositional:
        flow, therm, comp, degs = dg.diagnose_double_diffusive(
            solver, vec, *geom, thermal=par.thermal, heating=heating,
            comp_background=getattr(par, "comp_background", "differential"), **grad)
        extra = dict(CompBuoy=par.OmgTau ** 2 * par.BV2_comp, CompD=par.OmgTau * par.Ek / par.Schmidt)
    else:
        flow, therm, degs = dg.diagnose(solver, vec, *geom, thermal=par.thermal, heating=heating, **grad)
    # power_balance.dat is this driver's own file: its thermal / compositional balances carry the constant of the
    # background gradient that the heat equation has and utils4pp.thermal_advect leaves out (ricb / gap for
    # 'differential', operators.py:736, 802; the minus of operators.py:738 for a gradient of the run's own), so
    # that resid3 / resid4 close; flow.dat / thermal.dat below keep the reference's numbers
    def scale(background):
        if background == "differential":
            return dg.differential_gradient_factor(par.ricb) if par.ricb > 0 else 1.0
        return -1.0 if background in ("two zone", "user defined") else 1.0
    extra.update(advect_scale_thm=scale(heating), advect_scale_cmp=scale(getattr(par, "comp_background", "internal")))
    rows = []
    for i in range(vec.shape[1]):
        pb = dg.power_balance(flow[i], therm[i] if par.thermal else None, degs, lam[i], par.Ek,
                              getattr(par, "ViscosD", par.Ek), getattr(par, "Beyonce", 0.0), getattr(par, "ThermaD", 0.0),
                              comp=None if comp is None else comp[i], **extra)
        keys = ["KE", "KP", "KT", "Dkin", "Dint", "Wthm", "resid0", "resid1"]
        if par.thermal:
            keys += ["TE", "Dthm", "Wadv_thm", "resid3"]
        if par.compositional:
            keys += ["Wcmp", "CE", "Dcmp", "Wadv_cmp", "resid4"]
        rows.append([pb[k] for k in keys])
    np.savetxt("power_balance.dat", np.asarray(rows))
    # the files spin_doctor.py itself leaves behind (flow.dat, thermal.dat, compositional.dat; appended to, as
    # there: spin_doctor.py:370-396), with the reference's own scalings -- which need par.OmgTau
    if getattr(par, "OmgTau", None) is None:
        print("-kb_diagnose: parameters.py does not define OmgTau (spin_doctor.py needs it): flow.dat / thermal.dat not written")
        return
    vt, vi = dg.viscous_torques(vec, *geom, par.Ek)
    tables = dg.spin_doctor_tables(flow, therm if par.thermal else None, comp, degs, lam, par.Ek, par.OmgTau,
                                   getattr(par, "BV2", 0.0), getattr(par, "BV2_comp", 0.0), getattr(par, "Etherm", 0.0),
                                   getattr(par, "Ecomp", 0.0), vt, vi)
    for name, rows in tables.items():
        with open(name + ".dat", "ab") as f:
            np.savetxt(f, rows)
    if par.forcing == 0:  # spin_doctor.py:400-402: the running list of eigenvalues next to the per-run eigenvalues0.dat
        with open("eigenvalues.dat", "ab") as f:
            np.savetxt(f, np.c_[np.real(lam), np.imag(lam)])


def main(argv=None, device=0):
    from . import eps as kb

    argv = list(sys.argv[1:] if argv is None else argv)
    tic = timer()
    rundir = os.getcwd()
    par = import_parameters(rundir)
    opts = kb.Options(argv)
    N1, n, sizmat, symmB0 = kore_sizes(par)

    import glob
    on_device = (opts.hasName("kb_assemble") or opts.hasName("kb_operators")
                 or (not os.path.exists("A.npz") and bool(glob.glob("*.mtx"))))
    A = asm_inputs = None
    if on_device:
        from . import assembly as _assembly
        pp = _assembly.PhysicsParams.from_modules(par, run_utils(rundir))
        pp.check_supported()
        if opts.hasName("kb_operators") or not glob.glob("*.mtx"):
            from . import radial as _radial
            operators = _radial.radial_operators(pp, radprofs=_radial.run_profiles(pp))
        else:
            operators = _assembly.load_operators(".")
        asm_inputs = (pp, operators)
        layout = kb.ChainLayout.from_params(N1, par.m, par.lmax, par.symm, symmB0, par.hydro, par.magnetic,
                                            par.thermal, par.compositional)
    else:
        A = load_csr("A.npz")
        if A.shape[0] != sizmat:
            raise SystemExit("A.npz is %d x %d but parameters.py implies sizmat = %d" % (A.shape + (sizmat,)))
        layout = kb.ChainLayout.from_kore(A, N1, par.m, par.lmax, par.symm, symmB0, par.hydro, par.magnetic,
                                          par.thermal, par.compositional)
    success = 0

    if par.forcing == 0:  # ------------------------------------------------ eigenvalue problem
        E = kb.EPS(device)
        E.create()
        if on_device:
            E.setAssembly(*asm_inputs)
        else:
            E.setOperators(A, load_csr("B.npz"))
        E.setChainLayout(layout)
        E.setProblemType(kb.EPS.ProblemType.GNHEP)
        E.setDimensions(par.nev)
        E.setTolerances(par.tol, par.maxit)
        if par.which_eigenpairs not in _WHICH:
            raise SystemExit("unknown which_eigenpairs %r" % (par.which_eigenpairs,))
        E.setWhichEigenpairs(par.which_eigenpairs)
        E.setTarget(par.tau)
        E.setFromOptions(opts)
        E.solve()

        nconv = E.getConverged()
        if nconv > 0:
            k = np.zeros((1, nconv), dtype=complex)
            vec = np.zeros((sizmat, nconv), dtype=complex)
            v = np.zeros(sizmat, dtype=complex)
            for i in range(nconv):
                k[0, i] = E.getEigenpair(i, v)
                vec[:, i] = v
            eigval = np.hstack([np.real(k).T, np.imag(k).T])
            kb.savetxt("eigenvalues0.dat", eigval)
            write_fields(par, n, vec)
            success = nconv
            if opts.hasName("kb_diagnose"):
                write_power_balance(par, E._solver, vec, k[0])
            if opts.hasName("kb_npz"):
                # binary twin of eigenvalues0.dat + the field files: one uncompressed .npz, no text round trip
                np.savez("eigenpairs.npz", eigenvalues=k[0], vectors=vec,
                         fields=np.array([[nm, a, b] for nm, a, b in field_slices(par, n)], dtype=object).astype(str))
        else:
            print("No converged solution found")
            np.savetxt("no_conv_solution", [0])
        st = E.getStats()
        E.destroy()
    else:  # ------------------------------------------------------------- forced problem
        if on_device and not os.path.exists("B_forced.npz"):
            bvec = _assembly.forcing_vector(asm_inputs[0])  # assemble.py:278-329 (libration)
        else:
            b0 = load_csr("B_forced.npz")
            bvec = np.asarray(b0[:, 0].todense()).ravel().astype(complex)
        x = np.zeros(sizmat, dtype=complex)
        K = kb.KSP(device)
        K.create()
        if on_device:
            K.setAssembly(*asm_inputs)
        else:
            K.setOperators(A)
        K.setChainLayout(layout)
        K.setTolerances(rtol=par.tol, max_it=par.maxit)
        K.setFromOptions(opts)
        K.solve(bvec, x)
        st = K.getStats()
        K.destroy()
        if not np.all(np.isfinite(x)):
            print("Solver crashed, got nan's!")
        else:
            success = 1
            print("Solution(s) computed")
            write_fields(par, n, x.reshape(-1, 1))

    toc = timer()
    print("Solve done in", toc - tic, "seconds",
          "(GPU: factor %.1f ms, eigs %.1f ms)" % (st.get("factor_ms", 0.0), st.get("eigs_ms", 0.0)))
    with open("timing.dat", "ab") as f:
        np.savetxt(f, np.array([toc - tic]))
    return 0 if success >= 0 else 1


_WHICH = ("LM", "SM", "LR", "SR", "LI", "SI", "TM", "TR", "TI")

if __name__ == "__main__":
    sys.exit(main())
