#!/usr/bin/env python3
"""Large real-Kore parity + baseline: a case assembled by the UNMODIFIED reference in the build
container (tools/make_case.py, shipped to the GPU box in the git-ignored bigcases/ directory) is
solved by libkoreb200 and by the CPU oracle on the box's own host cores, same nev/ncv/sigma.

    python tools/run_bigcase.py bigcases/E1e-6 [--skip-oracle]
Writes gpurun_out/bigcase_<name>.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import kore_oracle as ko
    from kore_b200 import chain, lib
    d = sys.argv[1]
    skip = "--skip-oracle" in sys.argv
    m = json.load(open(os.path.join(d, "meta.json")))
    A = ko.load_csr(os.path.join(d, "A.npz"))
    B = ko.load_csr(os.path.join(d, "B.npz"))
    n = A.shape[0]
    tau = complex(m["rtau"], m["itau"])
    nev = m["nev"]
    ncv = max(ko.default_ncv(nev), 24)
    perm, nodeptr = chain.chain_from_params(m["N1"], m["m"], m["lmax"], m["symm"], m["symmB0"], m["hydro"],
                                            m["magnetic"], m["thermal"], m["compositional"])
    out = dict(case=os.path.basename(d.rstrip("/")), n=n, nnzA=int(A.nnz), nnzB=int(B.nnz), nev=nev, ncv=ncv,
               P=len(nodeptr) - 1, b=int(np.diff(nodeptr).max()), host_cpus=os.cpu_count())
    rng = np.random.default_rng(1)
    v0 = rng.standard_normal(n) + 1j * rng.standard_normal(n)

    s = lib.Solver(0)
    t0 = time.perf_counter()
    s.set_pencil(A, B)
    s.set_chain(perm, nodeptr)
    s.factor(tau)
    lam, X, info = s.eigs(nev, m["which_eigenpairs"], target=tau, ncv=ncv, tol=1e-12, maxit=100, v0=v0)
    out["gpu_e2e_s"] = time.perf_counter() - t0
    s.factor(tau)
    lam, X, info = s.eigs(nev, m["which_eigenpairs"], target=tau, ncv=ncv, tol=1e-12, maxit=100, v0=v0)
    res = ko.residuals(A, B, lam, X)
    out.update(gpu_factor_ms=info["factor_ms"], gpu_eigs_ms=info["eigs_ms"], gpu_applies=int(info["op_applies"]),
               gpu_nconv=int(info["nconv"]), gpu_max_resid=float(res.max()),
               gpu_eigs=[[float(z.real), float(z.imag)] for z in lam[:nev]])
    rhs = B @ v0
    x = s.solve(rhs)
    T = (A - tau * B).tocsr()
    out["gpu_solve_resid"] = float(np.linalg.norm(T @ x - rhs) / np.linalg.norm(rhs))
    out["gpu_solve_ms"] = s.stats()["solve_ms"]
    s.close()
    print(json.dumps(out), flush=True)

    if not skip:
        t0 = time.perf_counter()
        op = ko.ShiftInvert(A, B, tau)
        out["cpu_lu_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        xo = op.solve(rhs)
        out["cpu_solve_s"] = time.perf_counter() - t0
        out["solve_rel_diff"] = float(np.linalg.norm(x - xo) / np.linalg.norm(xo))
        t0 = time.perf_counter()
        lam_o, X_o, info_o = ko.eigs(A, B, tau, nev, m["which_eigenpairs"],
                                     ncv=(ncv if m["which_eigenpairs"] == "TM" else None), tol=1e-12, v0=v0, op=op)
        out["cpu_eigs_s"] = time.perf_counter() - t0
        out["cpu_applies"] = int(info_o["napply"])
        out["cpu_max_resid"] = float(ko.residuals(A, B, lam_o, X_o).max())
        out["eig_max_rel_diff"] = float(max(np.min(np.abs(lam - lo)) / abs(lo) for lo in lam_o))
        out["speedup_factor_plus_eigs"] = (out["cpu_lu_s"] + out["cpu_eigs_s"]) / (
            (out["gpu_factor_ms"] + out["gpu_eigs_ms"]) / 1e3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bigcase_%s.json" % out["case"]), "w"), indent=1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
