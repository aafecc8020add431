#!/usr/bin/env python3
"""Golden eigenvalues at the benchmark's OWN size: the synthetic pencil P = b = 600 (n = 360 000) of
bench.py solved by the CPU oracle (oracle/kore_oracle.py: block LU of the l-chain's fronts with LAPACK +
ARPACK on the explicit shift-invert operator, restating /root/reference/bin/solve.py:91-149).  A few
minutes on 8 cores.  Writes tests/golden/synthetic_P600_b600_eigs.json, which the GPU parity test
tests/test_gpu_parity.py::test_full_size_eigenvalues_against_the_cpu_oracle compares with.

    python tools/make_fullsize_golden.py [P b]"""
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
    from kore_b200 import synthetic
    P, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (600, 600)
    nev, ncv, tol, sigma = 10, 25, 1e-12, 1j
    A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
    v0 = synthetic.start_vector(A.shape[0], 1)
    t0 = time.perf_counter()
    op = ko.BlockShiftInvert(A, B, sigma, (perm, nodeptr))
    t_lu = time.perf_counter() - t0
    t0 = time.perf_counter()
    lam, X, info = ko.eigs(A, B, sigma, nev, "TM", ncv=ncv, tol=tol, v0=v0, op=op)
    t_eigs = time.perf_counter() - t0
    res = ko.residuals(A, B, lam, X)
    order = np.argsort(np.abs(lam - sigma))
    lam, res = lam[order], res[order]
    out = {
        "what": "CPU oracle (block LU + ARPACK) on kore_b200.synthetic.synthetic_pencil(%d, %d), sigma=1j, nev=10, "
                "ncv=25, tol=1e-12, eigenvalues sorted by distance from sigma" % (P, b),
        "P": P, "b": b, "n": int(A.shape[0]), "sigma": [0.0, 1.0], "nev": nev,
        "eigs": [[float(z.real), float(z.imag)] for z in lam],
        "oracle_residuals": [float(r) for r in res],
        "oracle_op_applies": int(info["napply"]), "factor_s": t_lu, "eigs_s": t_eigs, "host_cpus": os.cpu_count(),
    }
    path = os.path.join(ROOT, "tests", "golden", "synthetic_P%d_b%d_eigs.json" % (P, b))
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
