#!/usr/bin/env python3
"""Correctness and throughput of the batched complex128 product kernel of the l-sharded fast
path (kb_zgemm_batch, kore_b200/csrc/kb_shard.cu) against numpy:  python tools/dev_zgemm.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kore_b200 import lib  # noqa: E402


def run(s, m, n, k, transA, batch, reps, beta=0.5, alpha=-1.0):
    rng = np.random.default_rng(m + 7 * n + 13 * k + transA)
    A = (rng.standard_normal((k, m) if transA else (m, k)) + 1j * rng.standard_normal((k, m) if transA else (m, k)))
    B = rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))
    C0 = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    Cc = np.ascontiguousarray(C0.copy())
    ms = C.c_double(0.0)
    f = s.lib.kb_dbg_zgemm
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                  C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    # reps = 0: one launch with the given beta, so the result can be checked
    rc = f(s.h, m, n, k, transA, A.ctypes.data, B.ctypes.data, Cc.ctypes.data, alpha, beta, batch, 0, C.byref(ms))
    assert rc == 0, (rc, s.lib.kb_last_error(s.h))
    ref = alpha * ((A.T if transA else A) @ B) + beta * C0
    err = float(np.abs(Cc - ref).max() / np.abs(ref).max())
    t = None
    if reps > 0:
        rc = f(s.h, m, n, k, transA, A.ctypes.data, B.ctypes.data, Cc.ctypes.data, alpha, beta, batch, reps, C.byref(ms))
        assert rc == 0, (rc, s.lib.kb_last_error(s.h))
        t = ms.value
    return err, t


def main():
    s = lib.Solver(0)
    out = []
    for (m, n, k, tr) in [(600, 600, 600, 0), (600, 600, 600, 1), (148, 74, 148, 0), (74, 148, 74, 1), (37, 5, 91, 0),
                          (1072, 1072, 1072, 0), (428, 428, 428, 0)]:
        for batch in (1, 2, 4):
            reps = 20 if m >= 400 else 0
            err, t = run(s, m, n, k, tr, batch, reps)
            rec = {"m": m, "n": n, "k": k, "transA": tr, "batch": batch, "max_rel_err": err}
            if t:
                rec["ms_per_launch"] = t
                rec["tflops"] = 8.0 * m * n * k * batch / (t * 1e-3) / 1e12
            out.append(rec)
            print(json.dumps(rec), flush=True)
            assert err < 1e-13, rec
    s.close()


if __name__ == "__main__":
    main()
