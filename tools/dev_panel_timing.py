#!/usr/bin/env python3
import sys, os, ctypes as C
os.environ["KB_SWEEP_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic
P, b = int(sys.argv[1]), int(sys.argv[2])
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
s = lib.Solver(0)
s.set_pencil(A, B); s.set_chain(perm, nodeptr)
try:
    s.factor(1j)
except Exception as e:
    print('factor error (expected for ablations):', str(e)[:60])
L = lib.load()
out = np.zeros(256 * 8, dtype=np.int64)
L.kb_dbg_sweep_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.kb_dbg_sweep_timing(s.h, out.ctypes.data, 256)
n = out[3]
print("factor_ms", s.stats()["factor_ms"], "panels", n)
print("per panel cycles: load %.0f  loop %.0f (%.0f per column)  store %.0f" % (out[0] / n, out[1] / n, out[1] / n / 16, out[2] / n))
