#!/usr/bin/env python3
"""Developer probe (GPU box): cycle breakdown of the persistent strip factorisation."""
# cycle counters need the timing build: `make timing`, then KB_LIB_PATH=kore_b200/libkoreb200_timing.so
import sys, os, ctypes as C
os.environ["KB_SWEEP_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic
P, b = int(sys.argv[1]), int(sys.argv[2])
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
s = lib.Solver(0)
s.set_option(lib.OPT_REFINE, 0)
s.set_pencil(A, B); s.set_chain(perm, nodeptr)
for rep in range(2):
    s.factor(1j)
st = s.stats()
L = lib.load()
out = np.zeros(256 * 16, dtype=np.int64)
L.kb_dbg_factor_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
g = L.kb_dbg_factor_timing(s.h, out.ctypes.data, 148)
t = out[: g * 16].reshape(g, 16)
print("P=%d b=%d factor %.2f ms (%.2f TF/s executed)" % (P, b, st["factor_ms"], st["factor_flops"] / st["factor_ms"] / 1e9))
names = ["schur", "panel(rest)", "wait", "apply(fma)", "store+barrier", "S:poll", "S:barrier", "-", "P:isp scale", "P:barrier1", "P:resolve,publish,bar2",
         "P:fma,vote", "A:loads", "A:publish rows", "A:barrier", "-"]
nodes = (P + 1) // 2
for k in range(16):
    if names[k] != "-":
        print("  %-20s cycles per node: mean %9.0f  cta0 %9.0f  cta1 %9.0f max %9.0f" % (names[k], t[:, k].mean() / nodes, t[0, k] / nodes, t[1, k] / nodes, t[:, k].max() / nodes))
print("  total cycles per node %.0f" % (t.sum(axis=1).mean() / nodes))
rhs = B @ synthetic.start_vector(A.shape[0], 3)
x = s.solve(rhs)
T = (A - 1j * B).tocsr()
print("  residual %.2e" % (np.linalg.norm(T @ x - rhs) / np.linalg.norm(rhs)))
