#!/usr/bin/env python3
import sys, os, ctypes as C
os.environ["KB_SWEEP_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic
P, b = int(sys.argv[1]), int(sys.argv[2])
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
s = lib.Solver(0)
s.set_option(lib.OPT_REFINE, 0)
s.set_pencil(A, B); s.set_chain(perm, nodeptr); s.factor(1j)
rhs = B @ synthetic.start_vector(A.shape[0], 3)
for i in range(3):
    x = s.solve(rhs)
print("solve_ms", s.stats()["solve_ms"])
L = lib.load()
out = np.zeros(256 * 8, dtype=np.int64)
L.kb_dbg_sweep_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
g = L.kb_dbg_sweep_timing(s.h, out.ctypes.data, 256)
t = out[: g * 8].reshape(g, 8)
steps = P
names = ["pre", "collect", "couple", "stage wait", "gemv+publish", "shadow", "-", "-"] if not os.environ.get("KB_NO_ONEHOP") else ["pre", "phaseA(poll y)", "poll t", "sync after shfl", "final+store", "gemv j-loop", "shuffles", "-"]
print("per-step cycles (mean over CTAs; CTA0; max):")
for k in range(8):
    print("  %-15s mean %8.0f  cta0 %8.0f  max %8.0f" % (names[k], t[:, k].mean() / steps, t[0, k] / steps, t[:, k].max() / steps))
print("  total per step %.0f cycles" % (t.sum(axis=1).mean() / steps))
