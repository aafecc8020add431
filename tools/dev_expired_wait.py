"""Dev: the folded sweep with unreadable tags (KB_OPT_INJECT_FAULT = 3): its waits really expire.
Run plain or under compute-sanitizer."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kore_b200 import lib, synthetic

P = int(sys.argv[1]) if len(sys.argv) > 1 else 40
b = int(sys.argv[2]) if len(sys.argv) > 2 else 600
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
s = lib.Solver(0)
s.set_option(lib.OPT_WAIT_MS, 50)
s.set_pencil(A, B)
s.set_chain(perm, nodeptr)
s.factor(1j)
r = B @ synthetic.start_vector(A.shape[0], 3)
x0 = s.solve(r)
print("healthy solve ok", s.stats()["protocol_fallbacks"], flush=True)
s.set_option(lib.OPT_INJECT_FAULT, 3)
t0 = time.perf_counter()
try:
    x1 = s.solve(r)
    print("after fault: %.3f s" % (time.perf_counter() - t0), s.stats()["protocol_fallbacks"], hex(s.stats()["wait_error"]),
          np.linalg.norm(x1 - x0) / np.linalg.norm(x0), flush=True)
except Exception as e:
    print("FAILED after %.3f s:" % (time.perf_counter() - t0), e, flush=True)
