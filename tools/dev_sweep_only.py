#!/usr/bin/env python3
"""Dev: device time of one chain sweep (kb_solve_dev, no refinement), CUDA events.
usage: [KB_LIB_PATH=...] dev_sweep_only.py P b [reps]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from kore_b200 import lib, synthetic
P, b = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
n = A.shape[0]
s = lib.Solver(0)
s.set_option(lib.OPT_REFINE, 0)
s.set_pencil(A, B); s.set_chain(perm, nodeptr); s.factor(1j)
rhs = torch.from_numpy(B @ synthetic.start_vector(n, 3)).cuda()
x = torch.empty_like(rhs)
st = torch.cuda.ExternalStream(s.stream())
with torch.cuda.stream(st):
    for _ in range(3):
        s.solve_dev(rhs.data_ptr(), x.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        s.solve_dev(rhs.data_ptr(), x.data_ptr())
    e1.record(st)
torch.cuda.synchronize()
print("%s: %.3f ms per kb_solve_dev (sweep + permutations/scalings)" % (os.environ.get("KB_LIB_PATH", "default"), e0.elapsed_time(e1) / reps))
