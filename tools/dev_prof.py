#!/usr/bin/env python3
"""Small driver for ncu: synthetic chain PxB, one factor and NS solves (and optionally eigs)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic
P, b = int(sys.argv[1]), int(sys.argv[2])
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
s = lib.Solver(0)
s.set_option(lib.OPT_REFINE, 0)
s.set_pencil(A, B); s.set_chain(perm, nodeptr)
s.factor(1j)
rhs = B @ synthetic.start_vector(A.shape[0], 3)
for i in range(ns):
    x = s.solve(rhs)
print(s.stats())
