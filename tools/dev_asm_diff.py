#!/usr/bin/env python3
"""Dev: where a device-assembled matrix differs from the NumPy model (tests/assembly_model.py)."""
import json, os, sys
import numpy as np, scipy.sparse as sp
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import assembly_model as am
from kore_b200 import assembly as asm, lib
for name in sys.argv[1:]:
    d = os.path.join(ROOT, "tests", "golden", name)
    pp = asm.PhysicsParams.from_dict(json.load(open(os.path.join(d, "asm_params.json"))))
    ops = asm.load_operators_npz(os.path.join(d, "operators.npz"))
    progs = [("A", asm.build_program_A(pp, ops))]
    if pp.forcing == 0:
        progs.append(("B", asm.build_program_B(pp, ops)))
    for which, prog in progs:
        for fs in (None, 0.37):
            q = prog if fs is None else prog.with_final_scale(fs)
            M = am.evaluate(q)
            with lib.Solver(0) as s:
                s.assemble(q if which == "A" else None, q if which == "B" else None)
                ip, ix, v = s.get_assembled(which)
            D = sp.csr_matrix((v, ix, ip), shape=M.shape)
            same_pat = np.array_equal(D.indptr, M.indptr) and np.array_equal(D.indices, M.indices)
            print(name, which, "fs", fs, "nnz", D.nnz, M.nnz, "pattern", same_pat, flush=True)
            if same_pat:
                bad = np.flatnonzero(D.data != M.data)
                print("   differing values:", bad.size)
                rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
                for k in bad[:8]:
                    r, c = rows[k], M.indices[k]
                    print("    row %d (block %d, i %d) col %d (block %d): dev %r model %r" %
                          (r, r // prog.N1, r % prog.N1, c, c // prog.N1, D.data[k], M.data[k]))
                if bad.size:
                    print("    block rows hit:", np.unique(rows[bad] // prog.N1)[:20], "block cols:", np.unique(M.indices[bad] // prog.N1)[:20])
            else:
                dr = np.flatnonzero(np.diff(D.indptr) != np.diff(M.indptr))
                print("   rows with different counts:", dr[:10], "of", dr.size)
