#!/usr/bin/env python3
"""l-sharded parity check, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_sharded_check.py [case ...]
Every rank factors its segment of the chain, the reduced interface system goes over NCCL,
and the solve / eigenpairs are compared with the committed CPU-oracle values."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_case
    from kore_b200 import lib

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    names = sys.argv[1:] or ["spinover", "magnetic_small", "dormy"]
    ok = True
    for name in names:
        case = load_case(name)
        m = case.meta
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            assert lib.load().kb_nccl_unique_id(buf.ctypes.data) == 0
            uid = torch.from_numpy(buf)
        dist.broadcast(uid, 0)
        s = lib.Solver(local)
        s.set_pencil(case.A, case.B)
        s.set_chain(case.perm, case.nodeptr)
        s.set_sharding(rank, world, uid.numpy())
        s.factor(case.tau)
        x = s.solve(case.oracle["solve_rhs"])
        xo = case.oracle["solve_x"]
        rel = np.linalg.norm(x - xo) / np.linalg.norm(xo)
        T = (case.A - case.tau * case.B).tocsr()
        res = np.linalg.norm(T @ x - case.oracle["solve_rhs"]) / np.linalg.norm(case.oracle["solve_rhs"])
        lam, X, info = s.eigs(m["nev"], which=m["which_eigenpairs"], target=case.tau, tol=m["tol"], maxit=m["maxit"])
        d = max(np.min(np.abs(lam - lo)) / abs(lo) for lo in case.oracle["eig"])
        good = rel < 1e-9 and res < 1e-12 and d < 1e-9 and info["nconv"] >= m["nev"] and info["resid"].max() < 1e-9
        ok &= bool(good)
        path = {0: "one-gpu", 1: "general", 2: "fast"}[int(s.stats()["shard_path"])]
        if os.environ.get("KB_EXPECT_SHARD_PATH"):
            good = good and path == os.environ["KB_EXPECT_SHARD_PATH"]
        ok &= bool(good)
        print("rank %d/%d %-16s [%s] solve rel %.2e resid %.2e | eig max rel diff %.2e nconv %d maxres %.1e | factor %.1f ms eigs %.1f ms %s"
              % (rank, world, name, path, rel, res, d, info["nconv"], info["resid"].max(), info["factor_ms"],
                 info["eigs_ms"], "OK" if good else "FAIL"), flush=True)
        s.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
