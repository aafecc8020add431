#!/usr/bin/env python3
"""BASELINE.json config 5 on the GPU box: a forcing-frequency sweep, one refactor + one solve of
(A - i omega B) x = f per frequency (kore_b200/sweep.py), frequencies dealt across the ranks.

  (a) real Kore matrices (tests/golden/forced_small*, assembled by the reference with forcing = 7,
      m = 2, symm = 1): the full 256-point sweep omega_k = -2 + 4k/255 of SURVEY.md 8(d), every
      16th solution checked against SciPy SuperLU on the same matrices, and the CPU time of that
      LU + solve beside it;
  (b) the E = 1e-8 size (synthetic P = b = 600 pencil of bench.py): `--big-points` frequencies.

Run single-process or under torch.distributed.run (one rank per GPU; gloo is enough, the sweep
has no data-path collective).  Writes gpurun_out/forced_sweep.json on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=256)
    ap.add_argument("--big-points", type=int, default=8)
    ap.add_argument("--P", type=int, default=600)
    ap.add_argument("--b", type=int, default=600)
    ap.add_argument("--handles", type=int, default=8,
                    help="library handles (streams + host threads) per GPU for the small-pencil sweep")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "forced_sweep.json"))
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)

    import scipy.sparse.linalg as ssl
    from conftest import load_case
    from kore_b200 import sweep, synthetic

    def gather_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    rec = {"world": world}
    # ---- (a) real Kore forced-libration matrices
    cf, ce = load_case("forced_small"), load_case("forced_small_eig")
    f = np.asarray(cf.bf.todense()).ravel().astype(np.complex128)
    omegas = -2.0 + 4.0 * np.arange(a.points) / max(1, a.points - 1)
    # untimed: context, lazy kernel loading, first allocations (hundreds of ms, once per process)
    sweep.forced_sweep(ce.A, ce.B, f, omegas[:2], ce.perm, ce.nodeptr, device=local)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    mine, X, times = sweep.forced_sweep(ce.A, ce.B, f, omegas, ce.perm, ce.nodeptr, device=local,
                                        rank=rank, world=world)
    wall = gather_max(time.perf_counter() - t0)
    # the same sweep with several handles per GPU (independent frequencies side by side)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    mine_h, X_h, times_h = sweep.forced_sweep(ce.A, ce.B, f, omegas, ce.perm, ce.nodeptr, device=local,
                                              rank=rank, world=world, handles=a.handles)
    wall_h = gather_max(time.perf_counter() - t0)
    worst_h = gather_max(float(np.max(np.linalg.norm(X_h - X, axis=0) / np.linalg.norm(X, axis=0))))
    worst, cpu_s, nchk = 0.0, 0.0, 0
    for k in range(0, len(mine), 16):
        T = (ce.A - 1j * mine[k] * ce.B).tocsc()
        t1 = time.perf_counter()
        xo = ssl.splu(T).solve(f)
        cpu_s += time.perf_counter() - t1
        nchk += 1
        worst = max(worst, float(np.linalg.norm(X[:, k] - xo) / np.linalg.norm(xo)))
    worst = gather_max(worst)
    rec["kore_forced_small"] = {
        "n": int(ce.n), "points": int(a.points), "wall_s": wall, "s_per_factor_solve_wall": wall / a.points * world,
        "points_per_s_all_ranks": a.points / wall,
        "factor_ms_mean": float(times[:, 0].mean()), "solve_ms_mean": float(times[:, 1].mean()),
        "max_rel_err_vs_superlu": worst, "checked_points_per_rank": nchk,
        "cpu_superlu_s_per_factor_solve": cpu_s / max(1, nchk),
        "handles": int(a.handles), "wall_s_with_handles": wall_h, "points_per_s_all_ranks_with_handles": a.points / wall_h,
        "max_rel_diff_handles_vs_one": worst_h,
    }
    assert worst < 1e-9, worst

    # ---- (b) E = 1e-8 size
    if a.big_points > 0:
        A, B, perm, nodeptr = synthetic.synthetic_pencil(a.P, a.b)
        rhs = B @ synthetic.start_vector(A.shape[0], 1)
        om = np.linspace(0.9, 1.1, a.big_points * world)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        mine, X, times = sweep.forced_sweep(A, B, rhs, om, perm, nodeptr, device=local, rank=rank, world=world)
        wall = gather_max(time.perf_counter() - t0)
        T = (A - 1j * mine[-1] * B).tocsr()
        res = float(np.linalg.norm(T @ X[:, -1] - rhs) / np.linalg.norm(rhs))
        rec["synthetic_P%d_b%d" % (a.P, a.b)] = {
            "n": int(A.shape[0]), "points": int(len(om)), "wall_s_incl_ingest": wall,
            "factor_ms_mean": float(times[:, 0].mean()), "solve_ms_mean": float(times[:, 1].mean()),
            "s_per_factor_solve_device": float(times.sum(axis=1).mean() / 1e3),
            "points_per_s_all_ranks_device": world / float(times.sum(axis=1).mean() / 1e3),
            "rel_residual_last": gather_max(res),
        }
        assert res < 1e-9, res
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(rec, open(a.out, "w"), indent=1)
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
