"""CPU, world_size 2 over gloo: the host-side logic of the l-sharded path -- node ranges,
separators, what each rank contributes to the reduced interface system and what comes back
(tests/shard_model.py is the numpy statement of kore_b200/csrc/kb_shard.cu)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, q, fast=False):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import shard_model as sm
    from kore_b200 import chain
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = load_case(name)
    T = (c.A - c.tau * c.B).tocsr()
    Tp = T[c.perm][:, c.perm].tocsr()
    r = c.oracle["solve_rhs"][c.perm]
    P = len(c.nodeptr) - 1
    ranges = chain.split_ranges(P, world)
    seg = (sm.FastSegment if fast else sm.Segment)(Tp, c.nodeptr, ranges, rank)   # this rank's elimination only
    a_top, b_bot = seg.forward(r)
    mine = dict(R_above=seg.R_above, acc=seg.acc, C_sub=seg.C_sub, C_sup=seg.C_sup, a_top=a_top, b_bot=b_bot,
                bot=seg.bot)
    allc = [None] * world
    dist.all_gather_object(allc, mine)                     # the one exchange step
    # reduced system, redundantly on every rank
    G = world
    seps = [allc[j]["bot"] for j in range(G - 1)]
    Mr, z = [], []
    for j in range(G - 1):
        S = allc[j]["R_above"] - allc[j + 1]["acc"]
        rho = r[c.nodeptr[seps[j]]:c.nodeptr[seps[j] + 1]] - allc[j]["b_bot"] - allc[j + 1]["a_top"]
        if j > 0:
            S = S - allc[j]["C_sub"] @ Mr[j - 1] @ allc[j]["C_sup"]
            rho = rho - allc[j]["C_sub"] @ z[j - 1]
        Mr.append(np.linalg.inv(S))
        z.append(Mr[j] @ rho)
    xs = [None] * (G - 1)
    for j in range(G - 2, -1, -1):
        xs[j] = z[j] if j == G - 2 else z[j] - Mr[j] @ (allc[j + 1]["C_sup"] @ xs[j + 1])
    xsep = {seps[j]: xs[j] for j in range(G - 1)}
    part = seg.backward(xsep)
    pieces = [None] * world
    dist.all_gather_object(pieces, part)
    x = np.zeros_like(r)
    for p, v in xsep.items():
        x[c.nodeptr[p]:c.nodeptr[p + 1]] = v
    for d in pieces:
        for p, v in d.items():
            x[c.nodeptr[p]:c.nodeptr[p + 1]] = v
    xo = c.oracle["solve_x"][c.perm]
    q.put((rank, float(np.linalg.norm(x - xo) / np.linalg.norm(xo)), sorted(part.keys())[:1], ranges))
    dist.destroy_process_group()


@pytest.mark.parametrize("name,fast", [("m0_small", False), ("magnetic_small", False), ("magnetic_small", True),
                                       ("dormy", True)])
def test_two_rank_sharded_solve_over_gloo(name, fast):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q, fast)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rel, first, ranges in out:
        assert rel < 1e-9, (rank, rel)
    # the two ranks own disjoint interiors that start at their range's first node
    firsts = {rank: first[0] for rank, rel, first, ranges in out}
    ranges = out[0][3]
    assert firsts[0] == ranges[0][0] and firsts[1] == ranges[1][0]


def test_single_process_model_matches_oracle_for_many_ranks():
    import shard_model as sm
    c = load_case("m0_small")
    T = (c.A - c.tau * c.B).tocsr()
    Tp = T[c.perm][:, c.perm].tocsr()
    r = c.oracle["solve_rhs"][c.perm]
    xo = c.oracle["solve_x"][c.perm]
    for G in (1, 2, 4, 8):
        x = sm.sharded_solve(Tp, c.nodeptr, G, r)
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)


@pytest.mark.parametrize("name", ["m0_small", "dormy"])
def test_fast_path_model_matches_oracle_and_general_contributions(name):
    """The fast path's algebra (interior as a chain of its own, corner blocks of its inverse from
    identity-column chains, two sweep passes) gives the same reduced-system blocks as the spike
    recurrences of the general path and the oracle's solution, for every rank count; short
    interiors (one to three nodes: one-sided, no product chains) included."""
    import shard_model as sm
    from kore_b200 import chain
    c = load_case(name)
    T = (c.A - c.tau * c.B).tocsr()
    Tp = T[c.perm][:, c.perm].tocsr()
    r = c.oracle["solve_rhs"][c.perm]
    xo = c.oracle["solve_x"][c.perm]
    P = len(c.nodeptr) - 1
    for G in (2, 3, 8, P // 2):
        x = sm.sharded_solve(Tp, c.nodeptr, G, r, fast=True)
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo), G
    ranges = chain.split_ranges(P, 3)
    a, b = sm.Segment(Tp, c.nodeptr, ranges, 1), sm.FastSegment(Tp, c.nodeptr, ranges, 1)
    for name_ in ("R_above", "acc", "C_sub", "C_sup"):
        u, v = getattr(a, name_), getattr(b, name_)
        assert np.linalg.norm(u - v) <= 1e-9 * max(np.linalg.norm(u), 1e-300), name_
