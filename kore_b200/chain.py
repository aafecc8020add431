"""Chain (l-major) ordering of Kore's unknowns.

Kore orders unknowns section-major, ``[u | v | f | g | h | i]`` with ``nb``
blocks of ``N1`` radial coefficients per section
(/root/reference/bin/assemble.py:442,458,472,494,510,534; utils.py:26-36).
Re-ordering by spherical-harmonic degree ``l`` makes ``A - sigma B`` strictly
block-tridiagonal: node ``p`` holds every section block
that carries the p-th degree of ``ll = ut.ell(m,lmax,symm)[2]``.

Two builders:
  * :func:`chain_from_params`  -- from the Kore parameter set (the l lists of
    utils.py:174-183 and the section/parity rules of assemble.py:50-74);
  * :func:`chain_from_pattern` -- from the sparsity pattern alone (BFS level
    sets of the N1-block graph), used when no parameter metadata is available
    and as a cross-check: level sets of any connected graph are block
    tridiagonal by construction.

Both return ``(perm, nodeptr)``: ``perm[k]`` = original index of the k-th
unknown in chain order, ``nodeptr[p]:nodeptr[p+1]`` = rows of node ``p``.
Pure host logic (numpy); no GPU, no oracle.
"""
from __future__ import annotations

import numpy as np


def ell(m: int, lmax: int, vsymm: int):
    """The spherical-harmonic degrees of the problem, split by equatorial parity: (degrees of
    the first family, degrees of the second family, all degrees).  Same sets as the reference's
    index rule (utils.py:174-183): lmax - m + 1 consecutive degrees starting at m (at 1 for the
    axisymmetric case, where l = 0 carries nothing); the first family takes every other degree,
    beginning with the first one exactly when `m > 0` and `vsymm == 1` disagree."""
    axisymmetric = m == 0
    degrees = np.arange(lmax - m + 1, dtype=int) + (1 if axisymmetric else m)
    first = ((0 if axisymmetric else 1) + (1 if vsymm > 0 else 0)) % 2
    return degrees[first::2], degrees[1 - first::2], degrees


def section_degrees(m, lmax, symm, symmB0, hydro, magnetic, thermal, compositional):
    """Per-section degree list and base offset in units of n (assemble.py:50-74,
    442-534).  Returns list of (name, base_in_units_of_n, degrees)."""
    top, bot, _ = ell(m, lmax, symm)
    secs = []
    base = 0
    if hydro:
        secs.append(("u", base, top))
        secs.append(("v", base + 1, bot))
        base += 2
    if magnetic:
        # f follows ll_bot and g ll_top for an antisymmetric (dipole-like) B0,
        # swapped for a symmetric one (assemble.py:69-74; utils.py:47-55)
        if symmB0 == -1:
            secs.append(("f", base, bot))
            secs.append(("g", base + 1, top))
        else:
            secs.append(("f", base, top))
            secs.append(("g", base + 1, bot))
        base += 2
    if thermal:
        secs.append(("h", base, top))
        base += 1
    if compositional:
        secs.append(("i", base, top))
        base += 1
    return secs


def chain_from_params(N1, m, lmax, symm, symmB0=-1, hydro=1, magnetic=0, thermal=0,
                      compositional=0):
    """l-major permutation."""
    _, _, ll = ell(m, lmax, symm)
    secs = section_degrees(m, lmax, symm, symmB0, hydro, magnetic, thermal, compositional)
    nb = len(secs[0][2])
    n = N1 * nb
    perm = []
    nodeptr = [0]
    rad = np.arange(N1, dtype=np.int64)
    for l in ll:
        cnt = 0
        for (_, base, degs) in secs:
            k = np.nonzero(degs == l)[0]
            if k.size:
                perm.append(base * n + int(k[0]) * N1 + rad)
                cnt += N1
        if cnt:
            nodeptr.append(nodeptr[-1] + cnt)
    perm = np.concatenate(perm).astype(np.int64)
    return perm, np.asarray(nodeptr, dtype=np.int64)


def chain_from_pattern(indptr, indices, nrows, blk):
    """BFS level sets of the graph whose vertices are ``blk``-row blocks.

    The start block is a pseudo-peripheral vertex (two BFS sweeps).  Levels of
    a BFS are block tridiagonal by construction: an edge never spans more
    than one level."""
    nblk = nrows // blk
    assert nblk * blk == nrows
    rows = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(indptr))
    br = rows // blk
    bc = np.asarray(indices, dtype=np.int64) // blk
    key = np.unique(br * nblk + bc)
    gr, gc = key // nblk, key % nblk
    adj = [[] for _ in range(nblk)]
    for a, b in zip(gr, gc):
        if a != b:
            adj[a].append(b)
            adj[b].append(a)

    def bfs(start):
        level = -np.ones(nblk, dtype=np.int64)
        level[start] = 0
        frontier = [start]
        d = 0
        while frontier:
            nxt = []
            for v in frontier:
                for w in adj[v]:
                    if level[w] < 0:
                        level[w] = d + 1
                        nxt.append(w)
            frontier = nxt
            d += 1
        return level

    lev = bfs(0)
    if (lev < 0).any():
        raise ValueError("block graph is disconnected; cannot chain")
    far = int(np.argmax(lev))
    lev = bfs(far)
    far2 = int(np.argmax(lev))
    lev2 = bfs(far2)
    if lev2.max() > lev.max():
        lev = lev2
    order = np.lexsort((np.arange(nblk), lev))
    perm = (order[:, None] * blk + np.arange(blk)[None, :]).reshape(-1).astype(np.int64)
    counts = np.bincount(lev, minlength=int(lev.max()) + 1)
    nodeptr = np.concatenate([[0], np.cumsum(counts) * blk]).astype(np.int64)
    return perm, nodeptr


def check_block_tridiagonal(indptr, indices, perm, nodeptr):
    """True iff the permuted pattern only couples nodes p and p-1,p,p+1."""
    n = perm.size
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    node_of = np.searchsorted(nodeptr, np.arange(n), side="right") - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    pr = node_of[inv[rows]]
    pc = node_of[inv[np.asarray(indices, dtype=np.int64)]]
    return bool(np.all(np.abs(pr - pc) <= 1))


def split_ranges(nnodes, nparts):
    """Contiguous, near-equal split of the chain into ``nparts`` segments
    (multi-GPU l-sharding).  Returns list of (lo, hi)."""
    nparts = max(1, min(nparts, nnodes))
    base, rem = divmod(nnodes, nparts)
    out = []
    lo = 0
    for r in range(nparts):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out
