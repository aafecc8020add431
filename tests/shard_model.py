"""Host (numpy) model of the l-sharded factor/solve: the chain is cut into G contiguous
segments; the last node of every segment but the final one is a SEPARATOR; each rank
eliminates its interior (block Thomas + spikes towards the separator above it) and the
(G-1)-node reduced interface system is solved redundantly.

Test infrastructure for the host-side sharding logic and the reference semantics of
kore_b200/csrc/kb_shard.cu; it is NOT a product path (dense numpy, no GPU)."""
import numpy as np

from kore_b200 import chain


class Segment:
    """What ONE rank computes."""

    def __init__(self, Tp, nodeptr, ranges, g):
        G = len(ranges)
        lo, hi = ranges[g]
        self.g, self.G = g, G
        self.nodeptr = nodeptr
        self.top = lo - 1 if g > 0 else None            # separator above (node index)
        self.bot = hi - 1 if g < G - 1 else None        # separator below
        self.interior = list(range(lo, hi - 1 if g < G - 1 else hi))
        blk = lambda p, q: Tp[nodeptr[p]:nodeptr[p + 1], nodeptr[q]:nodeptr[q + 1]].toarray()
        self.blk = blk
        self.M, self.V, self.Gs = {}, {}, {}
        s = self.interior[0]
        S = blk(s, s)
        F = blk(s, self.top) if self.top is not None else None   # block (p, t)
        Gm = blk(self.top, s) if self.top is not None else None  # block (t, p)
        self.acc = None
        for p in self.interior:
            M = np.linalg.inv(S)
            self.M[p] = M
            if self.top is not None:
                V = M @ F
                H = Gm @ M
                self.V[p], self.Gs[p] = V, Gm
                self.acc = Gm @ V if self.acc is None else self.acc + Gm @ V
            nxt = p + 1
            if nxt in self.interior:
                W = M @ blk(p, nxt)
                S = blk(nxt, nxt) - blk(nxt, p) @ W
                if self.top is not None:
                    F = -blk(nxt, p) @ V
                    Gm = -H @ blk(p, nxt)
        # contributions to the reduced system
        e = self.interior[-1]
        self.R_above = self.C_sub = self.C_sup = None
        if self.bot is not None:
            W = self.M[e] @ blk(e, self.bot)
            self.R_above = blk(self.bot, self.bot) - blk(self.bot, e) @ W
            if self.top is not None:
                self.C_sub = -blk(self.bot, e) @ self.V[e]          # block (bot, top)
                self.C_sup = -(self.Gs[e] @ self.M[e]) @ blk(e, self.bot)  # block (top, bot)

    # ---- solve phases
    def forward(self, r):
        npt = self.nodeptr
        self.y = {}
        prev = None
        for p in self.interior:
            c = r[npt[p]:npt[p + 1]].copy()
            if prev is not None:
                c -= self.blk(p, prev) @ self.y[prev]
            self.y[p] = self.M[p] @ c
            prev = p
        a_top = None
        if self.top is not None:
            a_top = sum(self.Gs[p] @ self.y[p] for p in self.interior)
        b_bot = None
        if self.bot is not None:
            e = self.interior[-1]
            b_bot = self.blk(self.bot, e) @ self.y[e]
        return a_top, b_bot

    def backward(self, xsep):
        """xsep: dict separator node -> solution.  Returns dict interior node -> solution."""
        x = {}
        nxt_val, nxt = (xsep[self.bot], self.bot) if self.bot is not None else (None, None)
        for p in reversed(self.interior):
            v = self.y[p].copy()
            if nxt is not None:
                v -= self.M[p] @ (self.blk(p, nxt) @ nxt_val)
            if self.top is not None:
                v -= self.V[p] @ xsep[self.top]
            x[p] = v
            nxt, nxt_val = p, v
        return x


def reduced_solve(segs, r, nodeptr):
    """Assemble and solve the separator system from every rank's contributions (what each
    rank does redundantly after the all-gather)."""
    G = len(segs)
    seps = [segs[g].bot for g in range(G - 1)]
    R, Csub, Csup, rho = [], [None] * (G - 1), [None] * (G - 1), []
    fw = [s.forward(r) for s in segs]
    for j in range(G - 1):
        Rj = segs[j].R_above - segs[j + 1].acc
        R.append(Rj)
        rj = r[nodeptr[seps[j]]:nodeptr[seps[j] + 1]] - fw[j][1] - fw[j + 1][0]
        rho.append(rj)
        if j >= 1:
            Csub[j] = segs[j].C_sub     # block (sep_j, sep_{j-1})
            Csup[j] = segs[j].C_sup     # block (sep_{j-1}, sep_j)
    Mr, z = [], []
    for j in range(G - 1):
        S = R[j] if j == 0 else R[j] - Csub[j] @ Mr[j - 1] @ Csup[j]
        Mr.append(np.linalg.inv(S))
        c = rho[j] if j == 0 else rho[j] - Csub[j] @ z[j - 1]
        z.append(Mr[j] @ c)
    xs = [None] * (G - 1)
    for j in range(G - 2, -1, -1):
        xs[j] = z[j] if j == G - 2 else z[j] - Mr[j] @ (Csup[j + 1] @ xs[j + 1])
    return {seps[j]: xs[j] for j in range(G - 1)}


def sharded_solve(Tp, nodeptr, G, r):
    P = len(nodeptr) - 1
    ranges = chain.split_ranges(P, G)
    segs = [Segment(Tp, nodeptr, ranges, g) for g in range(G)]
    if G > 1:
        xsep = reduced_solve(segs, r, nodeptr)
    else:
        segs[0].forward(r)
        xsep = {}
    x = np.zeros_like(r)
    for p, v in xsep.items():
        x[nodeptr[p]:nodeptr[p + 1]] = v
    for s in segs:
        for p, v in s.backward(xsep).items():
            x[nodeptr[p]:nodeptr[p + 1]] = v
    return x
