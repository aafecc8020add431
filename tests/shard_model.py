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


class FastSegment:
    """What ONE rank computes on the fast path (factor_sharded_fast / sharded_sweeps_fast of
    kb_shard.cu): the interior is a block-tridiagonal system of its own, factored two-sided and
    solved by the folded sweep (tests/fold_model.py); the corner blocks of its inverse come from the
    folded sweep with identity-block right-hand sides, written as chains of dense products; a solve
    is two passes of the folded sweep around the reduced solve.  Same contributions as Segment."""

    def __init__(self, Tp, nodeptr, ranges, g):
        import fold_model as fm
        G = len(ranges)
        lo, hi = ranges[g]
        self.g, self.G, self.nodeptr = g, G, nodeptr
        self.top = lo - 1 if g > 0 else None
        self.bot = hi - 1 if g < G - 1 else None
        self.interior = list(range(lo, hi - 1 if g < G - 1 else hi))
        blk = lambda p, q: Tp[nodeptr[p]:nodeptr[p + 1], nodeptr[q]:nodeptr[q + 1]].toarray()
        self.blk = blk
        I = self.interior
        n = len(I)
        D = [blk(p, p) for p in I]
        L = [blk(p, p - 1) if k > 0 else None for k, p in enumerate(I)]
        U = [blk(p, p + 1) if k < n - 1 else None for k, p in enumerate(I)]
        self.mid = n // 2 if n >= 4 else n - 1          # kb_shard.cu: two-sided from four nodes on
        m = self.mid
        self.Ms = fm.two_sided_factor(D, L, U, m)
        self.FL, self.FU = fm.fold(self.Ms, L, U)
        self.M = {p: self.Ms[k] for k, p in enumerate(I)}
        FL, FU, Ms = self.FL, self.FU, self.Ms
        bf, bl = D[0].shape[0], D[-1].shape[0]
        # forward chains of the two identity column blocks
        Tf = {0: np.eye(bf, dtype=complex)}
        for k in range(m):
            Tf[k + 1] = -FL[k] @ Tf[k]
        Tl = {n - 1: np.eye(bl, dtype=complex)}
        for k in range(n - 1, m, -1):
            Tl[k - 1] = -FU[k] @ Tl[k]
        # backward: U_m = T_m; up with base Tf (column f) / zero (column l), down the other way round
        Uf, Ul = {m: Tf[m]}, {m: Tl[m]}
        for k in range(m, 0, -1):
            Uf[k - 1] = Tf[k - 1] - FU[k] @ Uf[k]
            Ul[k - 1] = -FU[k] @ Ul[k]
        for k in range(m, n - 1):
            Ul[k + 1] = Tl[k + 1] - FL[k] @ Ul[k]
            Uf[k + 1] = -FL[k] @ Uf[k]
        Gff, Glf = Ms[0] @ Uf[0], Ms[n - 1] @ Uf[n - 1]
        Gfl, Gll = Ms[0] @ Ul[0], Ms[n - 1] @ Ul[n - 1]
        self.corners = (Gff, Glf, Gfl, Gll)
        f, l = I[0], I[-1]
        self.R_above = self.acc = self.C_sub = self.C_sup = None
        if self.bot is not None:
            self.R_above = blk(self.bot, self.bot) - blk(self.bot, l) @ Gll @ blk(l, self.bot)
        if self.top is not None:
            self.acc = blk(self.top, f) @ Gff @ blk(f, self.top)
        if self.top is not None and self.bot is not None:
            self.C_sub = -blk(self.bot, l) @ Glf @ blk(f, self.top)
            self.C_sup = -blk(self.top, f) @ Gfl @ blk(l, self.bot)

    def _sweep(self, rhs):
        import fold_model as fm
        return fm.folded_sweep(self.Ms, self.FL, self.FU, rhs, self.mid)

    def forward(self, r):
        """Pass 1: the interior solve; what the separators need are its end values."""
        npt, I = self.nodeptr, self.interior
        self.r = [r[npt[p]:npt[p + 1]].copy() for p in I]
        y = self._sweep(self.r)
        a_top = self.blk(self.top, I[0]) @ y[0] if self.top is not None else None
        b_bot = self.blk(self.bot, I[-1]) @ y[-1] if self.bot is not None else None
        return a_top, b_bot

    def backward(self, xsep):
        """Pass 2: the interior solve with the separator values moved to the right-hand side."""
        I = self.interior
        rhs = [v.copy() for v in self.r]
        if self.top is not None:
            rhs[0] = rhs[0] - self.blk(I[0], self.top) @ xsep[self.top]
        if self.bot is not None:
            rhs[-1] = rhs[-1] - self.blk(I[-1], self.bot) @ xsep[self.bot]
        x = self._sweep(rhs)
        return {p: x[k] for k, p in enumerate(I)}


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


def sharded_solve(Tp, nodeptr, G, r, fast=False):
    P = len(nodeptr) - 1
    ranges = chain.split_ranges(P, G)
    cls = FastSegment if fast else Segment
    segs = [cls(Tp, nodeptr, ranges, g) for g in range(G)]
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
