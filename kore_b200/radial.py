"""Radial operators of Kore's equations.

Replaces /root/reference/bin/submatrices.py:28-590 and the functions of bin/utils.py it calls
(chebco :251-270, Dlam :893-905, Slam :908-922, csl0 / csl / Mlam :925-1062, labelit :130-157,
decode_label :85-127, remroco :187-197; h0 .. h3 / chebco_h / B0_norm :555-890 for the background
fields; compute_profiles.py for the profile tables) for every set-up the reference itself can run
but the free-decay mode of degree 2: hydrodynamic, Boussinesq (differential,
internal or the run's own background gradient; with or without thermal diffusion) and anelastic
(also with a viscosity profile) runs, with or without the composition equation, shell or full sphere, viscous or the inviscid full sphere,
magnetic with an axial, dipole, G21, Luo_S1, Luo_S2 or l = 1 free-decay background field and any
conductivity profile.  Every operator ``r^X D^Y`` (and ``r^X h^(j)(r) D^Y``, ``r^X f(r) [g(r)] D^Y``
with the field's scalar h or the run's profiles f, g) of a section is the product

    (C^(Y) -> C^(g) basis change)  x  (multiplication by r^X [f(r)] in the C^(Y) basis)  x  D^Y

with g the Gegenbauer order of the section, cut to its parity for a full sphere, its last `chop`
rows dropped and `chop` empty boundary rows put on top.  The result feeds `assembly.assemble`
directly (a dict label -> CSR, what `assembly.load_operators` returns for a directory of the
reference's ``.mtx`` files) or is written as ``.mtx`` for the reference's own assemble.py.

What is different from the reference.  Its multiplication matrices are filled entry by entry by
an interpreted triple loop -- for every entry (j, k) of the band an O(k) product for the first
linearisation coefficient c_s^lambda(j, k) and an O(k) recursion for the others, although the
Chebyshev coefficients of r^X they multiply vanish beyond degree X (minutes beyond N = 676,
utils.py:1006-1026).  Here

  * the first coefficients come from four tables of running products (`_Tables`): the products of
    csl0 depend on (k), (|j - k|, k) and (j, |j - k|) only, so a table row is built once by the
    same left-to-right sequence of IEEE operations and read by every entry that needs it;
  * the recursion runs for all entries of the band at once (NumPy arrays over the entries) and
    stops after the last term whose Chebyshev coefficient is non-zero;
  * basis changes and derivatives are dense row operations (`_Upper`) whose terms are accumulated
    in the order scipy's CSR product of the reference accumulates them -- the STORAGE order of
    the left factor's rows, which is not ascending for a product of products.

Every floating-point operation of the reference is performed, on the same operands and in the same
order, so the operators are the reference's to the bit -- with one caveat the reference has itself:
its band entries are `numpy.dot` products, whose summation order belongs to the host BLAS.
``dot="blas"`` (default) makes the same call on the same vectors (same bits as the reference run on
the same machine: the committed fixtures in this container); ``dot="ordered"`` sums left to right,
machine independent, and differs from the former by an ulp in a few entries of the widest bands.
tests/test_radial.py compares with the operators the unmodified reference wrote for every fixture
under tests/golden/.
"""
import math
import os

import numpy as np

from .assembly import PhysicsParams

TOL = 1e-9  # submatrices.py:55


# ---------------------------------------------------------------------------
# Chebyshev coefficients of the radial functions
# ---------------------------------------------------------------------------
def _nodes(N, ricb, rcmb):
    """radii of the N Chebyshev-Gauss points (utils.py:257-263)"""
    i = np.arange(0, N)
    xi = np.cos(np.pi * (i + 0.5) / N)
    if ricb == 0:
        return rcmb * xi
    return ricb + (rcmb - ricb) * (xi + 1) / 2.


def _dct_coefficients(samples, N, tol):
    """first N Chebyshev coefficients of the function sampled on the Gauss points, small ones
    set to zero (utils.py:266-270)"""
    import scipy.fft as sfft
    out = sfft.dct(samples, type=2) / N
    out[0] = out[0] / 2.
    out[np.absolute(out) <= tol] = 0.
    return out


def chebco(powr, N, tol, ricb, rcmb):
    """Chebyshev coefficients of r**powr on [ricb, rcmb] (or [-rcmb, rcmb] without inner core)."""
    return _dct_coefficients(_nodes(N, ricb, rcmb) ** powr, N, tol)


def chebco_function(func, N, tol, ricb, rcmb, rpower=0):
    """Chebyshev coefficients of r**rpower * func(r) (utils.py:200-247)."""
    r = _nodes(N, ricb, rcmb)
    return _dct_coefficients(func(r) if rpower == 0 else r ** rpower * func(r), N, tol)


# ---------------------------------------------------------------------------
# basis changes and derivatives (dense, N x N)
# ---------------------------------------------------------------------------
def basis_change(lamb, N):
    """C^(lamb) -> C^(lamb+1) coefficients: diagonals 0 and +2 (utils.py:908-922)."""
    S = np.zeros((N, N))
    i = np.arange(N)
    if lamb == 0:
        d0 = 0.5 * np.ones(N)
        d0[0] = 1.
        d1 = -0.5 * np.ones(N - 2)
    else:
        t = np.arange(0., N)
        d0 = lamb / (lamb + t)
        d1 = -lamb / (lamb + t[2:])
    S[i, i] = d0
    S[i[:-2], i[:-2] + 2] = d1
    return S


def derivative_diagonal(lamb, N, ricb, rcmb):
    """The one diagonal (offset +lamb) of the order-lamb derivative, C^(0) -> C^(lamb)
    (utils.py:893-905): scale * (lamb + i), i = 0 .. N-lamb-1."""
    if ricb == 0:
        const1 = (1 / rcmb) ** lamb
    else:
        const1 = (2. / (rcmb - ricb)) ** lamb
    const2 = float(math.factorial(lamb - 1)) * 2 ** (lamb - 1.)
    return (const1 * const2) * (lamb + np.arange(0, N - lamb)).astype(float)


def times_derivative(M, lamb, N, ricb, rcmb):
    """M @ D^lamb: column k of the result is column k - lamb of M times the diagonal entry."""
    if lamb == 0:
        return M
    d = derivative_diagonal(lamb, N, ricb, rcmb)
    out = np.zeros_like(M)
    out[:, lamb:] = M[:, : N - lamb] * d[None, :]
    return out


# ---------------------------------------------------------------------------
# multiplication matrices
# ---------------------------------------------------------------------------
class _Tables:
    """Running products of csl0 (utils.py:925-941) for one Gegenbauer order and truncation.

    csl0(s, lamb, j, k) multiplies four products; in the two ways Mlam calls it (s = 0, or s = k)
    they are P[c] = prod_{t<c} (lamb+t)/(1+t), Q[d][c] = prod_{t<c} (d+1+t)/(d+lamb+t) and
    R[j][e] = prod_{t<e} (2 lamb+j+t)/(lamb+j+t), each advanced as ``p = p*num/float(den)``."""

    def __init__(self, lamb, N, reach):
        self.lamb, self.N, self.reach = lamb, N, reach
        P = np.ones(N + 1)
        for t in range(N):
            P[t + 1] = P[t] * (lamb + t) / float(1 + t)
        self.P = P
        d = np.arange(reach + 1, dtype=float)
        Q = np.ones((reach + 1, N + 1))
        for t in range(N):
            Q[:, t + 1] = Q[:, t] * (d + 1 + t) / (d + lamb + t)
        self.Q = Q
        j = np.arange(N, dtype=float)
        R = np.ones((N, reach + 1))
        for t in range(reach):
            R[:, t + 1] = R[:, t] * (2 * lamb + j + t) / (lamb + j + t)
        self.R = R


_table_cache = {}


def _tables(lamb, N, reach):
    key = (lamb, N)
    t = _table_cache.get(key)
    if t is None or t.reach < reach:
        t = _table_cache[key] = _Tables(lamb, N, reach)
    return t


def multiplication(a0, lamb, vector_parity, dot="blas"):
    """Matrix of the multiplication by the series sum_i a0[i] C_i^(lamb) in the C^(lamb) basis
    (utils.py:962-1062); `vector_parity` as there (0 with an inner core).  Dense N x N."""
    a0 = np.asarray(a0, dtype=float)
    N = a0.size
    if not np.sum(abs(a0)) > 0:
        return None
    nzi = np.nonzero(a0)[0]
    bw = int(nzi.max())

    idj = idk = 0
    if vector_parity != 0:
        rpower_parity = 1 - 2 * (int(nzi[-1]) % 2)
        lamb_parity = 1 - 2 * (lamb % 2)
        overall_parity = vector_parity * rpower_parity * lamb_parity
        idj = int((1 - overall_parity) / 2)
        idk = int((1 - vector_parity * lamb_parity) / 2)

    if lamb == 0:
        # Chebyshev basis: half of Toeplitz + Hankel (utils.py:1030-1049)
        a2 = np.copy(a0)
        a2[0] = 2 * a2[0]
        i = np.arange(N)
        T = a2[np.abs(i[:, None] - i[None, :])]
        ij = i[:, None] + i[None, :]
        H = np.where(ij < N, a2[np.minimum(ij, N - 1)], 0.0)
        H[0, :] = 0.0
        out = 0.5 * (T + H)
        if vector_parity != 0:
            out[int((1 + overall_parity) / 2)::2, :] = 0.0
            out[:, int((1 + vector_parity * lamb_parity) / 2)::2] = 0.0
        return out

    # entries of the band: |k - j| <= bw + 1, parities as asked
    step = 2 if vector_parity != 0 else 1
    jr = np.arange(idj, N, step, dtype=np.int64)
    off = np.arange(-bw - 1, bw + 2, dtype=np.int64)
    J = np.repeat(jr, off.size)
    K = J + np.tile(off, jr.size)
    keep = (K >= 0) & (K < N)
    if vector_parity != 0:
        keep &= (K % 2) == idk
    J, K = J[keep], K[keep]
    E = K - J
    D = np.abs(E)
    up = E > 0                                   # s0 = k - j > 0: csl(s, lamb, k, k - j)
    s0 = np.where(up, E, 0)
    nterm = np.minimum(np.where(D <= bw, (bw - D) // 2 + 1, 0), K - s0 + 1)
    L = int(nterm.max()) if nterm.size else 0

    tb = _tables(lamb, N, bw + 1)
    Jf, Kf = J.astype(float), K.astype(float)
    # first coefficient, csl0(s0, lamb, k, |j - k|): ((((p1*p2)*p3)*p4)*(j'+k'+lamb-2 s))/(j'+k'+lamb-s)
    lo = ~up
    c = np.empty(J.size)
    c[lo] = ((tb.P[K[lo]] * tb.Q[D[lo], K[lo]]) * (Jf[lo] + lamb)) / (Jf[lo] + lamb)
    c[up] = ((((tb.P[E[up]] * tb.P[J[up]]) * tb.R[J[up], E[up]]) * tb.Q[0, J[up]]) * (Jf[up] + lamb)) / (Kf[up] + lamb)

    C = np.zeros((J.size, max(L, 1)))
    C[:, 0] = c
    jj, kk, s = K.copy(), D.copy(), s0.copy()    # csl's j, k, s (utils.py:944-958), int64 as there
    for i in range(1, L):
        tmp1 = (jj + kk + lamb - s) * (lamb + s) * (jj - s) * (2 * lamb + jj + kk - s) * (kk - s + lamb)
        tmp2 = (jj + kk + lamb - s + 1) * (s + 1) * (lamb + jj - s - 1) * (lamb + jj + kk - s) * (kk - s + 1)
        live = i < nterm
        nxt = np.zeros(J.size)
        nxt[live] = C[live, i - 1] * tmp1[live].astype(float) / tmp2[live].astype(float)
        C[:, i] = nxt
        kk = kk + 2
        s = s + 1

    # the Chebyshev coefficients each term multiplies: a1[2 s + j - k] = a0[|j - k| + 2 i]
    a1 = np.zeros(2 * N + 2 * max(L, 1))
    a1[:N] = a0
    A = a1[D[:, None] + 2 * np.arange(max(L, 1))[None, :]]
    A[np.arange(max(L, 1))[None, :] >= nterm[:, None]] = 0.0

    out = np.zeros((N, N))
    if dot == "ordered":
        acc = np.zeros(J.size)
        for i in range(L):
            acc = acc + A[:, i] * C[:, i]
        out[J, K] = acc
    else:
        # numpy.dot on vectors of the reference's lengths (k + 1 - s0): the summation order of the
        # host BLAS depends on the length and on where in the vector a term sits
        full = K - s0 + 1
        za = np.zeros(N + 1)
        zc = np.zeros(N + 1)
        vals = np.empty(J.size)
        for e in range(J.size):
            n, f = int(nterm[e]), int(full[e])
            za[:n] = A[e, :n]
            zc[:n] = C[e, :n]
            vals[e] = np.dot(za[:f], zc[:f])
            za[:n] = 0.0
            zc[:n] = 0.0
        out[J, K] = vals
    return out


# ---------------------------------------------------------------------------
# products in scipy's accumulation order
# ---------------------------------------------------------------------------
class _Upper:
    """An upper-banded N x N matrix as scipy holds it after ``A * B``: dense values plus the order
    in which a row stores its diagonals.  The reference multiplies its basis changes with
    scipy.sparse, whose CSR product (sparsetools csr_matmat) accumulates an entry of the result
    over the left factor's row in STORAGE order and stores a result row in reverse order of first
    touch -- so S3*S2*S1*S0 is not summed in ascending column order, and the last bit shows it."""

    def __init__(self, val, order):
        self.val, self.order = val, list(order)

    def __matmul__(self, other):
        if isinstance(other, _Upper):
            touched = []
            for a in self.order:
                for b in other.order:
                    if a + b not in touched:
                        touched.append(a + b)
            return _Upper(self.apply(other.val), touched[::-1])
        return self.apply(other)

    def apply(self, M):
        """self @ M for a dense matrix or vector M, terms in storage order, sums from zero"""
        N = self.val.shape[0]
        out = np.zeros(M.shape)
        for a in self.order:
            g = np.diagonal(self.val, a)
            if M.ndim == 1:
                out[: N - a] += g * M[a:]
            else:
                out[: N - a, :] += g[:, None] * M[a:, :]
        return out


class GegenbauerBases:
    """The basis changes of submatrices.py:138-170: S[d] takes C^(0) coefficients to C^(d),
    G[g][g - d] takes C^(d) to C^(g); `None` stands for the identity."""

    def __init__(self, N):
        S0, S1, S2, S3 = (_Upper(basis_change(k, N), (0, 2)) for k in range(4))
        S10 = S1 @ S0
        S21 = S2 @ S1
        S210 = S2 @ S10
        S32 = S3 @ S2
        S321 = S32 @ S1
        S3210 = S321 @ S0
        self.S = [None, S0, S10, S210, S3210]
        self.G = [[None], [None, S0], [None, S1, S10], [None, S2, S21, S210], [None, S3, S32, S321, S3210]]


# ---------------------------------------------------------------------------
# which operators a run needs (submatrices.py:200-443)
# ---------------------------------------------------------------------------
def _labelit(labels, section, rplus=0):
    """utils.py:130-157: section suffix, power of r raised by rplus ('q1' is r**-1)"""
    out = []
    for lab in labels:
        if rplus > 0:
            old = lab[:2]
            power = -1 if old == "q1" else int(old[1])
            lab = "r" + str(power + rplus) + lab[2:]
        out.append(lab + "_" + section)
    return out


def operator_labels(pp: PhysicsParams):
    """(labels, vector parities) of the radial operators of a run, in the reference's order."""
    vP = int((1 - 2 * (pp.m % 2)) * pp.symm)
    vT = -vP
    labels, par = [], []
    dip = pp.dipole and pp.ricb > 0
    quadrupole = bool(pp.magnetic) and field_degree(pp) == 2
    vF = vP if quadrupole else -vP               # the induced field has the flow's parity times the background field's
    vG = -vF
    vS = vP                                      # entropy perturbation
    if pp.magnetic and (pp.B0 not in BACKGROUND_FIELDS or (pp.B0 == "FDM" and pp.B0_l != 1)
                        or (pp.B0 == "dipole" and pp.ricb <= 0)):
        raise NotImplementedError("B0 = %r%s" % (pp.B0, "" if pp.ricb > 0 else " without inner core"))
    if pp.hydro:
        u = ["r2_D0", "r3_D1", "r4_D2"] + ["r3_D0", "r4_D1"] + ["r0_D0", "r2_D2", "r3_D3", "r4_D4"]
        par += [vP, vP, vP, vT, vT, vP, vP, vP, vP]
        if pp.anelastic:
            u += ["r1_lho1_D0", "r2_lho2_D0", "r3_lho3_D0", "r2_lho1_D1", "r3_lho2_D1",
                  "r3_lho1_D2", "r4_lho2_D2", "r4_lho3_D1", "r4_lho1_D3"]
            par += [vP] * 9
            if pp.variable_viscosity:  # viscosity profile and its first two derivatives (submatrices.py:230-237)
                w = ["r0_vsc0_D0", "r1_vsc0_lho1_D0", "r1_vsc1_D0", "r2_vsc0_D2", "r2_vsc0_lho1_D1", "r2_vsc0_lho2_D0",
                     "r2_vsc1_lho1_D0", "r2_vsc1_D1", "r2_vsc2_D0", "r3_vsc0_D3", "r3_vsc0_lho1_D2", "r3_vsc0_lho2_D1",
                     "r3_vsc0_lho3_D0", "r3_vsc1_D2", "r3_vsc1_lho1_D1", "r3_vsc1_lho2_D0", "r3_vsc2_lho1_D0", "r4_vsc0_D4",
                     "r4_vsc0_lho1_D3", "r4_vsc0_lho2_D2", "r4_vsc0_lho3_D1", "r4_vsc1_D3", "r4_vsc1_lho1_D2",
                     "r4_vsc1_lho2_D1", "r4_vsc2_D2", "r4_vsc2_lho1_D1"]
                u += w
                par += [vP] * len(w)
        if pp.magnetic:
            u += ["r1_h0_D1", "r2_h1_D1", "r2_h0_D2", "r3_h1_D2", "r0_h0_D0", "r1_h1_D0", "r2_h2_D0",
                  "r3_h3_D0", "r3_h0_D3", "r1_h0_D0", "r2_h1_D0", "r2_h0_D1", "r3_h1_D1", "r3_h0_D2"]
            par += [vF] * 9 + [vG] * 5
            if quadrupole:
                u += ["r3_h2_D1", "r3_h2_D0"]
                par += [vF, vG]
        if pp.thermal or pp.compositional:
            u += ["r3_buo0_D0"] if pp.anelastic else ["r4_D0"]
            par += [vS if pp.anelastic else vP]
        labels += _labelit(u, "u", 2 * dip)
        v = ["r2_D0"] + ["r1_D0", "r2_D1"] + ["r0_D0", "r1_D1", "r2_D2"]
        par += [vT, vP, vP, vT, vT, vT]
        if pp.anelastic:
            v += ["r1_lho1_D0", "r2_lho2_D0", "r2_lho1_D1"]
            par += [vT] * 3
            if pp.variable_viscosity:
                w = ["r0_vsc0_D0", "r1_vsc0_D1", "r1_vsc0_lho1_D0", "r2_vsc0_D2", "r2_vsc0_lho1_D1", "r2_vsc1_D1",
                     "r2_vsc1_lho1_D0", "r2_vsc0_lho2_D0", "r1_vsc1_D0"]
                v += w
                par += [vT] * len(w)
        if pp.magnetic:
            v += ["r0_h0_D1", "r0_h1_D0", "r1_h2_D0", "r1_h0_D2", "r0_h0_D0", "r1_h1_D0", "r1_h0_D1"]
            par += [vF] * 4 + [vG] * 3
        labels += _labelit(v, "v", 3 * dip)
    if pp.magnetic:
        b0 = "r2_rho0_D0" if pp.anelastic else "r2_D0"
        dif = "eho" if pp.anelastic else "eta"
        f = [b0] + ["r0_h0_D0", "r1_h1_D0", "r1_h0_D1", "r1_h0_D0"] + ["r0_%s0_D0" % dif, "r1_%s0_D1" % dif, "r2_%s0_D2" % dif]
        par += [vF, vP, vP, vP, vT, vF, vF, vF]
        labels += _labelit(f, "f", 2 * dip)
        g = [b0] + ["r0_h0_D1", "r1_h1_D1", "q1_h0_D0", "r0_h1_D0", "r1_h2_D0", "r1_h0_D2", "r0_h0_D0", "r1_h0_D1", "r1_h1_D0"]
        par += [vG] + [vP] * 6 + [vT] * 3
        if pp.anelastic:
            g += ["r0_h0_lho1_D0", "r1_h1_lho1_D0", "r1_h0_lho1_D0", "r1_h0_lho1_D1"]
            par += [vP, vT, vT, vT]
            g += ["r0_eho0_D0", "r1_eho0_D1", "r2_eho0_D2", "r1_eta1_rho0_D0", "r2_eta1_rho0_D1"]
        else:
            g += ["r0_eta0_D0", "r1_eta0_D1", "r2_eta0_D2", "r1_eta1_D0", "r2_eta1_D1"]
        par += [vG] * 5
        labels += _labelit(g, "g", 3 * dip)
    if pp.thermal and pp.anelastic:
        h = ["r2_roT0_D0", "r1_tds0_D0"]
        par += [vS, vP]
        if pp.ThermaD > 0:
            h += ["r0_krT0_D0", "r1_krT0_D1", "r2_krT1_D1", "r2_krT0_D2"]
            par += [vS] * 4
        labels += _labelit(h, "h", 0)
    elif pp.thermal:
        if pp.heating == "differential":
            h = ["r0_D0", "r3_D0"]
            par += [vP, vP]
            if pp.ThermaD > 0:
                h += ["r1_D0", "r2_D1", "r3_D2"]
                par += [vP, vP, vP]
        elif pp.heating in ("internal", "two zone", "user defined"):
            h = ["r2_D0"]
            par += [vP]
            if pp.ThermaD > 0:
                h += ["r0_D0", "r1_D1", "r2_D2"]
                par += [vP, vP, vP]
            if pp.heating != "internal":
                h += ["r0_drS0_D0"]                   # r dT/dr of the run's own background (cd_ent)
                par += [vP]
        else:
            raise NotImplementedError("heating = %r" % (pp.heating,))
        labels += _labelit(h, "h", 0)
    if pp.compositional:
        if pp.comp_background == "differential":
            ci = ["r0_D0", "r1_D0", "r2_D1", "r3_D2", "r3_D0"]
        else:
            ci = ["r2_D0", "r0_D0", "r1_D1", "r2_D2"]
        par += [vP] * len(ci)
        labels += _labelit(ci, "i", 0)
    if pp.ricb > 0:
        par = [0] * len(labels)
    return labels, par


def gegenbauer_orders(pp: PhysicsParams):
    """Gegenbauer order of every section = number of boundary rows with an inner core
    (submatrices.py:167-179)."""
    g = {"u": 4, "v": 2, "f": 2, "g": 2, "h": 2, "i": 2}
    if pp.magnetic and "conductor" in pp.innercore:
        g["f"] = 3
    if (pp.Ek == 0 or pp.ViscosD == 0) and pp.ricb == 0:
        g["u"], g["v"] = 2, 1
    if pp.ThermaD == 0:
        g["h"] = 0
    return g


# ---------------------------------------------------------------------------
# background magnetic field (utils.py:555-797, 800-890)
# ---------------------------------------------------------------------------
def _field_axial(r, rp, d):
    return [lambda: (1 / 2) * r ** (1 + rp), lambda: (1 / 2) * r ** rp,
            lambda: np.zeros_like(r), lambda: np.zeros_like(r)][d]()


def _field_dipole(r, rp, d):
    return [lambda: (1 / 2) * r ** (-2 + rp), lambda: -r ** (-3 + rp),
            lambda: 3 * r ** (-4 + rp), lambda: -12 * r ** (-5 + rp)][d]()


def _field_g21(r, rp, d):
    return [lambda: (1 / 6) * r ** (1 + rp) - (1 / 10) * r ** (3 + rp), lambda: (1 / 6) * r ** rp - (3 / 10) * r ** (2 + rp),
            lambda: (6 / 10) * r ** (1 + rp), lambda: (6 / 10) * r ** rp][d]()


def _field_luo_s1(r, rp, d):
    return [lambda: (5 - 3 * r ** 2) * r ** (1 + rp), lambda: (5 - 9 * r ** 2) * r ** rp,
            lambda: -18 * r ** (1 + rp), lambda: -18 * r ** rp][d]()


def _field_luo_s2(r, rp, d):
    return [lambda: (157 - 296 * r ** 2 + 143 * r ** 4) * r ** (2 + rp), lambda: 2 * r ** (1 + rp) * (157 - 592 * r ** 2 + 429 * r ** 4),
            lambda: 2 * r ** rp * (157 - 1776 * r ** 2 + 2145 * r ** 4), lambda: 24 * r ** (1 + rp) * (-296 + 715 * r ** 2)][d]()


def _bessel_j(l, x, d):
    """spherical Bessel function of the first kind (d = 0) or its derivative (d = 1); three-term
    series below 1e-3 (utils.py:433-473)"""
    import scipy.special as scsp
    x = np.asarray(x, dtype=float)
    out = np.zeros_like(x)
    small = x < 1e-3
    if np.any(small):
        xs = x[small]
        c1 = (2 ** l) * scsp.factorial(l) / scsp.factorial(2 * l + 1)
        c2 = -(2 ** l) * scsp.factorial(l + 1) / scsp.factorial(2 * l + 3)
        c3 = (2 ** l) * scsp.factorial(l + 2) / (2 * scsp.factorial(2 * l + 5))
        if d == 0:
            out[small] = c1 * xs ** l + c2 * xs ** (l + 2) + c3 * xs ** (l + 4)
        else:
            out[small] = c1 * l * xs ** (l - 1) + c2 * (l + 2) * xs ** (l + 1) + c3 * (l + 4) * xs ** (l + 3)
    out[~small] = scsp.spherical_jn(l, x[~small], derivative=d)
    return out


def _bessel_j_series(l, x, d):
    """second (d = 2) and third (d = 3) derivative of j_l for small arguments, l = 1 (utils.py:450-461)"""
    import scipy.special as scsp
    c2 = -(2 ** l) * scsp.factorial(l + 1) / scsp.factorial(2 * l + 3)
    c3 = (2 ** l) * scsp.factorial(l + 2) / (2 * scsp.factorial(2 * l + 5))
    if d == 2:
        return c2 * (l + 2) * (l + 1) * x ** l + c3 * (l + 4) * (l + 3) * x ** (l + 2)
    return c2 * (l + 2) * (l + 1) * l * x ** (l - 1) + c3 * (l + 4) * (l + 3) * (l + 2) * x ** (l + 1)


def _bessel_y(l, x, d):
    import scipy.special as scsp
    return scsp.spherical_yn(l, x, derivative=d)


_beta_cache = {}


def decay_mode_wavenumber(beta0, l, ricb):
    """Radial wavenumber of the poloidal free-decay mode of degree l in a shell with insulating
    boundaries (Zhang & Fearn 1995; utils.py:530-550): the root next to the guess beta0."""
    key = (beta0, l, ricb)
    if key not in _beta_cache:
        import scipy.optimize as so

        def f0(x, ric, l):
            if ric == 0:
                return _bessel_j(l - 1, x, 0)            # Gubbins & Roberts (1987)
            return _bessel_j(l + 1, x * ric, 0) * _bessel_y(l - 1, x, 0) - _bessel_j(l - 1, x, 0) * _bessel_y(l + 1, x * ric, 0)
        _beta_cache[key] = so.root(f0, beta0, args=(ricb, l)).x[0]
    return _beta_cache[key]


def _field_decay_mode(r, rp, d, b):
    """free-decay mode of degree 1 with an inner core (utils.py:588-597, 640-646, 697-698, 772-776)"""
    j, y = _bessel_j, _bessel_y
    x = b * r
    y0b, j0b = y(0, b, 0), j(0, b, 0)
    if d == 0:
        return (j(1, x, 0) * y0b - j0b * y(1, x, 0)) * r ** rp
    if d == 1:
        return (b * (j(1, x, 1) * y0b - j0b * y(1, x, 1))) * r ** rp
    if d == 2:
        return ((x ** 2 * j(0, x, 1) + 2 * (j(1, x, 0) - x * j(1, x, 1))) * y0b
                - j0b * (x ** 2 * y(0, x, 1) + 2 * (y(1, x, 0) - b * r * y(1, x, 1)))) * r ** (-2 + rp)
    return (-2 * b ** 2 * r ** 2 * j(0, b * r, 1) * y0b - 8 * j(1, b * r, 0) * y0b + 8 * b * r * j(1, b * r, 1) * y0b
            - b ** 3 * r ** 3 * j(1, b * r, 1) * y0b + 2 * b ** 2 * r ** 2 * j0b * y(0, b * r, 1) + 8 * j0b * y(1, b * r, 0)
            - 8 * b * r * j0b * y(1, b * r, 1) + b ** 3 * r ** 3 * j0b * y(1, b * r, 1)) * r ** (-3 + rp)


def _field_decay_mode_sphere(r, rp, d, b):
    """free-decay mode of degree 1 of a full sphere (utils.py:586-587, 637-638, 685-691, 759-766)"""
    j = _bessel_j
    x = b * r
    if d == 0:
        return j(1, x, 0) * r ** rp
    if d == 1:
        return b * j(1, x, 1) * r ** rp
    k = x < 1e-3
    x0, x1 = x[k], x[~k]
    out = np.zeros_like(r)
    if d == 2:
        out[k] = b ** 2 * _bessel_j_series(1, x0, 2) * r[k] ** rp
        out[~k] = b ** 2 * j(0, x1, 1) * r[~k] ** rp + (2 * (j(1, x1, 0) - x1 * j(1, x1, 1))) * r[~k] ** (-2 + rp)
    else:
        out[k] = b ** 3 * _bessel_j_series(1, x0, 3) * r[k] ** rp
        out[~k] = (-2 * x1 ** 2 * j(0, x1, 1) - 8 * j(1, x1, 0) + x1 * (8 - x1 ** 2) * j(1, x1, 1)) * r[~k] ** (-3 + rp)
    return out


# r**rp times the d-th derivative of the poloidal scalar h(r) of the background field; degree l = 1 but for Luo_S2
BACKGROUND_FIELDS = {"axial": _field_axial, "dipole": _field_dipole, "G21 dipole": _field_g21, "Luo_S1": _field_luo_s1,
                     "FDM": _field_decay_mode, "Luo_S2": _field_luo_s2}


def field_degree(pp):
    """degree l of the background field (utils.py:45-55); its equatorial symmetry is (-1)**l"""
    return 2 if pp.B0 == "Luo_S2" else (pp.B0_l if pp.B0 == "FDM" else 1)


def background_field(r, kind, rp, d, pp=None):
    """h0 .. h3 of utils.py:555-797: r**rp times the d-th derivative of h.  Without inner core the
    Chebyshev nodes cover [-rcmb, rcmb]: the values on r < 0 are the mirror image with the parity
    of the function, (-1)**(l + rp) for even d and (-1)**(l - 1 + rp) for odd d."""
    r = np.asarray(r, dtype=float)
    pos = r > 0
    if kind == "FDM":
        if pp is None or pp.B0_l != 1:
            raise NotImplementedError("free-decay modes of degree l > 1")
        mode = _field_decay_mode if pp.ricb > 0 else _field_decay_mode_sphere
        vals = mode(r[pos], rp, d, decay_mode_wavenumber(pp.beta, pp.B0_l, pp.ricb))
    elif kind == "dipole" and (pp is None or pp.ricb <= 0):
        raise NotImplementedError("a dipole (singular at the origin) without inner core")
    else:
        vals = BACKGROUND_FIELDS[kind](r[pos], rp, d)
    out = np.zeros_like(r)
    out[pos] = vals
    neg = r < 0
    if pp is not None and pp.ricb == 0 and np.count_nonzero(pos) == np.count_nonzero(neg):
        out[neg] = np.flipud(vals) * (-1) ** (field_degree(pp) - (d % 2) + rp)
    return out


def field_normalisation(pp: PhysicsParams):
    """utils.py:837-890 (B0_norm)"""
    l = field_degree(pp)
    L = l * (l + 1)
    kind, ricb = pp.B0, pp.ricb
    if pp.cnorm == "rms_cmb":
        return (np.sqrt(2 * l + 1) / (l * (l + 1) * background_field(np.array([1.0]), kind, 0, 0, pp)))[0]
    if pp.cnorm in ("mag_energy", "Schmitt2012"):
        N = 240
        i = np.arange(0, N)
        xk = np.cos((i + 0.5) * np.pi / N)
        sqx = np.sqrt(1 - xk ** 2)
        rk = 0.5 * (1 - ricb) * (xk + 1) + ricb
        r2 = rk ** 2
        y0 = background_field(rk, kind, 0, 0, pp)
        y1 = background_field(rk, kind, 0, 1, pp)
        f0 = 4 * np.pi * L / (2 * l + 1)
        f1 = (L + 1) * y0 ** 2
        f2 = 2 * rk * y0 * y1
        f3 = r2 * y1 ** 2
        integ = (np.pi / N) * ((1 - ricb) / 2) * np.sum(sqx * f0 * (f1 + f2 + f3))
        return (1 if pp.cnorm == "mag_energy" else 2) / np.sqrt(integ)
    return pp.cnorm


def chebyshev_derivative(ck, ricb, rcmb):
    """Chebyshev coefficients of the r-derivative of a Chebyshev series (utils.py:273-294)"""
    c = np.copy(ck)
    c[0] = 2. * c[0]
    s = np.size(c)
    out = np.zeros_like(c)
    out[-2] = 2. * (s - 1.) * c[-1]
    for k in range(s - 3, -1, -1):
        out[k] = out[k + 2] + 2. * (k + 1) * ck[k + 1]
    out[0] = out[0] / 2.
    return out / rcmb if ricb == 0 else 2 * out / (rcmb - ricb)


def profile_table(func, order, N, ricb, rcmb, tol=TOL):
    """Columns: Chebyshev coefficients of func(r) and of its first `order` derivatives
    (utils.py:317-328, chebify)."""
    cols = [chebco_function(func, N, tol, ricb, rcmb)]
    for _ in range(order):
        cols.append(chebyshev_derivative(cols[-1], ricb, rcmb))
    return np.stack(cols, axis=1)


def profile_tables(pp: PhysicsParams, rap):
    """The tables compute_profiles.py:39-91 stores in radProfs.mat, from a module `rap` with the
    run's radial functions (radial_profiles.py: density, log_density, viscosity, krT, roT,
    log_temperature, kappa_rho, tds, buoFac, magnetic_diffusivity)."""
    N, ricb, rcmb = pp.N, pp.ricb, pp.rcmb
    t = {}
    if pp.anelastic:
        for key, name, order in (("cd_rho", "density", 2), ("cd_lho", "log_density", 4), ("cd_vsc", "viscosity", 2),
                                 ("cd_krT", "krT", 1), ("cd_roT", "roT", 0)):
            t[key] = profile_table(getattr(rap, name), order, N, ricb, rcmb)
        if pp.thermal:
            for key, name, order in (("cd_lnT", "log_temperature", 1), ("cd_kho", "kappa_rho", 1), ("cd_tds", "tds", 0),
                                     ("cd_buo", "buoFac", 0)):
                t[key] = profile_table(getattr(rap, name), order, N, ricb, rcmb)
    if pp.thermal and not pp.anelastic and pp.heating in ("two zone", "user defined"):
        f = rap.twozone if pp.heating == "two zone" else rap.BVprof
        t["cd_ent"] = chebco_function(lambda r: f(r, pp.args), N, TOL, ricb, rcmb, rpower=1).reshape([N, 1])
    if pp.magnetic:
        t["cd_eta"] = profile_table(rap.magnetic_diffusivity, 1, N, ricb, rcmb)
        if pp.anelastic:
            t["cd_eho"] = profile_table(rap.eta_rho, 1, N, ricb, rcmb)
    return t


def series_product(ck1, ck2, tol=TOL):
    """Chebyshev coefficients of the product of two Chebyshev series (utils.py:342-348)"""
    M = multiplication(ck1, 0, 0)
    out = np.zeros_like(ck2) if M is None else _rows_times_vector(M, ck2)
    out[np.absolute(out) <= tol] = 0.0
    return out


def _rows_times_vector_or_zero(ck1, ck2):
    """the inner product of utils.cheb3Product (:332-339), not thresholded: Mlam(ck1, 0, 0) * ck2"""
    M = multiplication(ck1, 0, 0)
    return np.zeros_like(ck2) if M is None else _rows_times_vector(M, ck2)


def _rows_times_vector(M, x):
    """M @ x as scipy's CSR product forms it: every row summed over its non-zero entries in
    ascending column order, from zero"""
    out = np.zeros(M.shape[0])
    for k in range(M.shape[1]):
        col = M[:, k]
        nz = col != 0.0
        out[nz] += col[nz] * x[k]
    return out


def _decode(label):
    """(index of the power of r, derivative order of h or None, [(profile, its derivative order) ...],
    derivative order of the operator, section) -- utils.py:85-127"""
    parts = label.split("_")
    rx = 6 if parts[0] == "q1" else int(parts[0][1])
    hx, prof = None, []
    for q in parts[1:-2]:
        if q[0] == "h" and len(q) == 2:
            hx = int(q[1])
        else:
            prof.append((q[:3], int(q[3])))
    return rx, hx, prof, int(parts[-2][1]), parts[-1]


RPOWERS = [0, 1, 2, 3, 4, 5, -1]  # submatrices.py:119: powers of r that go with h (index 6 is 1/r)


def radial_operators(pp: PhysicsParams, radprofs=None, dot="blas", dense=False):
    """Every radial operator of the run, label -> CSR (N1 x N1, boundary rows empty, the
    reference's ``<label>.mtx``), from the parameters alone.  `radprofs`: the tables of
    compute_profiles.py (``cd_eta``, ``cd_lho`` ... -- `profile_tables` makes them from a run's
    radial_profiles module); needed for anelastic runs, and for magnetic ones whose conductivity is
    not the uniform one radial_profiles.py ships with."""
    import scipy.sparse as sp
    N, ricb, rcmb = pp.N, pp.ricb, pp.rcmb
    labels, parities = operator_labels(pp)
    bases = GegenbauerBases(N)
    gorder = gegenbauer_orders(pp)
    radprofs = dict(radprofs or {})
    if pp.magnetic and "cd_eta" not in radprofs:
        # radial_profiles.py:252-262 as shipped: uniform conductivity
        radprofs["cd_eta"] = profile_table(lambda r: 1. / np.ones_like(r), 1, N, ricb, rcmb)
    if pp.anelastic and "cd_lho" not in radprofs:
        raise ValueError("anelastic = 1: pass radprofs=profile_tables(pp, <the run's radial_profiles module>)")
    if pp.thermal and not pp.anelastic and pp.heating in ("two zone", "user defined") and "cd_ent" not in radprofs:
        raise ValueError("heating = %r: pass radprofs=profile_tables(pp, <the run's radial_profiles module>)" % (pp.heating,))
    cnorm = field_normalisation(pp) if pp.magnetic else None
    rp, rdh = {}, {}

    def rpower(rx):
        if rx not in rp:
            rp[rx] = chebco(rx, N, TOL, ricb, rcmb)
        return rp[rx]

    def c0_series(rx, hx, prof):
        # submatrices.py:461-508: the series the multiplication matrix is made of
        cks = [radprofs["cd_ent" if name == "drS" else "cd_" + name][:, order] for name, order in prof]
        if hx is not None:
            if (rx, hx) not in rdh:
                rdh[rx, hx] = cnorm * _dct_coefficients(background_field(_nodes(N, ricb, rcmb), pp.B0, RPOWERS[rx], hx, pp), N, TOL)
            return series_product(rdh[rx, hx], cks[0]) if cks else rdh[rx, hx]
        if len(cks) == 2:
            return series_product(rpower(rx), _rows_times_vector_or_zero(cks[0], cks[1]))
        if len(cks) == 1:
            return cks[0] if rx == 0 else series_product(rpower(rx), cks[0])
        return rpower(rx)

    mult = {}
    out = {}
    for lab, vp in zip(labels, parities):
        rx, hx, prof, dx, sec = _decode(lab)
        key = (lab[:-2], vp)
        if key not in mult:
            c0 = c0_series(rx, hx, prof)
            mult[key] = multiplication(c0 if dx == 0 else bases.S[dx].apply(c0), dx, vp, dot=dot)
        M = mult[key]
        gb = gorder[sec]
        G = bases.G[gb][gb - dx]
        if M is None:
            # a vanishing series (second derivative of an axial field ...): the reference's 0 * D
            M = np.zeros((N, N))
        M = times_derivative(M if G is None else G.apply(M), dx, N, ricb, rcmb)
        if ricb == 0:
            # parity of the operator as a function of r (submatrices.py:548-567)
            if hx is not None:
                operator_parity = (-1) ** (hx + field_degree(pp) + RPOWERS[rx] + dx)  # h has the parity of its degree
            elif prof and prof[0][0] == "drS":
                operator_parity = 1                                         # the run's gradient is to make an even operator
            elif prof:
                if len(prof) > 1 or prof[0][0] not in ("eta", "roT", "krT"):
                    raise NotImplementedError("operators of the profiles %r without inner core" % (prof,))
                operator_parity = 1 - ((rx + prof[0][1] + dx) % 2) * 2      # even profiles
            else:
                operator_parity = 1 - ((rx + dx) % 2) * 2
            overall = vp * operator_parity
            M = M[int((1 - overall) / 2)::2, int((1 - vp) / 2)::2]
            chop = gb // 2
        else:
            chop = gb
        if chop > 0:
            M = np.vstack([np.zeros((chop, M.shape[1])), M[:-chop, :]])
        out[lab] = M if dense else sp.csr_matrix(M)
    return out


def resolution_rule(Ek, m, ncpus=4, g=1.0):
    """(N, lmax) Kore gives a run of Ekman number Ek by default (parameters.py:6-16, 296-301):
    N = max(48, even(int(17 Ek^-0.2))), lmax - m + 1 the multiple of 2 ncpus next below g N."""
    out = int(17 * Ek ** -0.2) if Ek != 0 else 48
    N = max(48, out + out % 2)
    return N, int(2 * ncpus * (np.floor_divide(g * N, 2 * ncpus)) + m - 1)


class OperatorCache:
    """Radial operators by what they depend on: the truncation, the radii, which equations are on,
    the background field and profiles -- and, without inner core only, the parities that go with
    (m, symm).  A sweep over m, symmetry, Rayleigh / Ekman factors or boundary conditions at fixed
    resolution generates them once; a ramp whose resolution follows `resolution_rule` generates
    them once per truncation (the reference re-runs submatrices.py at every step)."""

    def __init__(self, radprofs=None, dot="blas"):
        self.radprofs, self.dot, self.store = radprofs, dot, {}

    def key(self, pp):
        labels, parities = operator_labels(pp)
        field = (pp.B0, pp.B0_l, pp.beta, str(pp.cnorm)) if pp.magnetic else None
        return (pp.N, pp.ricb, pp.rcmb, tuple(labels), tuple(parities), field, tuple(sorted(gegenbauer_orders(pp).items())))

    def get(self, pp):
        k = self.key(pp)
        if k not in self.store:
            self.store[k] = radial_operators(pp, radprofs=self.radprofs, dot=self.dot)
        return self.store[k]


def run_profiles(pp: PhysicsParams):
    """The profile tables of the run in the current directory: from its radProfs.mat when
    compute_profiles.py has been run, else from its radial_profiles module (bin/ is on sys.path
    once the driver has imported `parameters`); None when the run needs none."""
    own_gradient = pp.thermal and pp.heating in ("two zone", "user defined")
    if not (pp.anelastic or pp.magnetic or own_gradient):
        return None
    if os.path.exists("radProfs.mat"):
        import scipy.io as sio
        return {k: v for k, v in sio.loadmat("radProfs.mat").items() if k.startswith("cd_")}
    try:
        import radial_profiles as rap
    except ImportError:
        if pp.anelastic or own_gradient:
            raise
        return None  # magnetic, Boussinesq: the uniform conductivity radial_profiles.py ships with
    return profile_tables(pp, rap)


def write_mtx(directory, operators):
    """``<label>.mtx`` files as submatrices.py:588 writes them (for the reference's assemble.py)."""
    import scipy.io as sio
    import scipy.sparse as sp
    for lab, M in operators.items():
        sio.mmwrite(os.path.join(directory, lab + ".mtx"), sp.csr_matrix(M))


def main(argv=None):
    """``python -m kore_b200.radial [ncpus]`` in a run directory: the ``.mtx`` files of
    ``./bin/submatrices.py ncpus`` (the argument is accepted and ignored: one core, a second or so)."""
    import sys
    from timeit import default_timer as timer
    from .solve import import_parameters
    tic = timer()
    par = import_parameters(os.getcwd())
    try:
        import utils as ut
    except ImportError:
        ut = None
    pp = PhysicsParams.from_modules(par, ut)
    print("N =", pp.N, ", lmax =", pp.lmax)
    ops = radial_operators(pp, radprofs=run_profiles(pp))
    write_mtx(".", ops)
    print("Submatrices generated and written to disk in", timer() - tic, "seconds")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
