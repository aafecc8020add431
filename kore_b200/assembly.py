"""Device-side assembly of Kore's pencil.

The reference builds ``A`` and ``B`` on the host (/root/reference/bin/assemble.py:432-1171):
for every spherical-harmonic degree ``l`` and every pair of coupled sections it forms a sparse
N1 x N1 block as a short linear combination of banded radial operators read from ``*.mtx``
(written by bin/submatrices.py:533-590), appends dense boundary-condition rows, concatenates
COO triplets, converts to CSR and divides by ``||B||_F``: tens of seconds of interpreted scipy
per assembly (45 s at E = 1e-7), then ~250 MB from disk through the host to the solver.

Here the host only writes down WHAT each block is -- an *assembly program*: a table of blocks,
each a few groups of ``(coefficient, radial operator)`` terms with their scalar factors -- and
the library evaluates it on the GPU straight into the CSR the layout build consumes
(csrc/kb_assemble.cu, `kb_assemble`).  The program is a few hundred KB whatever the truncation;
re-assembly at another Rayleigh number or forcing frequency is a new coefficient table and one
kernel pass.

Bit compatibility.  Every entry of the reference's matrices is the result of a fixed sequence
of IEEE double operations (scipy.sparse scalar products and sums evaluated left to right,
no fused multiply-add).  The program records that sequence, not just the mathematical
coefficient: a group is ``s_k * (... s_1 * ((c_0 x_0 + c_1 x_1) + c_2 x_2 ...))`` and the
evaluators (the kernel, and the NumPy model in tests/assembly_model.py) perform exactly these
operations, so the assembled CSR equals the reference's bit for bit (tests/test_assembly.py).

Scope: hydrodynamic (sections u, v), Boussinesq thermal (section h) and -- to rounding, not to the
bit, see `_magnetic_blocks` -- magnetic (sections f, g; any degree-1 background field, insulating
boundaries) problems, viscous,
with or without inner core, eigenvalue (forcing = 0) and forced runs (forcing = 7, 9, 10: the modes
that work in the reference; A and the forcing vector) -- BASELINE.json configs 1 to 5 -- and anelastic (density-stratified) runs, bit for bit (with a viscosity profile: its two viscous blocks to rounding).  Quadrupolar
background fields and conducting inner cores
raise NotImplementedError (their pencils still enter through `kb_set_pencil`).

The physics restated here (which operators enter which block with which coefficient):
momentum equation operators.py:22-195, buoyancy :386-405, heat equation :699-775; block
placement assemble.py:432-760, 1012-1075; boundary rows assemble.py:1188-1345 with the
Chebyshev end-point tables of bc_variables.py:8-32.
"""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass, field

import numpy as np

from .chain import section_degrees

MAX_SCALES = 4
MAX_BAND = 1023  # widest operator band (2 H + 1) the assembly kernel accepts (ka_check, csrc/kb_assemble.cu)


@dataclass
class PhysicsParams:
    """The entries of Kore's ``parameters`` / ``utils`` modules the assembly depends on."""
    hydro: int = 1
    magnetic: int = 0
    thermal: int = 0
    compositional: int = 0
    anelastic: int = 0
    variable_viscosity: int = 0
    # anelastic runs with stress-free boundaries: d(ln rho)/dr at the inner / outer boundary (assemble.py:1195-1198;
    # `from_modules` evaluates them with the run's own radial_profiles.py when it is importable)
    lho1_icb: float = None
    lho1_cmb: float = None
    m: int = 1
    lmax: int = 8
    N: int = 8
    symm: int = -1
    ricb: float = 0.35
    rcmb: float = 1.0
    Ek: float = 1e-3
    bci: int = 1
    bco: int = 1
    bci_thermal: int = 0
    bco_thermal: int = 0
    heating: str = "differential"
    args: object = None           # [rc, h, rsy] of the 'two zone' temperature gradient (parameters.py:203-206)
    forcing: int = 0
    forcing_frequency: float = 0.0
    forcing_amplitude_cmb: float = 0.0
    forcing_amplitude_icb: float = 0.0
    Gaspard: float = 1.0
    ViscosD: float = 1e-3
    Beyonce: float = 0.0
    ThermaD: float = 0.0
    # compositional runs (parameters.py:228-271; OmgTau is commented out in every shipped file and has to be set)
    comp_background: str = "internal"
    OmgTau: float = None
    BV2_comp: float = 0.0
    Schmidt: float = 1.0
    bci_compositional: int = 1
    bco_compositional: int = 1
    # magnetic runs (parameters.py:103-137, 275-278)
    B0: str = "axial"
    B0_l: int = 1                 # degree of the free-decay-mode field (parameters.py:113)
    beta: float = 3.0             # guess for its radial wavenumber (parameters.py:112)
    innercore: str = "insulator"
    mantle: str = "insulator"
    Hendrik: float = 0.0
    MagnetD: float = 0.0
    # thin conducting layer on the fluid side of a boundary ('TWA'; parameters.py:115-131)
    mu: float = 1.0
    c_icb: float = 0.0
    c1_icb: float = 0.0
    c_cmb: float = 0.0
    c1_cmb: float = 0.0
    cnorm: object = "mag_energy"  # normalisation of the background field (parameters.py:144-155; radial.py only)

    @classmethod
    def from_modules(cls, par, ut=None):
        """From an imported Kore ``parameters`` module (and ``utils`` for rcmb)."""
        kw = {}
        for f in cls.__dataclass_fields__:
            if hasattr(par, f):
                kw[f] = getattr(par, f)
        if ut is not None and hasattr(ut, "rcmb"):
            kw["rcmb"] = ut.rcmb
        if kw.get("anelastic") and ut is not None and 0 in (kw.get("bci", 1), kw.get("bco", 1)):
            # the two numbers the stress-free rows of an anelastic run need, computed as the reference does
            import bc_variables as bv
            import radial_profiles as rap
            lho = ut.chebco_f(rap.log_density, par.N, par.ricb, ut.rcmb, 1e-9)
            kw["lho1_icb"], kw["lho1_cmb"] = float(np.dot(lho, bv.Ta[:, 1])), float(np.dot(lho, bv.Tb[:, 1]))
        return cls(**kw)

    @classmethod
    def from_dict(cls, d):
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})

    # derived (utils.py:19-45)
    @property
    def wf(self):
        return 0 if self.forcing == 0 else self.forcing_frequency

    @property
    def N1(self):
        return self.N if self.ricb > 0 else self.N // 2

    @property
    def nb(self):
        return (self.lmax - self.m + 1) // 2

    @property
    def sizmat(self):
        return self.N1 * self.nb * (2 * self.hydro + 2 * self.magnetic + self.thermal + self.compositional)

    @property
    def dipole(self):
        """the background field whose equations carry two more powers of r (operators.py: `cdipole`)"""
        return bool(self.magnetic) and self.B0 == "dipole"

    def check_supported(self):
        if (self.lmax - self.m + 1) % 2:
            raise ValueError("lmax - m + 1 must be even (parameters.py:301-303): lmax = %d, m = %d" % (self.lmax, self.m))
        bad = []
        if self.magnetic:
            # every degree-1 field of the reference shares the axial field's block structure; only the radial
            # operators r^X h^(j) D^Y differ (operators.py:225-228, 249-252, 285-287, 318-320 ...)
            if self.B0 not in ("axial", "dipole", "G21 dipole", "Luo_S1", "FDM") or (
                    self.B0 == "FDM" and self.B0_l != 1):
                bad.append("B0 = %r" % (self.B0,))
            if self.innercore not in ("insulator", "TWA") or self.mantle not in ("insulator", "TWA"):
                bad.append("innercore / mantle other than 'insulator' or 'TWA'")
            if self.ricb <= 0 and self.B0 == "dipole":
                bad.append("B0 = %r without inner core" % (self.B0,))
        if self.compositional:
            if self.OmgTau is None:
                bad.append("compositional = 1 without OmgTau (parameters.py:268-271 leaves it commented out; "
                           "operators.py:423, 824 need it)")
            if self.anelastic or self.comp_background not in ("internal", "differential"):
                bad.append("compositional = 1 with anelastic = 1 or comp_background = %r" % (self.comp_background,))
        if self.anelastic:
            if self.ricb <= 0:
                bad.append("anelastic = 1 without inner core")
            if self.dipole:
                bad.append("anelastic = 1 with a dipole field (the reference's viscous terms then refer to operators "
                           "submatrices.py does not generate)")
            if (self.bci == 0 and self.lho1_icb is None) or (self.bco == 0 and self.lho1_cmb is None):
                bad.append("anelastic = 1 with stress-free boundaries but without lho1_icb / lho1_cmb "
                           "(d ln(rho) / dr at the boundaries; PhysicsParams.from_modules(par, ut) computes them)")
        if (self.Ek == 0 or self.ViscosD == 0) and not (self.Ek == 0 and self.ViscosD == 0 and self.ricb == 0
                                                        and not self.magnetic and not self.anelastic):
            bad.append("Ek = 0 or ViscosD = 0 other than the inviscid full sphere (Ek = ViscosD = 0, ricb = 0)")
        if not self.hydro:
            bad.append("hydro = 0")
        if self.forcing not in (0, 7, 9, 10):
            bad.append("forcing = %d (the reference's own code for it refers to undefined names)" % self.forcing)
        if self.thermal and self.ThermaD < 0:
            bad.append("thermal = 1 with ThermaD < 0")
        if self.thermal and self.ThermaD == 0 and self.anelastic:
            bad.append("anelastic = 1 without thermal diffusion")
        if self.thermal and self.heating not in ("differential", "internal", "two zone", "user defined"):
            bad.append("heating = %r" % (self.heating,))
        if bad:
            raise NotImplementedError("device-side assembly does not cover: " + ", ".join(bad)
                                      + " -- assemble on the host and use kb_set_pencil")


# ---------------------------------------------------------------------------
# radial operators
# ---------------------------------------------------------------------------
def load_operators(directory="."):
    """Every ``*.mtx`` of `directory` as CSR, keyed by its label (what operators.py:10-13 does
    with globals)."""
    import scipy.io as sio
    import scipy.sparse as sp
    ops = {}
    for fn in sorted(glob.glob(os.path.join(directory, "*.mtx"))):
        ops[os.path.basename(fn)[:-4]] = sp.csr_matrix(sio.mmread(fn))
    if not ops:
        raise FileNotFoundError("no *.mtx radial operators in %r (run submatrices.py first)" % directory)
    return ops


def save_operators_npz(path, ops):
    store = {}
    for k, M in ops.items():
        M = M.tocsr()
        store[k + "/data"], store[k + "/indices"], store[k + "/indptr"] = M.data, M.indices, M.indptr
        store[k + "/shape"] = np.asarray(M.shape)
    np.savez_compressed(path, **store)


def load_operators_npz(path):
    import scipy.sparse as sp
    z = np.load(path)
    labels = sorted({k.split("/")[0] for k in z.files})
    return {k: sp.csr_matrix((z[k + "/data"], z[k + "/indices"], z[k + "/indptr"]), shape=tuple(z[k + "/shape"]))
            for k in labels}


def endpoint_table(x, N, pmax, ric, rcmb):
    """``T[k, p]`` = p-th derivative with respect to r of the Chebyshev polynomial T_k at the end
    point x = -1 (r = ric) or x = +1 (r = rcmb), k = 0..N, p = 0..pmax:
    T_k^(p)(+-1) = (+-1)^(k+p) prod_{i<p} (k^2 - i^2) / (2 i + 1), times (2 / (rcmb - ric))^p for
    the change of variable.  The products are accumulated in the order bc_variables.py:8-32 uses,
    so that the boundary rows match the reference's to the bit."""
    T = np.zeros((N + 1, pmax + 1))
    for k in range(N + 1):
        T[k, 0] = x ** k
        acc = 1.0
        for i in range(pmax):
            acc = acc * (k ** 2 - i ** 2) / (2 * i + 1)
            T[k, i + 1] = x ** (k + i + 1) * acc * (2 / (rcmb - ric)) ** (i + 1)
    return T


# ---------------------------------------------------------------------------
# the program
# ---------------------------------------------------------------------------
@dataclass
class Group:
    """One summand of a block: ``sign * s_k(...s_1(sum_j c_j x_j))`` added to the real (part 0)
    or imaginary (part 1) component."""
    part: int
    sign: int
    scales: list
    terms: list  # (coefficient, operator label)


@dataclass
class AsmProgram:
    """Flat arrays of an assembly program (the argument of `kb_assemble`)."""
    N1: int
    nblockrows: int
    H: int
    is_complex: bool
    op_labels: list
    ops: np.ndarray        # float64 [nop, N1, 2H+1]: ops[k, i, d] = R_k[i, i + d - H]
    bc: np.ndarray         # float64 [nbc, N1]: dense boundary rows
    br_chop: np.ndarray    # int32 [nblockrows]: boundary rows at the top of every block of this block row
    br_bc: np.ndarray      # int32 [nblockrows]: first row of `bc` of this block row's boundary rows
    blk_ptr: np.ndarray    # int32 [nblockrows + 1]
    blk_col: np.ndarray    # int32 [nblk]: block column, ascending inside a block row
    blk_grp: np.ndarray    # int32 [nblk + 1]
    grp_part: np.ndarray   # int32 [ngrp]
    grp_sign: np.ndarray   # int32 [ngrp]
    grp_nsc: np.ndarray    # int32 [ngrp]
    grp_sc: np.ndarray     # float64 [ngrp, MAX_SCALES]
    grp_term: np.ndarray   # int32 [ngrp + 1]
    term_coef: np.ndarray  # float64 [nterm]
    term_op: np.ndarray    # int32 [nterm]
    final_scale: float = 1.0
    use_final: bool = False
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.N1 * self.nblockrows

    def with_final_scale(self, s):
        import copy
        q = copy.copy(self)
        q.final_scale, q.use_final = float(s), True
        return q


class _Builder:
    def __init__(self, pp: PhysicsParams, operators: dict, is_complex: bool):
        self.pp = pp
        self.opsrc = operators
        self.is_complex = is_complex
        self.labels = []
        self.blocks = {}  # (block row, block col) -> [Group]

    def op_id(self, label):
        if label not in self.opsrc:
            raise KeyError("radial operator %s.mtx is missing" % label)
        if label not in self.labels:
            self.labels.append(label)
        return self.labels.index(label)

    def add(self, brow, bcol, group: Group):
        if len(group.scales) > MAX_SCALES:
            raise ValueError("more than %d scalar factors in one group" % MAX_SCALES)
        self.blocks.setdefault((int(brow), int(bcol)), []).append(group)

    def finish(self, nblockrows, br_chop, br_bc, bc_rows, meta):
        N1 = self.pp.N1
        H = 0
        csr = []
        for lab in self.labels:
            M = self.opsrc[lab].tocoo()
            if M.shape != (N1, N1):
                raise ValueError("operator %s is %r, expected %r" % (lab, M.shape, (N1, N1)))
            nz = M.data != 0
            if nz.any():
                H = max(H, int(np.abs(M.col[nz] - M.row[nz]).max()))
            csr.append(M)
        if 2 * H + 1 > MAX_BAND:
            raise ValueError("radial operators wider than +-%d are not supported (found +-%d)" % ((MAX_BAND - 1) // 2, H))
        W = 2 * H + 1
        ops = np.zeros((max(1, len(self.labels)), N1, W))
        for k, M in enumerate(csr):
            ops[k, M.row, M.col - M.row + H] = M.data
        keys = sorted(self.blocks)
        blk_ptr = np.zeros(nblockrows + 1, dtype=np.int32)
        for (r, _) in keys:
            blk_ptr[r + 1] += 1
        blk_ptr = np.cumsum(blk_ptr).astype(np.int32)
        blk_col = np.array([c for (_, c) in keys], dtype=np.int32)
        blk_grp, gp, gs, gn, gsc, gt, tc, to = [0], [], [], [], [], [0], [], []
        for key in keys:
            for g in self.blocks[key]:
                gp.append(g.part)
                gs.append(g.sign)
                gn.append(len(g.scales))
                gsc.append(list(map(float, g.scales)) + [1.0] * (MAX_SCALES - len(g.scales)))
                for (c, lab) in g.terms:
                    tc.append(float(c))
                    to.append(self.op_id_existing(lab))
                gt.append(len(tc))
            blk_grp.append(len(gp))
        return AsmProgram(
            N1=N1, nblockrows=nblockrows, H=H, is_complex=self.is_complex, op_labels=list(self.labels), ops=ops,
            bc=np.ascontiguousarray(bc_rows, dtype=np.float64).reshape(-1, N1),
            br_chop=np.asarray(br_chop, dtype=np.int32), br_bc=np.asarray(br_bc, dtype=np.int32),
            blk_ptr=blk_ptr, blk_col=blk_col, blk_grp=np.asarray(blk_grp, dtype=np.int32),
            grp_part=np.asarray(gp, dtype=np.int32), grp_sign=np.asarray(gs, dtype=np.int32),
            grp_nsc=np.asarray(gn, dtype=np.int32), grp_sc=np.asarray(gsc, dtype=np.float64).reshape(-1, MAX_SCALES),
            grp_term=np.asarray(gt, dtype=np.int32), term_coef=np.asarray(tc, dtype=np.float64),
            term_op=np.asarray(to, dtype=np.int32), meta=meta)

    def op_id_existing(self, lab):
        return self.labels.index(lab)


def _lin(b, *terms):
    """[(coefficient, label)] with the labels registered."""
    out = []
    for c, lab in terms:
        b.op_id(lab)
        out.append((c, lab))
    return out


def _sections(pp):
    secs = section_degrees(pp.m, pp.lmax, pp.symm, -1, pp.hydro, pp.magnetic, pp.thermal, pp.compositional)
    return {name: (base * pp.nb, np.asarray(degs)) for (name, base, degs) in secs}


def _block_of(sec, l):
    base, degs = sec
    k = np.nonzero(degs == l)[0]
    return None if k.size == 0 else base + int(k[0])


def _boundary_rows(pp, l=None):
    """Dense boundary rows per section (assemble.py:1188-1345, non-anelastic: the log-density
    terms vanish); their number is the number of rows submatrices.py:582-588 leaves empty.  Only the
    radial boundary-flow forcing (forcing = 9) with a stress-free outer boundary makes them depend on
    the degree `l` (assemble.py:1236-1241)."""
    ric = pp.ricb if pp.ricb > 0 else -pp.rcmb
    Ta = endpoint_table(-1, pp.N - 1, 4, ric, pp.rcmb)
    Tb = endpoint_table(1, pp.N - 1, 5, ric, pp.rcmb)
    s = (pp.symm + 1) // 2
    if pp.ricb > 0:
        Tbu = Tbv = Tbh = Tb
    else:
        ixu, ixv = (pp.m + 1 - s) % 2, (pp.m + s) % 2
        Tbu, Tbv, Tbh = Tb[ixu::2, :], Tb[ixv::2, :], Tb[ixu::2, :]
    rows = {}
    la = pp.lho1_icb if (pp.anelastic and pp.lho1_icb is not None) else 0.
    lb = pp.lho1_cmb if (pp.anelastic and pp.lho1_cmb is not None) else 0.
    u = [Tbu[:, 0]]
    if pp.forcing == 9:
        L = l * (l + 1)
        u.append(Tbu[:, 2] - (2 - L) * Tbu[:, 0] / pp.rcmb ** 2 if pp.bco == 0 else Tbu[:, 1] + Tbu[:, 0])
    else:
        u.append(pp.rcmb * Tbu[:, 2] - lb * Tbu[:, 1] if pp.bco == 0 else Tbu[:, 1])
    v = [-pp.rcmb * Tbv[:, 1] + (1 + pp.rcmb * lb) * Tbv[:, 0] if pp.bco == 0 else Tbv[:, 0]]
    h = [Tbh[:, 0] if pp.bco_thermal == 0 else Tbh[:, 1]]
    if pp.Ek == 0:
        u, v = [Tbu[:, 0]], []  # inviscid (full sphere): no penetration, nothing on the toroidal scalar (assemble.py:1205-1227, 1266)
    if pp.ricb > 0:
        u.append(Ta[:, 0])
        u.append(pp.ricb * Ta[:, 2] - la * Ta[:, 1] if pp.bci == 0 else Ta[:, 1])
        v.append(-pp.ricb * Ta[:, 1] + (1 + pp.ricb * la) * Ta[:, 0] if pp.bci == 0 else Ta[:, 0])
        h.append(Ta[:, 0] if pp.bci_thermal == 0 else Ta[:, 1])
    if pp.ThermaD == 0:
        h = []  # no thermal diffusion: a first-order equation in the C^(0) basis, no boundary rows (submatrices.py:177-178)
    if pp.compositional:  # assemble.py:1345-1380
        ci = [Tbh[:, 0] if pp.bco_compositional == 0 else Tbh[:, 1]]
        if pp.ricb > 0:
            ci.append(Ta[:, 0] if pp.bci_compositional == 0 else Ta[:, 1])
        rows["i"] = np.array(ci)
    width = len(u[0])
    rows["u"], rows["v"], rows["h"] = np.array(u), np.array(v).reshape(len(v), width), np.array(h).reshape(len(h), width)
    if pp.magnetic:
        if pp.ricb > 0:
            Tbf = Tbg = Tb
        else:
            # full sphere: the outer row only, on the Chebyshev polynomials of the section's parity -- the field
            # induced by an antisymmetric background field has the opposite one to the flow's (assemble.py:1479-1492)
            Tbf, Tbg = Tb[(pp.m + s) % 2::2, :], Tb[(pp.m + 1 - s) % 2::2, :]

        def thin_layer(T, Tg, rj, eps, c, c1):
            # a thin conducting layer between the fluid and the insulator (assemble.py:1383-1467)
            mu_vf = 1 / pp.mu
            F, F1, F2 = rj * T[:, 0], rj * T[:, 1] + T[:, 0], 2 * T[:, 1] + rj * T[:, 2]
            G, G1 = rj * Tg[:, 0], rj * Tg[:, 1] + Tg[:, 0]
            kj = (l + 0.5) * eps - 0.5
            nabF = F2 - l * (l + 1) * F / rj
            return (mu_vf * F1 + (kj / rj) * F + eps * kj * c * F1 + eps * c1 * rj * (mu_vf + 0.5 * eps * kj * c) * nabF,
                    G + eps * rj * c1 * G1)
        # insulating inner core and mantle: the field matches a potential field on either side
        # (assemble.py:1545-1572 inner row, 1479-1536 outer row); inner boundary first
        if pp.mantle == "TWA":
            fo, go = thin_layer(Tbf, Tbg, pp.rcmb, 1, pp.c_cmb, pp.c1_cmb)
        else:
            fo, go = (l + 1) * Tbf[:, 0] + pp.rcmb * Tbf[:, 1], Tbg[:, 0]
        if pp.ricb > 0:
            if pp.innercore == "TWA":
                fi, gi = thin_layer(Ta, Ta, pp.ricb, -1, pp.c_icb, pp.c1_icb)
            else:
                fi, gi = l * Ta[:, 0] - pp.ricb * Ta[:, 1], Ta[:, 0]
            rows["f"], rows["g"] = np.array([fi, fo]), np.array([gi, go])
        else:
            rows["f"], rows["g"] = np.array([fo]), np.array([go])
    return rows


def _coriolis_down(l, m):
    # coupling of degree l to degree l - 1 (operators.py:72, 98)
    return (l ** 2 - 1) * np.sqrt(l ** 2 - m ** 2) / (2 * l - 1.)


def _coriolis_up(l, m):
    # coupling of degree l to degree l + 1 (operators.py:83, 109)
    return l * (l + 2.) * np.sqrt((l + m + 1.) * (l - m + 1)) / (2. * l + 3.)


def build_program_A(pp: PhysicsParams, operators: dict) -> AsmProgram:
    """Assembly program of ``A`` (before the division by ||B||_F)."""
    pp.check_supported()
    secs = _sections(pp)
    b = _Builder(pp, operators, True)
    m, wf = pp.m, pp.wf
    RE, IM = 0, 1
    G, V, Bf, Td = pp.Gaspard, pp.ViscosD, pp.Beyonce, pp.ThermaD
    # with a dipole background field the momentum equations are multiplied by two (2curl) / three (1curl)
    # more powers of r (operators.py:33-43 and every `cdipole` branch)
    U = lambda k, d: "r%d_D%d_u" % (k + (2 if pp.dipole else 0), d)  # noqa: E731
    W = lambda k, d: "r%d_D%d_v" % (k + (3 if pp.dipole else 0), d)  # noqa: E731

    for l in secs["u"][1]:  # ---- poloidal momentum (2curl) rows
        l = int(l)
        L = l * (l + 1)
        r = _block_of(secs["u"], l)
        b.add(r, r, Group(IM, +1, [L, wf], _lin(b, (L, U(2, 0)), (-2, U(3, 1)), (-1, U(4, 2)))))
        b.add(r, r, Group(IM, +1, [2 * m, G], _lin(b, (-L, U(2, 0)), (2, U(3, 1)), (1, U(4, 2)))))
        if pp.anelastic and pp.variable_viscosity:
            # ... and a viscosity profile nu(r) (operators.py:154-168).  The reference nests two of the sums
            # (2(L-1)(a + b), 6(a + b)); here they are distributed, so this block agrees to rounding, not to the bit
            b.add(r, r, Group(RE, -1, [L, V], _lin(
                b, (L * (2 - L), "r0_vsc0_D0_u"), (-(L + 2), "r1_vsc0_lho1_D0_u"), (-2 * (L + 1), "r1_vsc1_D0_u"),
                (-2 * (L - 1), "r2_vsc0_lho2_D0_u"), (-2 * (L - 1), "r2_vsc1_lho1_D0_u"), (2 - L, "r2_vsc2_D0_u"),
                (1, "r3_vsc0_lho3_D0_u"), (2, "r3_vsc1_lho2_D0_u"), (1, "r3_vsc2_lho1_D0_u"),
                (2 - L, "r2_vsc0_lho1_D1_u"), (6, "r3_vsc0_lho2_D1_u"), (6, "r3_vsc1_lho1_D1_u"),
                (1, "r4_vsc0_lho3_D1_u"), (2, "r4_vsc1_lho2_D1_u"), (1, "r4_vsc2_lho1_D1_u"), (2 * (L + 1), "r2_vsc1_D1_u"),
                (2 * L, "r2_vsc0_D2_u"), (5, "r3_vsc0_lho1_D2_u"), (-4, "r3_vsc1_D2_u"), (2, "r4_vsc0_lho2_D2_u"),
                (2, "r4_vsc1_lho1_D2_u"), (-1, "r4_vsc2_D2_u"),
                (-4, "r3_vsc0_D3_u"), (1, "r4_vsc0_lho1_D3_u"), (-2, "r4_vsc1_D3_u"), (-1, "r4_vsc0_D4_u"))))
        elif pp.anelastic:  # viscous force with the density stratification (operators.py:146-152)
            b.add(r, r, Group(RE, -1, [L, V], _lin(
                b, (-L * (l + 2) * (l - 1), "r0_D0_u"), (-(L + 2), "r1_lho1_D0_u"), (-2 * (L - 1), "r2_lho2_D0_u"),
                (1, "r3_lho3_D0_u"), (-(L - 2), "r2_lho1_D1_u"), (6, "r3_lho2_D1_u"), (1, "r4_lho3_D1_u"),
                (2 * L, "r2_D2_u"), (5, "r3_lho1_D2_u"), (2, "r4_lho2_D2_u"), (-4, "r3_D3_u"), (1, "r4_lho1_D3_u"),
                (-1, "r4_D4_u"))))
        else:
            b.add(r, r, Group(RE, -1, [L, V], _lin(b, (-L * (l + 2) * (l - 1), U(0, 0)), (2 * L, U(2, 2)),
                                                   (-4, U(3, 3)), (-1, U(4, 4)))))
        c = _block_of(secs["v"], l - 1)
        if c is not None:
            b.add(r, c, Group(RE, +1, [2 * _coriolis_down(l, m), G], _lin(b, (l - 1, U(3, 0)), (-1, U(4, 1)))))
        c = _block_of(secs["v"], l + 1)
        if c is not None:
            b.add(r, c, Group(RE, +1, [2 * _coriolis_up(l, m), G], _lin(b, (-(l + 2), U(3, 0)), (-1, U(4, 1)))))
        if pp.thermal:
            b.add(r, _block_of(secs["h"], l), Group(RE, +1, [L, Bf], _lin(b, (1, "r3_buo0_D0_u" if pp.anelastic else U(4, 0)))))
        if pp.compositional:  # operators.py:408-423
            b.add(r, _block_of(secs["i"], l), Group(RE, +1, [L, pp.OmgTau ** 2 * pp.BV2_comp], _lin(b, (1, U(4, 0)))))

    for l in secs["v"][1]:  # ---- toroidal momentum (1curl) rows
        l = int(l)
        L = l * (l + 1)
        r = _block_of(secs["v"], l)
        c = _block_of(secs["u"], l - 1)
        if c is not None:
            b.add(r, c, Group(RE, +1, [2 * _coriolis_down(l, m), G], _lin(b, (l - 1, W(1, 0)), (-1, W(2, 1)))))
        c = _block_of(secs["u"], l + 1)
        if c is not None:
            b.add(r, c, Group(RE, +1, [2 * _coriolis_up(l, m), G], _lin(b, (-(l + 2), W(1, 0)), (-1, W(2, 1)))))
        b.add(r, r, Group(IM, +1, [L, wf], _lin(b, (1, W(2, 0)))))
        b.add(r, r, Group(IM, +1, [-2 * m, G], _lin(b, (1, W(2, 0)))))
        if pp.anelastic and pp.variable_viscosity:  # operators.py:181-185
            b.add(r, r, Group(RE, -1, [L, V], _lin(
                b, (-L, "r0_vsc0_D0_v"), (-3, "r1_vsc0_lho1_D0_v"), (-1, "r2_vsc1_lho1_D0_v"), (-1, "r2_vsc0_lho2_D0_v"),
                (-1, "r1_vsc1_D0_v"), (2, "r1_vsc0_D1_v"), (-1, "r2_vsc0_lho1_D1_v"), (1, "r2_vsc1_D1_v"), (1, "r2_vsc0_D2_v"))))
        elif pp.anelastic:  # operators.py:176-179
            b.add(r, r, Group(RE, -1, [L, V], _lin(b, (-L, "r0_D0_v"), (-3, "r1_lho1_D0_v"), (-1, "r2_lho2_D0_v"),
                                                   (2, "r1_D1_v"), (-1, "r2_lho1_D1_v"), (1, "r2_D2_v"))))
        else:
            b.add(r, r, Group(RE, -1, [L, V], _lin(b, (-L, W(0, 0)), (2, W(1, 1)), (1, W(2, 2)))))

    if pp.thermal:  # ---- heat equation rows
        gap = pp.rcmb - pp.ricb
        diff = pp.heating == "differential"
        for l in secs["h"][1]:
            l = int(l)
            L = l * (l + 1)
            r = _block_of(secs["h"], l)
            c = _block_of(secs["u"], l)
            if pp.anelastic:  # entropy equation (operators.py:709-711, 731-732, 765-768)
                b.add(r, c, Group(RE, +1, [L], _lin(b, (-1, "r1_tds0_D0_h"))))
                b.add(r, r, Group(RE, +1, [Td], _lin(b, (-L, "r0_krT0_D0_h"), (2, "r1_krT0_D1_h"), (1, "r2_krT1_D1_h"),
                                                     (1, "r2_krT0_D2_h"))))
                b.add(r, r, Group(IM, -1, [wf], _lin(b, (1, "r2_roT0_D0_h"))))
            elif diff:
                b.add(r, c, Group(RE, +1, [pp.ricb, 1. / gap, L], _lin(b, (1, "r0_D0_h"))))
                if Td > 0:
                    b.add(r, r, Group(RE, +1, [Td], _lin(b, (-L, "r1_D0_h"), (2, "r2_D1_h"), (1, "r3_D2_h"))))
                b.add(r, r, Group(IM, -1, [wf], _lin(b, (1, "r3_D0_h"))))
            else:
                # internal heating, or a background gradient of the run's own (r dT/dr in radial_profiles.twozone /
                # BVprof, operators.py:737-738)
                own = pp.heating in ("two zone", "user defined")
                b.add(r, c, Group(RE, +1, [L], _lin(b, (-1, "r0_drS0_D0_h") if own else (1, "r2_D0_h"))))
                if Td > 0:
                    b.add(r, r, Group(RE, +1, [Td], _lin(b, (-L, "r0_D0_h"), (2, "r1_D1_h"), (1, "r2_D2_h"))))
                b.add(r, r, Group(IM, -1, [wf], _lin(b, (1, "r2_D0_h"))))

    if pp.compositional:  # ---- composition equation rows (operators.py:776-824; no frequency term in A, as there)
        gap = pp.rcmb - pp.ricb
        for l in secs["i"][1]:
            l = int(l)
            L = l * (l + 1)
            r = _block_of(secs["i"], l)
            c = _block_of(secs["u"], l)
            Dc = [pp.OmgTau, pp.Ek, 1. / pp.Schmidt]
            if pp.comp_background == "differential":
                b.add(r, c, Group(RE, +1, [pp.ricb, 1. / gap, L], _lin(b, (1, "r0_D0_i"))))
                b.add(r, r, Group(RE, +1, Dc, _lin(b, (-L, "r1_D0_i"), (2, "r2_D1_i"), (1, "r3_D2_i"))))
            else:
                b.add(r, c, Group(RE, +1, [L], _lin(b, (1, "r2_D0_i"))))
                b.add(r, r, Group(RE, +1, Dc, _lin(b, (-L, "r0_D0_i"), (2, "r1_D1_i"), (1, "r2_D2_i"))))

    if pp.magnetic:
        _magnetic_blocks(b, pp, secs)

    return _finish(b, pp, secs, with_bc=True)


def _magnetic_blocks(b, pp, secs):
    """Lorentz force in the momentum rows, induction and magnetic diffusion in the field rows, for an
    antisymmetric background field (axial or dipole): only the dipole-type couplings (degree l to l +- 1
    across the families, l to l inside a family) exist (operators.py:198-384 lorentz, :433-468 b,
    :471-643 induction, :646-692 magnetic_diffusion; placement assemble.py:640-668, 765-797, 820-1008).

    NOT bit-compatible with the reference, unlike the rest of the program: the reference evaluates the
    induction coefficients in numpy.float128 (operators.py:479) and nests its sums; here every block is
    ONE left-to-right sum with double coefficients (the induction ones rounded from the same extended
    precision products), which agrees with the reference's entries to a few ulp of the block's largest
    term (tests: 1e-13 of the block maximum; eigenvalues to 1e-9)."""
    m, wf = pp.m, pp.wf
    RE, IM = 0, 1
    H, Md = pp.Hendrik, pp.MagnetD
    d2 = 2 if pp.dipole else 0   # extra powers of r in the 2curl momentum and in the poloidal induction rows
    d3 = 3 if pp.dipole else 0   # ... in the 1curl momentum and in the toroidal induction rows
    u_ = lambda k, h, d: "r%d_h%d_D%d_u" % (k + d2, h, d)  # noqa: E731
    v_ = lambda k, h, d: "r%d_h%d_D%d_v" % (k + d3, h, d)  # noqa: E731
    f_ = lambda k, h, d: "r%d_h%d_D%d_f" % (k + d2, h, d)  # noqa: E731
    g_ = lambda k, h, d: "r%d_h%d_D%d_g" % (k + d3, h, d)  # noqa: E731
    sq = np.sqrt

    for l in secs["u"][1]:  # ---- Lorentz force, poloidal momentum rows
        l = int(l)
        L = l * (l + 1)
        r = _block_of(secs["u"], l)
        c = _block_of(secs["f"], l - 1)
        if c is not None:
            C = sq(l ** 2 - m ** 2) * (l ** 2 - 1) / (2 * l - 1)
            b.add(r, c, Group(RE, +1, [C, H], _lin(
                b, (-2 * (l ** 2 + 2), u_(1, 0, 1)), (-2 * (l - 2), u_(2, 1, 1)), (-(l - 4), u_(2, 0, 2)), (-(l - 2), u_(3, 1, 2)),
                (L * (l + 2), u_(0, 0, 0)), (L * (l - 4), u_(1, 1, 0)), (l, u_(2, 2, 0)), (l, u_(3, 3, 0)), (2, u_(3, 0, 3)))))
        c = _block_of(secs["f"], l + 1)
        if c is not None:
            C = sq((1 + l + m) * (1 + l - m)) * l * (l + 2) / (2 * l + 3)
            b.add(r, c, Group(RE, +1, [C, H], _lin(
                b, (-2 * (l ** 2 + 2 * l + 3), u_(1, 0, 1)), (2 * (l + 3), u_(2, 1, 1)), (l + 5, u_(2, 0, 2)), (l + 3, u_(3, 1, 2)),
                (-L * (l - 1), u_(0, 0, 0)), (-L * (l + 5), u_(1, 1, 0)), (-(l + 1), u_(2, 2, 0)), (-(l + 1), u_(3, 3, 0)),
                (2, u_(3, 0, 3)))))
        c = _block_of(secs["g"], l)
        b.add(r, c, Group(IM, +1, [2 * m, H], _lin(
            b, (-1, u_(1, 0, 0)), (-(l ** 2 + l - 1), u_(2, 1, 0)), (1, u_(2, 0, 1)), (1, u_(3, 1, 1)), (1, u_(3, 0, 2)))))

    for l in secs["v"][1]:  # ---- Lorentz force, toroidal momentum rows
        l = int(l)
        L = l * (l + 1)
        r = _block_of(secs["v"], l)
        c = _block_of(secs["f"], l)
        b.add(r, c, Group(IM, +1, [m, H], _lin(
            b, (4, v_(0, 0, 1)), (-2 * L, v_(0, 1, 0)), (-L, v_(1, 2, 0)), (2, v_(1, 0, 2)))))
        c = _block_of(secs["g"], l - 1)
        if c is not None:
            C = sq((l - m) * (l + m)) * (l ** 2 - 1) / (2 * l - 1)
            b.add(r, c, Group(RE, +1, [C, H], _lin(b, (l - 2, v_(0, 0, 0)), (l, v_(1, 1, 0)), (-2, v_(1, 0, 1)))))
        c = _block_of(secs["g"], l + 1)
        if c is not None:
            C = -sq((l + m + 1) * (l + 1 - m)) * l * (l + 2) / (2 * l + 3)
            b.add(r, c, Group(RE, +1, [C, H], _lin(b, (l + 3, v_(0, 0, 0)), (l + 1, v_(1, 1, 0)), (2, v_(1, 0, 1)))))

    ld = np.longdouble

    def ext(C, *terms):
        # coefficients formed in extended precision as the reference does, then rounded once
        return [(float(ld(C) * ld(c)), lab) for c, lab in terms]

    # anelastic runs: the field equations carry the density (operators.py:445-462, 660-689) and the toroidal
    # induction three more terms with d ln(rho)/dr (:571-629)
    rho = "rho0_" if pp.anelastic else ""

    def eta(k, p, d, s_):
        k += d2 if s_ == "f" else d3
        if not pp.anelastic:
            return "r%d_eta%d_D%d_%s" % (k, p, d, s_)
        return ("r%d_eho0_D%d_%s" % (k, d, s_)) if p == 0 else ("r%d_eta1_rho0_D%d_%s" % (k, d, s_))
    gl_ = lambda k, h, d: "r%d_h%d_lho1_D%d_g" % (k + d3, h, d)  # noqa: E731
    for l in secs["f"][1]:  # ---- poloidal field rows: induction by the flow, diffusion, time derivative
        l = int(l)
        L = l * (l + 1)
        lx = ld(l)
        r = _block_of(secs["f"], l)
        c = _block_of(secs["u"], l - 1)
        if c is not None:
            C = np.sqrt(lx ** 2 - m ** 2) * (lx ** 2 - 1) / (2 * lx - 1)
            b.add(r, c, Group(RE, +1, [], _lin(b, *ext(C, (lx - 2, f_(0, 0, 0)), (lx, f_(1, 1, 0)), (-2, f_(1, 0, 1))))))
        c = _block_of(secs["u"], l + 1)
        if c is not None:
            C = np.sqrt((lx + 1) ** 2 - m ** 2) * lx * (lx + 2) / (2 * lx + 3)
            b.add(r, c, Group(RE, +1, [], _lin(b, *ext(C, (-(lx + 3), f_(0, 0, 0)), (-(lx + 1), f_(1, 1, 0)), (-2, f_(1, 0, 1))))))
        c = _block_of(secs["v"], l)
        b.add(r, c, Group(IM, +1, [], _lin(b, (-2 * m, f_(1, 0, 0)))))
        b.add(r, r, Group(IM, +1, [L, wf], _lin(b, (1, "r%d_%sD0_f" % (2 + d2, rho)))))
        b.add(r, r, Group(RE, -1, [L, Md], _lin(b, (-L, eta(0, 0, 0, "f")), (2, eta(1, 0, 1, "f")), (1, eta(2, 0, 2, "f")))))

    sg = -1 if pp.dipole else 1  # sign of the eta' terms of the toroidal diffusion (operators.py:684, 689)
    for l in secs["g"][1]:  # ---- toroidal field rows
        l = int(l)
        L = l * (l + 1)
        lx = ld(l)
        r = _block_of(secs["g"], l)
        c = _block_of(secs["u"], l)
        if pp.dipole:
            terms = ((1, g_(0, 0, 1)), (1, g_(1, 1, 1)), (-(lx ** 2 + lx + 1), g_(-1, 0, 0)), (1, g_(0, 1, 0)),
                     (ld(L) / 2, g_(1, 2, 0)), (1, g_(1, 0, 2)))
        else:
            terms = ((1, g_(0, 0, 1)), (1, g_(1, 1, 1)), (-(ld(L) + 1), "q1_h0_D0_g"), (1, g_(0, 1, 0)),
                     (ld(L) / 2, g_(1, 2, 0)), (1, g_(1, 0, 2)))
        b.add(r, c, Group(IM, +1, [], _lin(b, *ext(2 * m, *terms))))
        if pp.anelastic:
            half = (lx ** 2 + lx + 2) if pp.dipole else (ld(L) + 2)
            b.add(r, c, Group(IM, +1, [], _lin(b, *ext(2 * m, (-ld(L) / 2, gl_(1, 1, 0)), (-half / 2, gl_(0, 0, 0)), (-1, gl_(1, 0, 1))))))
        c = _block_of(secs["v"], l - 1)
        if c is not None:
            C = (lx ** 2 - 1) * np.sqrt(lx ** 2 - m ** 2) / (2 * lx - 1)
            b.add(r, c, Group(RE, +1, [], _lin(b, *ext(C, (lx, g_(0, 0, 0)), (-2, g_(1, 0, 1)), (lx - 2, g_(1, 1, 0))))))
            if pp.anelastic:
                b.add(r, c, Group(RE, +1, [], _lin(b, *ext(2 * C, (1, gl_(1, 0, 0))))))
        c = _block_of(secs["v"], l + 1)
        if c is not None:
            C = lx * (lx + 2) * np.sqrt((lx + 1) ** 2 - m ** 2) / (3 + 2 * lx)
            b.add(r, c, Group(RE, +1, [], _lin(b, *ext(C, (-2, g_(1, 0, 1)), (-(lx + 1), g_(0, 0, 0)), (-(lx + 3), g_(1, 1, 0))))))
            if pp.anelastic:
                b.add(r, c, Group(RE, +1, [], _lin(b, *ext(2 * C, (1, gl_(1, 0, 0))))))
        b.add(r, r, Group(IM, +1, [L, wf], _lin(b, (1, "r%d_%sD0_g" % (2 + d3, rho)))))
        b.add(r, r, Group(RE, -1, [L, Md], _lin(b, (2, eta(1, 0, 1, "g")), (-L, eta(0, 0, 0, "g")), (1, eta(2, 0, 2, "g")),
                                                (sg, eta(1, 1, 0, "g")), (sg, eta(2, 1, 1, "g")))))


def build_program_B(pp: PhysicsParams, operators: dict) -> AsmProgram:
    """Assembly program of ``B`` (real; before the division by its Frobenius norm)."""
    pp.check_supported()
    secs = _sections(pp)
    b = _Builder(pp, operators, False)
    for l in secs["u"][1]:
        l = int(l)
        L = l * (l + 1)
        r = _block_of(secs["u"], l)
        k = 2 if pp.dipole else 0
        b.add(r, r, Group(0, -1, [L], _lin(b, (L, "r%d_D0_u" % (2 + k)), (-2, "r%d_D1_u" % (3 + k)), (-1, "r%d_D2_u" % (4 + k)))))
    for l in secs["v"][1]:
        l = int(l)
        r = _block_of(secs["v"], l)
        b.add(r, r, Group(0, -1, [l * (l + 1)], _lin(b, (1, "r%d_D0_v" % (5 if pp.dipole else 2)))))
    if pp.magnetic:
        rho = "rho0_" if pp.anelastic else ""
        for name, lab in (("f", "r%d_%sD0_f" % (4 if pp.dipole else 2, rho)), ("g", "r%d_%sD0_g" % (5 if pp.dipole else 2, rho))):
            for l in secs[name][1]:
                l = int(l)
                r = _block_of(secs[name], l)
                b.add(r, r, Group(0, -1, [l * (l + 1)], _lin(b, (1, lab))))
    if pp.thermal:
        lab = "r2_roT0_D0_h" if pp.anelastic else ("r3_D0_h" if pp.heating == "differential" else "r2_D0_h")
        for l in secs["h"][1]:
            r = _block_of(secs["h"], int(l))
            b.add(r, r, Group(0, +1, [], _lin(b, (1, lab))))
    if pp.compositional:
        lab = "r3_D0_i" if pp.comp_background == "differential" else "r2_D0_i"
        for l in secs["i"][1]:
            r = _block_of(secs["i"], int(l))
            b.add(r, r, Group(0, +1, [], _lin(b, (1, lab))))
    return _finish(b, pp, secs, with_bc=False)


def _finish(b, pp, secs, with_bc):
    nbr = pp.sizmat // pp.N1
    br_chop = np.zeros(nbr, dtype=np.int32)
    br_bc = np.full(nbr, -1, dtype=np.int32)
    bc_rows = np.zeros((0, pp.N1))
    if with_bc:
        # sections whose boundary rows depend on the degree: u under forcing = 9 with a stress-free outer
        # boundary, f (poloidal field) always
        per_degree = (["u"] if pp.forcing == 9 and pp.bco == 0 else []) + (["f"] if pp.magnetic else [])
        rows = _boundary_rows(pp, 2)
        first = {}
        stack = []
        at = 0
        for name in secs:
            first[name] = at
            stack.append(rows[name])
            at += rows[name].shape[0]
        for name, (base, degs) in secs.items():
            br_chop[base:base + len(degs)] = rows[name].shape[0]
            br_bc[base:base + len(degs)] = first[name]
        for name in per_degree:
            base, degs = secs[name]
            for k, l in enumerate(degs):
                r = _boundary_rows(pp, int(l))[name]
                br_bc[base + k] = at
                stack.append(r)
                at += r.shape[0]
        bc_rows = np.vstack(stack)
    meta = {"sections": {k: (int(v[0]), [int(x) for x in v[1]]) for k, v in secs.items()}}
    return b.finish(nbr, br_chop, br_bc, bc_rows, meta)


def forcing_vector(pp: PhysicsParams):
    """Right-hand side of the forced problem, as the dense complex vector solve.py:211-216 makes of
    B_forced.npz.  forcing = 7, libration in longitude as a boundary flow (assemble.py:278-329): two
    entries, in the outer and inner boundary rows of the toroidal l = 1 block (m = 0, axial) or of the
    poloidal l = 2 block (m = 2), no-slip boundaries, symm = 1."""
    pp.check_supported()
    secs = _sections(pp)
    b = np.zeros(pp.sizmat, dtype=np.complex128)
    if pp.forcing in (9, 10):
        # radial velocity forcing at the boundaries, poloidal l = 2 (forcing = 9, m = 2: assemble.py:360-390) or
        # l = m (forcing = 10, unit amplitude at the outer boundary: assemble.py:392-422)
        if pp.symm != 1 or (pp.forcing == 9 and not (pp.m == 2 and pp.bci == 1)):
            raise ValueError("radial boundary forcing needs symm = 1 (forcing = 9: also m = 2 and bci = 1)")
        l = 2 if pp.forcing == 9 else pp.m
        L = l * (l + 1)
        row = _block_of(secs["u"], l) * pp.N1
        b[row] = pp.forcing_amplitude_cmb * pp.rcmb / L if pp.forcing == 9 else 1. / L * 1.0
        b[row + 1] = pp.forcing_amplitude_icb * pp.ricb / L if pp.forcing == 9 else 0.0
        return b
    if pp.forcing != 7:
        raise NotImplementedError("forcing vector of forcing = %d" % pp.forcing)
    if not (pp.symm == 1 and pp.bci == 1 and pp.bco == 1 and pp.m in (0, 2)):
        raise ValueError("longitudinal libration needs m = 0 or 2, symm = 1 and no-slip boundaries")
    w = pp.forcing_frequency
    if pp.m == 0:
        row = _block_of(secs["v"], 1) * pp.N1
        b[row] = 1j * w * pp.forcing_amplitude_cmb / 2
        b[row + 1] = 1j * w * pp.forcing_amplitude_icb * pp.ricb / 2
    else:
        l = 2
        L = l * (l + 1)
        row = _block_of(secs["u"], l) * pp.N1
        b[row] = -w / L * pp.forcing_amplitude_cmb
        b[row + 1] = -w / L * pp.forcing_amplitude_icb * (pp.ricb ** 2)
    return b


def frobenius_norm(values):
    """||B||_F the way assemble.py:583 gets it (scipy.sparse.linalg.norm of the CSR = the
    2-norm of its value array, numpy.linalg.norm -> BLAS dot) on the same values in the same
    order.  This is the reference's number to the bit ON THE SAME MACHINE: a threaded BLAS splits
    the dot product by core count, so the last bit of the norm -- and with it the last bit of
    every entry of A / ||B||_F -- differs between hosts for the reference too."""
    return float(np.linalg.norm(np.asarray(values)))


def assemble(solver, pp: PhysicsParams, operators: dict, bnorm=None):
    """Assemble the pencil of `pp` on the GPU of `solver` (a `kore_b200.lib.Solver`) and make
    it the handle's pencil, as `set_pencil` would with the reference-assembled CSR.

    Eigenvalue runs (forcing = 0): B is assembled un-normalised, its values come back once for
    the Frobenius norm (assemble.py:583-585; `frobenius_norm`), then A and B are assembled with
    the factor 1 / ||B||_F.  `bnorm`: use this norm instead (a repeated assembly with the same B;
    tests pin the norm of the machine that wrote the fixtures).  Forced runs: A only,
    un-normalised (assemble.py:1164).  Returns a dict with the programs and the norm."""
    progA = build_program_A(pp, operators)
    out = {"progA": progA, "bnorm": None}
    if pp.forcing == 0:
        progB = build_program_B(pp, operators)
        if bnorm is None:
            solver.assemble(None, progB)
            _, _, bvals = solver.get_assembled("B")
            bnorm = frobenius_norm(bvals)
        out["bnorm"] = bnorm
        out["progB"] = progB
        solver.assemble(progA.with_final_scale(1. / bnorm), progB.with_final_scale(1. / bnorm))
    else:
        solver.assemble(progA, None)
    return out
