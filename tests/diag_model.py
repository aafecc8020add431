"""NumPy model of the diagnostics kernel (kore_b200/csrc/kb_diag.cu): per-degree volume integrals of
an eigenvector / forced solution -- kinetic energy, viscous (kinetic and internal) dissipation,
buoyancy power; thermal energy, dissipation, advection -- by Chebyshev-Gauss quadrature.  Test
infrastructure: the same series evaluation (forward three-term recurrences of T_j and its first three
derivatives at the quadrature nodes) and the same integrands as the kernel, vectorised over the nodes."""
import numpy as np


def cheb_series(c, x, scale, nder):
    """[f, f', f'', f'''][:nder+1] at the points x of f = sum_j c_j T_j(x), derivatives with respect
    to r = x / scale + const (each derivative picks up one factor `scale`)."""
    K = len(x)
    T = np.zeros((4, 3, K))  # derivative order, (previous, current, next), node
    T[0, 0], T[0, 1] = 1.0, x
    T[1, 1] = 1.0
    acc = np.zeros((4, K), dtype=complex)
    acc[0] = c[0] * T[0, 0] + (c[1] * T[0, 1] if len(c) > 1 else 0)
    acc[1] = c[1] * T[1, 1] if len(c) > 1 else 0
    for j in range(1, len(c) - 1):
        # T_{j+1} = 2 x T_j - T_{j-1} and its derivatives: d^p: 2 p T^(p-1)_j + 2 x T^(p)_j - T^(p)_{j-1}
        for p in range(3, -1, -1):
            T[p, 2] = 2 * x * T[p, 1] - T[p, 0] + (2 * p * T[p - 1, 1] if p else 0)
        for p in range(4):
            acc[p] += c[j + 1] * T[p, 2]
            T[p, 0], T[p, 1] = T[p, 1].copy(), T[p, 2].copy()
    return [acc[p] * scale ** p for p in range(nder + 1)]


def diagnose(x, meta, heating="differential"):
    """(flow [n_l, 6], thermal [n_lp, 3]) of the solution vector x (Kore ordering), full radial domain."""
    from kore_b200 import chain
    N, ricb, rcmb, m, lmax, symm = meta["N"], meta["ricb"], 1.0, meta["m"], meta["lmax"], meta["symm"]
    N1, n = meta["N1"], meta["n"]
    lp, lt, ll = chain.ell(m, lmax, symm)
    k = np.arange(N)
    xk = np.cos((k + 0.5) * np.pi / N)
    Ra, Rb = ricb, rcmb
    rk = 0.5 * (Rb - Ra) * (xk + 1) + Ra
    r0 = ricb if ricb > 0 else -rcmb
    x0 = 2 * (rk - r0) / (rcmb - r0) - 1
    scale = 2.0 / (rcmb - r0)
    w = (np.pi / N) * np.sqrt(1 - xk ** 2) * (Rb - Ra) / 2
    s = (symm + 1) // 2
    iP, iT = (m + 1 - s) % 2, (m + s) % 2

    def coeffs(sec, idx, parity):
        c = x[sec * n + idx * N1: sec * n + (idx + 1) * N1]
        if ricb > 0:
            return c
        full = np.zeros(N, dtype=complex)
        full[parity::2] = c
        return full

    thermal = bool(meta["thermal"])
    flow = np.zeros((len(ll), 6))
    therm = np.zeros((len(lp), 3))
    r2 = rk ** 2
    for a, l in enumerate(ll):
        L = l * (l + 1)
        f0 = 4 * np.pi / (2 * l + 1)
        if l in lp:
            i = list(lp).index(l)
            P0, P1, P2, P3 = cheb_series(coeffs(0, i, iP), x0, scale, 3)
            q0 = L * P0 / rk
            s0 = P1 + P0 / rk
            q1 = (L * P1 - q0) / rk
            s1 = P2 + q1 / L
            q2 = (L * P2 - 2 * q1) / rk
            s2 = P3 + q2 / L
            ke = f0 * (r2 * abs(q0) ** 2 + r2 * L * abs(s0) ** 2)
            dk = 2 * np.real(f0 * (L * r2 * np.conj(s0) * s2 + 2 * rk * L * np.conj(s0) * s1
                                   - L ** 2 * np.conj(s0) * s0 - (l ** 2 + l + 2) * np.conj(q0) * q0
                                   + 2 * rk * np.conj(q0) * q1 + r2 * np.conj(q0) * q2
                                   + 2 * L * (np.conj(q0) * s0 + q0 * np.conj(s0))))
            di = 2 * f0 * (L * abs(q0 + rk * s1 - s0) ** 2 + 3 * abs(rk * q1) ** 2 + L * (l - 1) * (l + 2) * abs(s0) ** 2)
            wt = 0
            if thermal:
                (h0, h1, h2) = cheb_series(coeffs(2, i, iP), x0, scale, 2)
                wt = f0 * r2 * L * 2 * np.real(np.conj(P0) * h0)
                fr = 1 / rk if heating == "differential" else r2
                therm[i, 0] = np.sum(w * f0 * r2 * abs(h0) ** 2)
                therm[i, 1] = np.sum(w * f0 * (2 * rk * 2 * np.real(h0 * np.conj(h1)) + r2 * 2 * np.real(h0 * np.conj(h2))
                                               - 2 * L * abs(h0) ** 2))
                therm[i, 2] = np.sum(w * f0 * fr * L * 2 * np.real(np.conj(P0) * h0))
        else:
            i = list(lt).index(l)
            T0, T1, T2 = cheb_series(coeffs(1, i, iT), x0, scale, 2)
            ke = f0 * r2 * L * abs(T0) ** 2
            dk = 2 * np.real(f0 * (L * r2 * np.conj(T0) * T2 + 2 * rk * L * np.conj(T0) * T1 - L ** 2 * np.conj(T0) * T0))
            di = 2 * f0 * (L * abs(rk * T1 - T0) ** 2 + L * (l - 1) * (l + 2) * abs(T0) ** 2)
            wt = 0
        flow[a, 0], flow[a, 1], flow[a, 2] = np.sum(w * ke), np.sum(w * dk), np.sum(w * di)
        flow[a, 4] = np.sum(w * wt) if thermal else 0.0
    return flow, therm
