"""Host-side (numpy, vectorised) metric algebra needed by the post-processing around the traced endpoints: metric
components, tetrads, circular-orbit four-velocities, LNRF velocities.  None of this is the hot path (it is evaluated a
few hundred times per call, at source positions and disc hits), it only has to agree with the reference's formulas:

    metric components       src/metrics/kerr-metric.jl:4-30, src/metrics/johannsen-psaltis-ad.jl:4-26
    tetradframe             src/orthonormalization.jl:36-105
    lnrbasis                src/orthonormalization.jl:116-123
    CircularOrbits          src/orbits/circular-orbits.jl:10-130
    local_velocity, lorentz_factor   src/corona/flux-calculations.jl:13-36

All functions accept numpy arrays for r (complex allowed: the radial metric derivative is taken by a complex step,
which is exact to rounding for these rational metrics)."""
from __future__ import annotations

import math

import numpy as np

from . import api


def metric_components(m, r, theta):
    """(g_tt, g_rr, g_θθ, g_φφ, g_tφ), each broadcast over r, θ."""
    r = np.asarray(r)
    theta = np.asarray(theta, np.float64)
    if isinstance(m, api.MorrisThorneWormhole):  # morris-thorne-ad.jl:4-15 (one power of sin θ in g_φφ, as there)
        b2l2 = m.b**2 + r * r
        zero = 0 * r + 0 * theta
        return np.stack(np.broadcast_arrays(zero - 1, zero + 1, b2l2 + zero, b2l2 * np.sin(theta), zero))
    if isinstance(m, api.DilatonAxion):  # dilaton-axion-ad.jl:8-46
        M, a, be, b = m.M, m.a, m.beta, m.b
        bb, ba, bab = m.ratios()
        cth, sth = np.cos(theta), np.sin(theta)
        sigma = r * r + a * a * cth**2
        delta = r * r + a * a - 2 * M * r
        dhat = delta - (be**2 + 2 * b * r) - M * (M + 2 * b) * bb**2
        shat = sigma - (be**2 + 2 * b * r) + M**2 * bb * (bb - 2 * a * cth)
        dl = r * r - 2 * b * r + a * a
        W = 1 + (bab * (2 * cth - bab) + ba**2) / sth**2
        A = dl**2 - dhat * (W * a * sth) ** 2
        return np.stack(np.broadcast_arrays(-(dhat - a * a * sth**2) / shat, dhat * 0 + shat / dhat, shat + 0 * r, A * sth**2 / shat,
                                            -a * (dl - dhat * W) * sth**2 / shat))
    M, a = m.M, m.a
    c2 = np.cos(theta) ** 2
    s2 = np.sin(theta) ** 2
    sigma = r * r + a * a * c2
    delta = r * r - 2 * M * r + a * a
    if isinstance(m, api.KerrMetric):
        tt = -(1 - 2 * M * r / sigma)
        rr = sigma / delta
        thth = sigma + 0 * r
        phph = s2 * (r * r + a * a + 2 * M * a * a * r * s2 / sigma)
        tph = -2 * M * a * r * s2 / sigma
    elif isinstance(m, api.JohannsenPsaltisMetric):
        h = m.eps3 * M**3 * r / sigma**2
        tt = -(1 + h) * (1 - 2 * M * r / sigma)
        rr = sigma * (1 + h) / (delta + a * a * s2 * h)
        thth = sigma + 0 * r
        phph = s2 * (r * r + a * a + 2 * a * a * M * r * s2 / sigma) + h * a * a * (sigma + 2 * M * r) * s2 * s2 / sigma
        tph = -2 * a * M * r * s2 * (1 + h) / sigma
    elif isinstance(m, api.JohannsenMetric):
        f = m.eps3 * M**3 / r
        sigma = r * r + a * a * c2 + f
        A1 = 1 + m.alpha13 * (M / r) ** 3
        A2 = 1 + m.alpha22 * (M / r) ** 2
        A5 = 1 + m.alpha52 * (M / r) ** 2
        r2a2 = r * r + a * a
        den = (r2a2 * A1 - a * a * A2 * s2) ** 2
        tt = -sigma * (delta - a * a * A2**2 * s2) / den
        rr = sigma / (delta * A5)
        thth = sigma + 0 * r
        phph = sigma * s2 * (r2a2**2 * A1**2 - a * a * delta * s2) / den
        tph = -a * sigma * s2 * (r2a2 * A1 * A2 - delta) / den
    elif isinstance(m, api.BumblebeeMetric):
        tt = -(1 - 2 * M / r)
        rr = r * r / ((r * r - 2 * M * r) / (m.l + 1))
        thth = r * r + 0 * c2
        phph = r * r * s2
        tph = -2 * M * a * s2 / r
    elif isinstance(m, api.KerrNewmanMetric):
        delta = delta + m.Q**2
        r2a2 = r * r + a * a
        tt = (a * a * s2 - delta) / sigma
        rr = sigma / delta
        thth = sigma + 0 * r
        phph = (s2 / sigma) * (r2a2**2 - a * a * s2 * delta)
        tph = (a * s2 / sigma) * (delta - r2a2)
    else:
        raise ValueError(f"metric {type(m).__name__} is outside the B200 scope")
    return np.stack(np.broadcast_arrays(tt, rr, thth, phph, tph))


def metric_dr(m, r, theta, h=1e-30):
    """∂g/∂r of the five components (complex step)."""
    r = np.asarray(r, np.float64)
    return np.imag(metric_components(m, r + 1j * h, theta)) / h


def metric_matrix(g):
    """4×4 matrix (…, 4, 4) from components (5, …)."""
    g = np.asarray(g)
    out = np.zeros(g.shape[1:] + (4, 4), g.dtype)
    out[..., 0, 0], out[..., 1, 1], out[..., 2, 2], out[..., 3, 3] = g[0], g[1], g[2], g[3]
    out[..., 0, 3] = out[..., 3, 0] = g[4]
    return out


def inverse_metric_components(g):
    d = g[0] * g[3] - g[4] * g[4]
    return np.stack([g[3] / d, 1 / g[1], 1 / g[2], g[0] / d, -g[4] / d])


def dot(g, u, v):
    """g_{μν} u^μ v^ν with components g (5, …) and vectors (4, …)."""
    return (g[0] * u[0] * v[0] + g[1] * u[1] * v[1] + g[2] * u[2] * v[2] + g[3] * u[3] * v[3]
            + g[4] * (u[0] * v[3] + u[3] * v[0]))


def _charged_omega(m, r, dg, q, mu, contra_rotating):
    """CircularOrbits.Ω(::KerrNewmanMetric; q) (kerr-newman-ad.jl:113-146): the angular velocity of a circular orbit of a
    charged particle, root of  ½ ∂_r(g_tt + 2ω g_tφ + ω² g_φφ) − q (∂_r A_φ ω + ∂_r A_t) / u^t  (the reference writes the
    force through F^r_κ g_rr; g^rr g_rr = 1), Newton from Ω = r / 100 like `Roots.find_zero((f, f'), Ω_init)`."""
    g = metric_components(m, r, np.pi / 2)
    Sig = r * r  # equatorial plane
    At_r = m.Q * (Sig - 2 * r * r) / Sig**2  # ∂_r (r Q / Σ) at θ = π/2
    Ap_r = -m.a * At_r                        # A_φ = −a sin²θ A_t

    def f(w):
        delta = w * w * dg[3] + 2 * w * dg[4] + dg[0]
        arg = -(w * w * g[3] + 2 * w * g[4] + g[0]) / mu**2
        inv_ut = np.sign(arg) * np.sqrt(abs(arg))
        return 0.5 * delta - (Ap_r * w + At_r) * q * inv_ut

    w = -r / 100 if contra_rotating else r / 100
    for _ in range(100):
        h = 1e-7 * max(abs(w), 1e-3)
        step = f(w) / ((f(w + h) - f(w - h)) / (2 * h))
        w -= step
        if abs(step) <= 1e-15 * max(abs(w), 1e-300):
            break
    return w


def circular_fourvelocity(m, r, contra_rotating=False, q=0.0, mu=1.0):
    """CircularOrbits.fourvelocity(m, r) in the equatorial plane: (u^t, 0, 0, u^φ) (circular-orbits.jl:10-130); `q`, `mu`:
    charged particles in the Kerr–Newman field (kerr-newman-ad.jl:104-146)."""
    r = np.asarray(r, np.float64)
    th = np.pi / 2
    dg = metric_dr(m, r, th)
    if q != 0.0:
        if not hasattr(m, "Q") or r.ndim:
            raise ValueError("charged circular orbits: scalar radius in a KerrNewmanMetric")
        omega = _charged_omega(m, float(r), dg, q, mu, contra_rotating)
    else:
        disc = np.sqrt(dg[4] ** 2 - dg[0] * dg[3])
        omega = -(dg[4] + disc) / dg[3] if contra_rotating else -(dg[4] - disc) / dg[3]
    gi = inverse_metric_components(metric_components(m, r, th))
    A = -(omega * gi[0] - gi[4])
    B = omega * gi[4] - gi[3]
    den = B * B * gi[0] + 2 * A * B * gi[4] + A * A * gi[3]
    d = -np.sign(den) * np.sqrt(1 / np.abs(den))
    ut_lo, uph_lo = B * d, A * d
    vt = gi[0] * ut_lo + gi[4] * uph_lo
    vph = gi[4] * ut_lo + gi[3] * uph_lo
    z = np.zeros_like(vt)
    return np.stack([vt, z, z, vph])


def _gram_schmidt(start, basis, G):
    v = start.astype(np.float64)
    for _ in range(2):  # second pass = re-orthogonalisation (the reference loops until the projection is below 4 eps)
        for e in basis:
            v = v - (v @ G @ e) / (e @ G @ e) * e
    return v / np.sqrt(abs(v @ G @ v))


def _searchsortedfirst(v, x):
    """Julia's binary search, 1-based result, also on the unsorted masks `tetradframe` feeds it."""
    lo, hi = 0, len(v) + 1
    while lo < hi - 1:
        mid = (lo + hi) >> 1
        if v[mid - 1] < x:
            lo = mid
        else:
            hi = mid
    return hi


def _permute(x):
    return (x[0], x[3], x[1], x[2])


def tetradframe(G, v):
    """Gram–Schmidt tetrad whose first leg is v/|v|, returned in coordinate order (t, r, θ, φ)
    (orthonormalization.jl:75-103).  G: 4×4 metric, v: 4-vector.  Like the reference's, the construction needs a
    velocity with at least one vanishing spatial component (its start vectors are the occupancy masks of v)."""
    G = np.asarray(G, np.float64)
    v = np.asarray(v, np.float64)
    v1 = v / np.sqrt(abs(v @ G @ v))
    state = tuple(bool(c != 0) for c in v1)
    if sum(state) == 1:
        state = (True, False, False, True)
    permutations = _searchsortedfirst(state[1:], True)
    v2 = _gram_schmidt(np.array(state, float), (v1,), G)
    state = tuple(a or b for a, b in zip(state, _permute(state)))
    v3 = _gram_schmidt(np.array(state, float), (v1, v2), G)
    state = tuple(a or b for a, b in zip(state, _permute(state)))
    v4 = _gram_schmidt(np.array(state, float), (v1, v2, v3), G)
    ret = (v1, v2, v3, v4)
    for _ in range(2, permutations + 1):
        ret = _permute(ret)
    return ret


def tetradframe_matrix(m, x, v):
    """Columns = tetrad legs (t, r, θ, φ)."""
    G = metric_matrix(metric_components(m, x[1], x[2]))
    return np.stack(tetradframe(G, v), axis=1)


def lnr_velocity_phi(m, r, theta, v):
    """𝒱^(φ) = v·e^(φ) / v·e^(t) in the locally non-rotating frame (Bardeen+73 3.9; flux-calculations.jl:13-18),
    closed form of the ZAMO co-basis: e^(t) = N dt, e^(φ) = √g_φφ (dφ − ω dt)."""
    g = metric_components(m, r, theta)
    omega = -g[4] / g[3]
    N = np.sqrt(-(g[0] - g[4] ** 2 / g[3]))  # lapse: −1/g^tt = −(g_tt − g_tφ²/g_φφ)
    return np.sqrt(g[3]) * (v[3] - omega * v[0]) / (N * v[0])


def lorentz_factor(m, r, theta, v):
    """flux-calculations.jl:32-35"""
    return 1.0 / np.sqrt(1.0 - lnr_velocity_phi(m, r, theta, v) ** 2)


def proper_area(m, r, theta):
    """`_proper_area`, corona/emissivity.jl:170-174: 2π √(g_rr g_φφ)"""
    g = metric_components(m, r, theta)
    return 2 * np.pi * np.sqrt(g[1] * g[3])


# --------------------------------------------------------------------------- a small host-side Tsit5 (set-up ODEs)
_TS_A = [[], [0.161], [-0.008480655492356989, 0.335480655492357], [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
         [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
         [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
         [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]]
_TS_BT = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552,
          -0.45808210592918697, 0.015151515151515152]


def tsit5_solve(f, u0, t_end, *, dtmax, abstol=1e-6, reltol=1e-3, terminate=None, maxiters=100000, return_times=False):
    """`solve(ODEProblem(f, u0, (0, t_end)), Tsit5(); dtmax, callback = DiscreteCallback(terminate, terminate!))` with
    OrdinaryDiffEq's defaults (abstol 1e-6, reltol 1e-3, PI controller, Hairer-Wanner initial step), every accepted step
    saved: the set-up ODEs of the reference that run on the host (the isobar of `PolishDoughnut`,
    src/geometry/discs/polish-doughnut.jl:66-100).  Returns the list of saved states (initial state first), with their times
    when `return_times` is set."""
    u = np.asarray(u0, np.float64)
    norm = lambda x: math.sqrt(float(np.mean(x * x)))  # noqa: E731  (ODE_DEFAULT_NORM)
    t, f0 = 0.0, np.asarray(f(u), np.float64)
    sk = abstol + np.abs(u) * reltol
    d0, d1 = norm(u / sk), norm(f0 / sk)
    dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    dt0 = min(dt0, dtmax)
    f1 = np.asarray(f(u + dt0 * f0), np.float64)
    d2 = norm((f1 - f0) / sk) / dt0
    md = max(d1, d2)
    dt1 = max(1e-6, dt0 * 1e-3) if md <= 1e-15 else 10.0 ** (-(2.0 + math.log10(md)) / 5.0)
    dt = min(100.0 * dt0, dt1, dtmax)
    beta1, beta2, gamma, qmin, qmax, qold = 7.0 / 50.0, 2.0 / 25.0, 0.9, 0.2, 10.0, 1e-4
    out, times, k1 = [u.copy()], [0.0], f0
    for _ in range(maxiters):
        dt = min(dt, t_end - t)
        ks = [k1]
        for s in range(1, 7):
            acc = sum(a * kk for a, kk in zip(_TS_A[s], ks))
            ks.append(np.asarray(f(u + dt * acc), np.float64))
        unew = u + dt * sum(a * kk for a, kk in zip(_TS_A[6], ks[:6]))
        err = dt * sum(b * kk for b, kk in zip(_TS_BT, ks))
        eest = norm(err / (abstol + np.maximum(np.abs(u), np.abs(unew)) * reltol))
        if eest == 0.0:
            q = 1.0 / qmax
        else:
            q11 = eest**beta1
            q = max(1.0 / qmax, min(1.0 / qmin, q11 / qold**beta2 / gamma))
        if eest > 1.0:
            dt = dt / min(1.0 / qmin, q11 / gamma)
            continue
        qold = max(eest, 1e-4)
        t += dt
        u, k1 = unew, ks[6]
        out.append(u.copy())
        times.append(t)
        if (terminate is not None and terminate(u)) or t >= t_end:
            break
        dt = min(dt / q, dtmax)
    return (times, out) if return_times else out
