"""Cunningham transfer functions on top of the device tracer.

What the reference does (src/transfer-functions/cunningham-transfer-functions.jl:336-426,
src/tracing/precision-solvers.jl:133-241, 401-451): for an emission radius rₑ it walks N image-plane angles θ, root-finds
the image-plane offset r(θ) whose geodesic lands on the disc at rₑ, takes the redshift g and the Jacobian
|∂(α,β)/∂(g,rₑ)| there, refines g_min / g_max with two golden-section searches and forms

    f = g √(g✶(1 − g✶)) |∂(α,β)/∂(g,rₑ)| / (π rₑ),      g✶ = (g − g_min) / (g_max − g_min).

Every one of those ~1300 geodesics per radius is an ordinary endpoint trace, i.e. the hot path this library puts on the
GPU.  The reference runs them one after the other (one radius per thread); here the *control* stays on the host, in the
reference's own order, but it is written in lock step over all (rₑ, θ) pairs: every Newton / golden-section iteration
is one `gb200_render` launch over all still-active pairs of all radii, so a table of 150 radii costs the same number of
launches as a single radius.  Derivatives: the reference pushes dual numbers through the integrator; this version takes
central differences of traces run at a tighter tolerance (rays are cheap here), which reproduces the reference's
transfer-function literals (test/smoke-tests/cunningham-transfer-functions.jl:25-39) well inside their tolerance.

The tracer is injected (`prober`): the product default is `DeviceProber` (C ABI, GPU, no fallback); the CPU test-suite
plugs the oracle into the same orchestration to pin the host logic without a GPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Callable, Optional, Sequence

import numpy as np

from . import api

_GOLDEN = 0.5 * (3.0 - math.sqrt(5.0))


@dataclass
class CunninghamTransferData:
    """src/transfer-functions/types.jl `CunninghamTransferData`: samples sorted by image-plane angle."""

    f: np.ndarray
    g_star: np.ndarray
    t: np.ndarray
    gmin: float
    gmax: float
    r_e: float
    theta: Optional[np.ndarray] = None  # the angles the samples were taken at (not stored by the reference)


class DeviceProber:
    """(α, β) → (g, ρ, t) at the disc through `gb200_render` with explicit impact-parameter lists.

    ρ = r sin θ at the intersection (`_equatorial_project`), NaN for rays that do not intersect."""

    def __init__(self, m, x, d, max_time=None, chart=None, ensemble=None, **solver_kwargs):
        self.m, self.x = m, np.asarray(x, np.float64)
        self.thick = None
        if isinstance(d, api.ThinDisc):  # `_promote_disc_for_transfer_functions`, cunningham-transfer-functions.jl:2-5
            d = api.DatumPlane(0.0)
        elif isinstance(d, api.ShakuraSunyaev):
            # thick disc: offsets are found on `datumplane(d, rₑ)` (one plane height per ray), visibility and the
            # Jacobian on the disc itself (`_rear_workhorse(::AbstractThickAccretionDisc)`, :274-333)
            self.thick = d
            d = api.DatumPlane(0.0)
        elif not isinstance(d, api.DatumPlane):
            raise ValueError("transfer functions on the device are implemented for thin discs, datum planes and ShakuraSunyaev")
        self.d = d
        self.max_time = 2 * self.x[1] if max_time is None else max_time
        self.chart = chart if chart is not None else api.chart_for_metric(m, 2 * self.x[1])
        self.ensemble = ensemble if ensemble is not None else api.EnsembleB200()
        self.solver_kwargs = solver_kwargs
        self.pfs = [api.ConstPointFunctions.redshift(m, x) @ api.ConstPointFunctions.filter_intersected(),
                    api.ConstPointFunctions.radius() @ api.ConstPointFunctions.filter_intersected(),
                    api.ConstPointFunctions.coordinate_time() @ api.ConstPointFunctions.filter_intersected(),
                    api.ConstPointFunctions.radius()]  # unfiltered: where a ray that missed the plane ended up
        self.plunging = None
        self.launches = 0
        self.rays = 0

    def config(self, alpha, beta, tol=None, height=None, thick=False, chart=None, callback=None):
        kw = dict(self.solver_kwargs)
        if tol is not None:
            kw["abstol"] = kw["reltol"] = tol
        if callback is not None:
            kw["callback"] = callback
        geometry = self.thick if thick else self.d
        return api.tracing_configuration(self.m, self.x, api.ImpactParameters(alpha, beta, None if thick else height),
                                         geometry, self.max_time, chart=self.chart if chart is None else chart,
                                         ensemble=self.ensemble, **kw)

    def _plunging(self):
        if self.plunging is None and not isinstance(self.m, api.KerrMetric):
            self.plunging = api.interpolate_plunging_velocities(self.m, self.ensemble)
        return self.plunging

    def evaluate(self, config):
        return api.apply_point_functions(config, self.pfs, plunging=self._plunging())

    def evaluate_points(self, config):
        return api.solve_tracing_problem(config)

    def cross_section(self, rho, group=None):
        return self.thick.cross_section(rho)

    def __call__(self, alpha, beta, tol=None, height=None, thick=False, callback=None, with_end_radius=False, group=None):
        """(g, ρ, t) at the intersection: with the datum plane(s) (one height per ray if `height` is given), or with the
        thick disc itself (`thick=True`, the Jacobian's traces).  `with_end_radius` adds r|sin θ| of the end point
        whatever its status (the reference's root finder reads that for rays that missed, precision-solvers.jl:124)."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        beta = np.ascontiguousarray(beta, np.float64)
        if alpha.size == 0:
            z = np.zeros(0)
            return (z, z, z, z) if with_end_radius else (z, z, z)
        self.launches += 1
        self.rays += alpha.size
        out = self.evaluate(self.config(alpha, beta, tol, height=height, thick=thick, callback=callback))
        return (out[0], out[1], out[2], out[3]) if with_end_radius else (out[0], out[1], out[2])

    def points(self, alpha, beta, height=None, thick=False, default_chart=False, group=None):
        """`GeodesicPoint`s of the same rays (status, λ_max, x): what the thick-disc visibility test compares."""
        alpha = np.ascontiguousarray(alpha, np.float64)
        beta = np.ascontiguousarray(beta, np.float64)
        self.launches += 1
        self.rays += alpha.size
        chart = api.chart_for_metric(self.m) if default_chart else None
        return self.evaluate_points(self.config(alpha, beta, height=height, thick=thick, chart=chart))


class CellProber:
    """Several probers — one per (metric, observer) cell of a transfer-function table — behind the prober interface.
    Every call carries `group`, the cell of each ray; the rays of all cells go out in ONE `gb200_render_batch` call
    (one launch per cell on the stream pool), so a probe round of the lock-step orchestration costs one host round trip
    for the whole table instead of one per cell (`make_transfer_function_table`, cunningham-transfer-functions.jl:507-530,
    runs its cells one after the other)."""

    def __init__(self, probers):
        self.probers = list(probers)
        self.thick = self.probers[0].thick
        self.launches = 0
        self.rays = 0

    def cross_section(self, rho, group):
        out = np.empty(len(rho))
        for c in np.unique(group):
            sel = group == c
            out[sel] = self.probers[c].thick.cross_section(rho[sel])
        return out

    def evaluate_batch(self, configs, cells):
        plungings = [self.probers[c]._plunging() for c in cells]
        return api.apply_point_functions_batch(configs, self.probers[0].pfs, plungings)

    def __call__(self, alpha, beta, tol=None, height=None, thick=False, callback=None, with_end_radius=False, group=None):
        alpha = np.ascontiguousarray(alpha, np.float64)
        beta = np.ascontiguousarray(beta, np.float64)
        n = alpha.size
        out = np.full((4, n), np.nan)
        if n:
            cells, sels = _split_by_group(group)
            configs = [self.probers[c].config(alpha[s], beta[s], tol, height=None if height is None else height[s], thick=thick,
                                              callback=callback) for c, s in zip(cells, sels)]
            self.launches += 1
            self.rays += n
            for s, img in zip(sels, self.evaluate_batch(configs, cells)):
                out[:, s] = img
        return (out[0], out[1], out[2], out[3]) if with_end_radius else (out[0], out[1], out[2])

    def points(self, alpha, beta, height=None, thick=False, default_chart=False, group=None):
        """Per cell (two calls per sample set, off the probe-round path)."""
        alpha = np.asarray(alpha, np.float64)
        n = alpha.size
        status = np.zeros(n, np.int32)
        lam = np.zeros(n)
        x = np.zeros((4, n))
        for c, s in zip(*_split_by_group(group)):
            gp = self.probers[c].points(alpha[s], np.asarray(beta)[s], height=None if height is None else height[s], thick=thick,
                                        default_chart=default_chart)
            status[s], lam[s], x[:, s] = gp.status, gp.lambda_max, gp.x
        return _Points(status, lam, x)


def _split_by_group(group):
    """cells present in `group` and, per cell, the (ascending) indices of its rays: one stable sort."""
    group = np.asarray(group)
    order = np.argsort(group, kind="stable")
    sorted_g = group[order]
    cuts = np.nonzero(np.diff(sorted_g))[0] + 1
    return sorted_g[np.concatenate([[0], cuts])] if group.size else np.zeros(0, int), np.split(order, cuts)


@dataclass
class _Points:
    status: np.ndarray
    lambda_max: np.ndarray
    x: np.ndarray


@dataclass
class TransferFunctionSetup:
    """`_TransferFunctionSetup`, cunningham-transfer-functions.jl:7-58."""

    theta_offset: float = 0.3
    zero_atol: float = 1e-7
    N: int = 80
    N_extrema: int = 17
    h: float = 1e-6
    max_iter: int = 50
    # central-difference Jacobian (the reference uses dual numbers): relative step on the image plane and the
    # integrator tolerance of the differenced traces
    fd_step: float = 2e-5
    fd_tol: float = 1e-12
    # origin of the polar coordinates on the image plane (`_rθ_to_αβ`, precision-solvers.jl:1-7)
    alpha0: float = 0.0
    beta0: float = 0.0


def theta_samples(setup: TransferFunctionSetup) -> np.ndarray:
    """The N fixed angles, clustered where g has its extrema (cunningham-transfer-functions.jl:366-370)."""
    K = setup.N // 5
    o = 2 * setup.theta_offset
    return np.concatenate([np.linspace(-o, o, K), np.linspace(-math.pi / 2, 3 * math.pi / 2, setup.N - 2 * K),
                           np.linspace(math.pi - o, math.pi + o, K)])


def find_offset_for_radius(prober, r_target, theta, setup: TransferFunctionSetup = TransferFunctionSetup(), initial_r=None,
                           height=None, group=None):
    """Batched `_find_offset_for_radius` (precision-solvers.jl:133-241): for each pair (r_target[i], theta[i]) the
    image-plane offset r with ρ(r cos θ, r sin θ) = r_target.  Newton steps on ρ(r) − r_target, safeguarded by the
    bracket the monotonicity of ρ(r) provides (the lower end starts inside the hole, like the reference's contrapoint).
    One launch per iteration for all unconverged pairs.  Returns (r, g, t); r is NaN where no offset was found
    (|ρ − r_target| > 1e-4 r_target after `max_iter` iterations)."""
    r_target = np.asarray(r_target, np.float64)
    theta = np.asarray(theta, np.float64)
    n = r_target.size
    ct, st = np.cos(theta), np.sin(theta)
    r = np.maximum(20.0, r_target) if initial_r is None else np.array(initial_r, np.float64)
    lo = np.zeros(n)
    hi = np.full(n, np.inf)
    g = np.full(n, np.nan)
    t = np.full(n, np.nan)
    y = np.full(n, np.inf)
    best = r.copy()  # the evaluated offset with the smallest |ρ − r_target| so far
    stall = np.zeros(n, int)
    active = np.ones(n, bool)
    rel = 1e-6  # forward-difference step for dρ/dr
    for _ in range(setup.max_iter + 1):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        ra = r[idx]
        rb = ra * (1 + rel)
        hq = {} if height is None else {"height": np.concatenate([height[idx], height[idx]])}
        if group is not None:
            hq["group"] = np.concatenate([group[idx], group[idx]])
        gq, rho, tq, rho_end = prober(np.concatenate([ra * ct[idx], rb * ct[idx]]) + setup.alpha0,
                                      np.concatenate([ra * st[idx], rb * st[idx]]) + setup.beta0, with_end_radius=True, **hq)
        m = idx.size
        ya = rho[:m] - r_target[idx]
        yb = rho[m:] - r_target[idx]
        # no intersection: the ray fell into the hole (its end point projects inside the target: the root lies further
        # out) or left the domain before reaching the plane (end point far outside: the root lies further in)
        lost = ~np.isfinite(ya)
        ya = np.where(lost, np.where(rho_end[:m] < r_target[idx], -np.inf, np.inf), ya)
        for yc, rc, off in ((ya, ra, 0), (np.where(np.isfinite(yb), yb, -np.inf), rb, m)):  # both traces are candidates
            improved = np.abs(yc) < np.abs(y[idx])
            k = idx[improved]
            y[k], best[k], g[k], t[k] = yc[improved], rc[improved], gq[off:off + m][improved], tq[off:off + m][improved]
            stall[idx] = np.where(improved, 0, stall[idx] + (off > 0))
        # converged, or stuck at the resolution of ρ(r) (the integrator's own error, ~reltol·ρ, can exceed zero_atol
        # at large radii): three iterations without improvement on an already good offset
        done = (np.abs(y[idx]) <= setup.zero_atol) | ((stall[idx] >= 3) & (np.abs(y[idx]) <= 1e-6 * r_target[idx]))
        active[idx[done]] = False
        below = ya < 0
        lo[idx] = np.where(below, np.maximum(lo[idx], ra), lo[idx])
        hi[idx] = np.where(~below, np.minimum(hi[idx], ra), hi[idx])
        with np.errstate(all="ignore"):
            dy = (yb - ya) / (rb - ra)
            nxt = ra - ya / dy
        # outside the bracket (or no derivative): bisect it, or double while there is no upper end yet
        bad = ~np.isfinite(nxt) | (nxt <= lo[idx]) | (nxt >= hi[idx])
        with np.errstate(all="ignore"):
            bis = np.where(np.isfinite(hi[idx]), 0.5 * (lo[idx] + hi[idx]), 2.0 * ra)
        nxt = np.where(bad, bis, nxt)
        r[idx] = np.where(done, ra, nxt)
    poor = ~(np.abs(y) <= 1e-4 * r_target)
    return np.where(poor, np.nan, best), g, t


def jacobian_ab_gr(prober, alpha, beta, setup: TransferFunctionSetup = TransferFunctionSetup(), thick=False, group=None):
    """|∂(ρ, g)/∂(α, β)|⁻¹ (`jacobian_∂αβ_∂gr`, precision-solvers.jl:401-451) by central differences: four traces
    per point at tolerance `fd_tol`, all points in one launch."""
    alpha = np.asarray(alpha, np.float64)
    beta = np.asarray(beta, np.float64)
    n = alpha.size
    h = setup.fd_step * np.maximum(np.hypot(alpha, beta), 1.0)
    a = np.concatenate([alpha + h, alpha - h, alpha, alpha])
    b = np.concatenate([beta, beta, beta + h, beta - h])
    gq = {} if group is None else {"group": np.tile(group, 4)}
    if thick:  # on the disc itself, upper hemisphere only (precision-solvers.jl:411,424-426)
        g, rho, _ = prober(a, b, setup.fd_tol, thick=True, callback=api.domain_upper_hemisphere(), **gq)
    else:
        g, rho, _ = prober(a, b, setup.fd_tol, **gq)
    with np.errstate(all="ignore"):
        drho_da = (rho[:n] - rho[n:2 * n]) / (2 * h)
        drho_db = (rho[2 * n:3 * n] - rho[3 * n:]) / (2 * h)
        dg_da = (g[:n] - g[n:2 * n]) / (2 * h)
        dg_db = (g[2 * n:3 * n] - g[3 * n:]) / (2 * h)
        return np.abs(1.0 / (drho_da * dg_db - drho_db * dg_da))


class _Workhorse:
    """`_rear_workhorse` for thin discs (cunningham-transfer-functions.jl:253-272): θ → (g, J, t), batched."""

    def __init__(self, prober, setup):
        self.prober, self.setup = prober, setup

    def __call__(self, r_e, theta, group=None):
        r, g, t = find_offset_for_radius(self.prober, r_e, theta, self.setup, group=group)
        if np.any(np.isnan(r)):
            k = int(np.nonzero(np.isnan(r))[0][0])
            raise RuntimeError(f"Transfer function integration failed (rₑ={r_e[k]}, θ={theta[k]}).")
        J = jacobian_ab_gr(self.prober, r * np.cos(theta) + self.setup.alpha0, r * np.sin(theta) + self.setup.beta0, self.setup,
                           group=group)
        return g, J, t


class _ThickWorkhorse:
    """`_thick_workhorse` (cunningham-transfer-functions.jl:274-333), batched: the offset is found on the datum plane
    through the disc surface at rₑ, the same ray is then traced against the disc itself, and the sample only counts
    (finite J) if it ends in the same way at (nearly) the same place, i.e. if that patch of the surface is visible."""

    def __init__(self, prober, setup):
        self.prober, self.setup, self.disc = prober, setup, prober.thick

    def __call__(self, r_e, theta, group=None):
        r_e = np.asarray(r_e, np.float64)
        gq = {} if group is None else {"group": group}
        h = self.prober.cross_section(r_e, group)
        r, g, t = find_offset_for_radius(self.prober, r_e, theta, self.setup, height=h, group=group)
        if np.any(np.isnan(r)):
            k = int(np.nonzero(np.isnan(r))[0][0])
            raise RuntimeError(f"Transfer function integration failed (rₑ={r_e[k]}, θ={theta[k]}).")
        alpha = r * np.cos(theta) + self.setup.alpha0
        beta = r * np.sin(theta) + self.setup.beta0
        gp = self.prober.points(alpha, beta, height=h, **gq)
        # the reference re-traces with the default chart and stops at 1.1 λ_max of the datum-plane point: a disc hit
        # later than that is no hit
        gt = self.prober.points(alpha, beta, thick=True, default_chart=True, **gq)
        status = np.where((gt.status == api.StatusCodes.IntersectedWithGeometry) & (gt.lambda_max > 1.1 * gp.lambda_max),
                          api.StatusCodes.NoStatus, gt.status)
        dist = np.linalg.norm(gp.x - gt.x, axis=0)
        close = dist <= 1e-3 * np.maximum(np.linalg.norm(gp.x, axis=0), np.linalg.norm(gt.x, axis=0))  # isapprox(rtol = 1e-3)
        ok = (status == gp.status) & close
        J = np.full(r.size, np.nan)
        if ok.any():
            J[ok] = jacobian_ab_gr(self.prober, alpha[ok], beta[ok], self.setup, thick=True, group=None if group is None else group[ok])
        J[~np.isfinite(J)] = np.nan  # `is_visible = isfinite(J)`; invisible samples keep g and t, J = NaN (utils.jl:71-78)
        return g, J, t


def _golden_sections(fn, lower, upper, iterations, rel_tol=math.sqrt(np.finfo(float).eps), abs_tol=np.finfo(float).eps):
    """Optim.jl's `GoldenSection` univariate minimiser (the algorithm `_search_extremal!` calls with
    `iterations = N_extrema − 1`, cunningham-transfer-functions.jl:391-426), run in lock step for a batch of
    independent problems: `fn(x, mask)` evaluates the objective of every problem with mask set.  1 + iterations
    evaluations per problem.  Returns the best objective value found per problem."""
    lower = np.array(lower, np.float64)
    upper = np.array(upper, np.float64)
    xm = lower + _GOLDEN * (upper - lower)
    active = np.ones(xm.size, bool)
    fm = fn(xm, active)
    for _ in range(iterations):
        tolx = rel_tol * np.abs(xm) + abs_tol
        mid = 0.5 * (upper + lower)
        active &= ~(np.abs(xm - mid) <= 2 * tolx - 0.5 * (upper - lower))
        if not active.any():
            break
        right = (upper - xm) > (xm - lower)
        xn = np.where(right, xm + _GOLDEN * (upper - xm), xm - _GOLDEN * (xm - lower))
        fnew = fn(xn, active)
        better = fnew < fm
        a = active
        new_lower = np.where(right, np.where(better, xm, lower), np.where(better, lower, xn))
        new_upper = np.where(right, np.where(better, upper, xn), np.where(better, xm, upper))
        lower = np.where(a, new_lower, lower)
        upper = np.where(a, new_upper, upper)
        upd = a & better
        xm = np.where(upd, xn, xm)
        fm = np.where(upd, fnew, fm)
    return fm


def cunningham_transfer_functions(m, x, d, radii: Sequence[float], *, prober: Optional[Callable] = None,
                                  setup: Optional[TransferFunctionSetup] = None, chart=None, max_time=None, ensemble=None,
                                  groups=None, **kwargs) -> list:
    """`cunningham_transfer_function(m, x, d, rₑ; N, chart, max_time, ...)` for every rₑ in `radii` at once (the loop
    `interpolated_transfer_branches` threads over, cunningham-transfer-functions.jl:428-462)."""
    setup_keys = {"theta_offset", "zero_atol", "N", "N_extrema", "h", "max_iter", "fd_step", "fd_tol", "alpha0", "beta0"}
    if setup is None:
        alias = {"θ_offset": "theta_offset", "α₀": "alpha0", "β₀": "beta0"}
        skw = {alias.get(k, k): kwargs.pop(k) for k in list(kwargs) if alias.get(k, k) in setup_keys}
        setup = TransferFunctionSetup(**skw)
    if prober is None:
        prober = DeviceProber(m, x, d, max_time=max_time, chart=chart, ensemble=ensemble, **kwargs)
    radii = np.atleast_1d(np.asarray(radii, np.float64))
    R = radii.size
    groups = None if groups is None else np.asarray(groups, np.int64)  # cell of each radius (CellProber)
    gq = (lambda sel: {}) if groups is None else (lambda sel: {"group": sel})
    work = _ThickWorkhorse(prober, setup) if getattr(prober, "thick", None) is not None else _Workhorse(prober, setup)
    th0 = theta_samples(setup)
    N = th0.size
    M = N + 2 * setup.N_extrema
    thetas = np.full((R, M), np.nan)
    gs = np.full((R, M), np.nan)
    Js = np.full((R, M), np.nan)
    ts = np.full((R, M), np.nan)
    g, J, t = work(np.repeat(radii, N), np.tile(th0, R), **gq(None if groups is None else np.repeat(groups, N)))
    thetas[:, :N] = th0
    gs[:, :N], Js[:, :N], ts[:, :N] = g.reshape(R, N), J.reshape(R, N), t.reshape(R, N)

    # `_search_extremal!`: problems 0..R-1 minimise g near θ = 0, R..2R-1 maximise near θ = π; every probe is kept
    fill = np.full(2 * R, 0)
    sign = np.concatenate([np.ones(R), -np.ones(R)])
    re2 = np.concatenate([radii, radii])
    grp2 = None if groups is None else np.concatenate([groups, groups])

    def objective(theta, mask):
        idx = np.nonzero(mask)[0]
        th = theta[idx].copy()
        pole = (np.abs(th) < 1e-4) | (np.abs(np.abs(th) - math.pi) < 1e-4)
        th = np.where(pole, th + 1e-4, th)
        gq_, Jq, tq = work(re2[idx], th, **gq(None if grp2 is None else grp2[idx]))
        out = np.full(theta.size, np.inf)
        rr = idx % R
        col = N + np.where(idx < R, 0, setup.N_extrema) + fill[idx]
        thetas[rr, col], gs[rr, col], Js[rr, col], ts[rr, col] = th, gq_, Jq, tq
        fill[idx] += 1
        out[idx] = sign[idx] * gq_
        return out

    off = setup.theta_offset
    best = _golden_sections(objective, np.concatenate([np.full(R, -off), np.full(R, math.pi - off)]),
                            np.concatenate([np.full(R, off), np.full(R, math.pi + off)]), setup.N_extrema - 1)
    out = []
    for k in range(R):
        used = np.isfinite(thetas[k])
        th, gk, Jk, tk = thetas[k, used], gs[k, used], Js[k, used], ts[k, used]
        gmin, gmax = _check_gmin_gmax(best[k], -best[R + k], radii[k], gk)
        order = np.argsort(th, kind="stable")
        th, gk, Jk, tk = th[order], gk[order], Jk[order], tk[order]
        Jk = (gmax - gmin) * Jk  # ∂g → ∂g✶
        gstar = (gk - gmin) / (gmax - gmin)
        with np.errstate(invalid="ignore"):
            f = (1.0 / (math.pi * radii[k])) * gk * np.sqrt(gstar * (1.0 - gstar)) * Jk
        out.append(CunninghamTransferData(f, gstar, tk, float(gmin), float(gmax), float(radii[k]), th))
    return out


def _check_gmin_gmax(gmin, gmax, r_e, gs):
    """transfer-functions/utils.jl:80-112"""
    if np.isnan(gmin):
        gmin = np.min(gs)
    if np.isnan(gmax):
        gmax = np.max(gs)
    if gmin == gmax:
        gmin, gmax = np.min(gs), np.max(gs)
        if gmin == gmax:
            raise RuntimeError(f"Cannot use extrema (rₑ = {r_e})")
    return min(gmin, np.min(gs)), max(gmax, np.max(gs))


def cunningham_transfer_function(m, x, d, r_e: float, **kwargs) -> CunninghamTransferData:
    """Single-radius form, cunningham-transfer-functions.jl:336-389."""
    return cunningham_transfer_functions(m, x, d, [r_e], **kwargs)[0]


def measure_ctf(ctf: CunninghamTransferData) -> float:
    """The scalar the reference's smoke test pins (test/smoke-tests/cunningham-transfer-functions.jl:19-21)."""
    return float(np.sum(ctf.f * ctf.g_star) / len(ctf.f))


def transfer_function_table(metrics, observers, d, radii_of, *, ensemble=None, setup: Optional[TransferFunctionSetup] = None,
                            prober_cls=DeviceProber, **kwargs) -> list:
    """Cunningham transfer functions of many (metric, observer) cells in one lock-step computation: the device-side
    form of `make_transfer_function_table` (cunningham-transfer-functions.jl:507-530), which loops `for a in a_range,
    θ in θ_range` and computes one cell after the other.  `metrics[c]`, `observers[c]` describe cell c, `d` is the disc
    (or a callable m -> disc, for discs that depend on the metric), `radii_of(m)` the emission radii of a cell.
    Returns, per cell, the list of `CunninghamTransferData` of its radii."""
    ensemble = ensemble if ensemble is not None else api.EnsembleB200()
    probers, radii, groups = [], [], []
    for c, (m, x) in enumerate(zip(metrics, observers)):
        disc = d(m) if callable(d) else d
        probers.append(prober_cls(m, x, disc, ensemble=ensemble, **kwargs))
        r = np.atleast_1d(np.asarray(radii_of(m), np.float64))
        radii.append(r)
        groups.append(np.full(r.size, c))
    cell = CellProber(probers)
    ctfs = cunningham_transfer_functions(None, None, None, np.concatenate(radii), prober=cell, setup=setup or TransferFunctionSetup(),
                                         groups=np.concatenate(groups))
    out, k = [], 0
    for r in radii:
        out.append(ctfs[k:k + r.size])
        k += r.size
    return out
