"""Cunningham transfer functions on top of the device tracer.

What the reference does (src/transfer-functions/cunningham-transfer-functions.jl:336-426,
src/tracing/precision-solvers.jl:73-241, 401-451): for an emission radius rₑ it walks N image-plane angles θ, root-finds
the image-plane offset r(θ) whose geodesic lands on the disc at rₑ (Newton on ρ(r) − rₑ with dρ/dr from a dual number
pushed through the integrator), takes the redshift g and the Jacobian |∂(α,β)/∂(g,rₑ)| there (two partials pushed
through the integrator), refines g_min / g_max with two golden-section searches and forms

    f = g √(g✶(1 − g✶)) |∂(α,β)/∂(g,rₑ)| / (π rₑ),      g✶ = (g − g_min) / (g_max − g_min).

Every one of those geodesics is a forward-mode trace on the device (`gb200_trace_dual`: the integrator state carries the
partials, the step-size control sees them through the error norm as DiffEqBase's does, the event time moves with the
parameters).  The reference runs them one after the other (one radius per thread); here the *control* stays on the host,
in the reference's own order and with its own update rules, but it is written in lock step over all (rₑ, θ) pairs: every
Newton / golden-section iteration is one launch over all still-active pairs of all radii, so a table of 150 radii costs
the same number of launches as a single radius.

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
        self.ensemble = ensemble if ensemble is not None else api.default_ensemble()
        self.solver_kwargs = solver_kwargs
        self.pfs = [api.ConstPointFunctions.redshift(m, x) @ api.ConstPointFunctions.filter_intersected(),
                    api.ConstPointFunctions.radius() @ api.ConstPointFunctions.filter_intersected(),
                    api.ConstPointFunctions.coordinate_time() @ api.ConstPointFunctions.filter_intersected(),
                    api.ConstPointFunctions.radius()]  # unfiltered: where a ray that missed the plane ended up
        self.plunging = None
        self.launches = 0
        self.rays = 0
        self.norm_mode = api.cabi.DUAL_NORM_WITH_PARTIALS  # DiffEqBase's norm on dual-valued states

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

    def evaluate_dual(self, config, arrays, norm_mode):
        return api.trace_dual(config, arrays, norm_mode, plunging=self._plunging())

    def cross_section(self, rho, group=None):
        return self.thick.cross_section(rho)

    def disc_inner_radius(self, group=None):
        return self.thick.inner_radius

    def dual(self, alpha, beta, dalpha, dbeta, height=None, thick=False, callback=None, group=None):
        """Forward-mode trace of the rays (α, β) whose partials are (dalpha[k], dbeta[k]), k < 1 or 2: returns the filled
        `DualArrays` (g, dg, rho, drho, x, status, ...).  ρ = r sin θ of the end point whatever its status."""
        arrays = api.cabi.DualArrays(alpha, beta, dalpha, dbeta, None if thick else height)
        if arrays.n == 0:
            return arrays
        self.launches += 1
        self.rays += arrays.n
        cfg = self.config(arrays.alpha, arrays.beta, None, height=None, thick=thick, callback=callback)
        return self.evaluate_dual(cfg, arrays, self.norm_mode)

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

    def disc_inner_radius(self, group):
        return np.array([p.thick.inner_radius for p in self.probers])[np.asarray(group)]

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

    def evaluate_dual_batch(self, configs, arrays, cells):
        plungings = [self.probers[c]._plunging() for c in cells]
        return api.trace_dual_batch(configs, arrays, self.probers[0].norm_mode, plungings)

    def dual(self, alpha, beta, dalpha, dbeta, height=None, thick=False, callback=None, group=None):
        """All cells' rays in ONE `gb200_trace_dual_batch` call; returns one merged `DualArrays`-like result."""
        merged = api.cabi.DualArrays(alpha, beta, dalpha, dbeta, None if thick else height)
        if merged.n == 0:
            return merged
        cells, sels = _split_by_group(group)
        parts = [api.cabi.DualArrays(merged.alpha[s], merged.beta[s], merged.dalpha[:, s], merged.dbeta[:, s],
                                     None if (thick or height is None) else merged.height[s]) for s in sels]
        configs = [self.probers[c].config(pa.alpha, pa.beta, None, height=None, thick=thick, callback=callback) for c, pa in zip(cells, parts)]
        self.launches += 1
        self.rays += merged.n
        self.evaluate_dual_batch(configs, parts, cells)
        for s, pa in zip(sels, parts):
            merged.status[s], merged.lambda_max[s] = pa.status, pa.lambda_max
            merged.x[:, s], merged.v[:, s] = pa.x, pa.v
            merged.g[s], merged.rho[s] = pa.g, pa.rho
            merged.dg[:, s], merged.drho[:, s] = pa.dg, pa.drho
            merged.naccept[s], merged.nreject[s], merged.flags[s] = pa.naccept, pa.nreject, pa.flags
        return merged

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
    """`_TransferFunctionSetup`, cunningham-transfer-functions.jl:7-58, and the keyword defaults of
    `_find_offset_for_radius`, precision-solvers.jl:133-147."""

    theta_offset: float = 0.3
    zero_atol: float = 1e-7
    N: int = 80
    N_extrema: int = 17
    h: float = 1e-6
    max_iter: int = 50
    contrapoint_bias: float = 2.0
    # origin of the polar coordinates on the image plane (`_rθ_to_αβ`, precision-solvers.jl:1-7)
    alpha0: float = 0.0
    beta0: float = 0.0


def theta_samples(setup: TransferFunctionSetup) -> np.ndarray:
    """The N fixed angles, clustered where g has its extrema (cunningham-transfer-functions.jl:366-370)."""
    K = setup.N // 5
    o = 2 * setup.theta_offset
    return np.concatenate([np.linspace(-o, o, K), np.linspace(-math.pi / 2, 3 * math.pi / 2, setup.N - 2 * K),
                           np.linspace(math.pi - o, math.pi + o, K)])


def _bracket_offsets(step_y, lo, hi, atol, max_halvings=80):
    """`find_zero(f, (contra, x), atol = zero_atol)` (Roots.jl bisection on a bracketing interval) for a batch of
    independent intervals: `step_y(idx, x)` evaluates ρ(x) − rₑ for the problems `idx`.  Returns the roots, NaN where the
    end points do not bracket a sign change (Roots.jl raises there)."""
    n = lo.size
    idx = np.arange(n)
    flo, fhi = step_y(idx, lo), step_y(idx, hi)
    root = np.where(np.abs(flo) <= np.abs(fhi), lo, hi).astype(np.float64)
    best = np.minimum(np.abs(flo), np.abs(fhi))
    ok = np.sign(flo) * np.sign(fhi) < 0
    root[~ok & (best > atol)] = np.nan
    active = ok & (best > atol)
    lo, hi, flo = lo.copy(), hi.copy(), flo.copy()
    for _ in range(max_halvings):
        k = np.nonzero(active)[0]
        if k.size == 0:
            break
        mid = 0.5 * (lo[k] + hi[k])
        fm = step_y(k, mid)
        better = np.abs(fm) < best[k]
        root[k[better]], best[k[better]] = mid[better], np.abs(fm[better])
        same = np.sign(fm) == np.sign(flo[k])
        lo[k] = np.where(same, mid, lo[k])
        flo[k] = np.where(same, fm, flo[k])
        hi[k] = np.where(same, hi[k], mid)
        active[k] = (np.abs(fm) > atol) & (hi[k] - lo[k] > 4 * np.finfo(float).eps * np.abs(mid))
    return root


def find_offset_for_radius(prober, r_target, theta, setup: TransferFunctionSetup = TransferFunctionSetup(), initial_r=None,
                           height=None, group=None, r_min=None):
    """`_find_offset_for_radius` (precision-solvers.jl:133-241) for a batch of (r_target[i], theta[i]) pairs: the
    image-plane offset x with ρ(x cos θ, x sin θ) = r_target.  The reference's iteration, pair by pair -- Newton steps
    x − y / y′ with y′ = dρ/dx read off a dual number pushed through the trace, a contrapoint inside the hole that pulls
    overshoots back (`contrapoint_bias`), cycle detection on the relative decrease with a bracketing finish -- run in
    lock step: every `step` of every still-active pair goes out in one launch.  Returns (x, point) with point the
    `DualArrays`-like record (g, rho, x, status, ...) of the last trace of each pair; x is NaN where the reference
    returns NaN (negative offset, or |y| > 1e-4 r_target at the end)."""
    r_target = np.asarray(r_target, np.float64)
    theta = np.asarray(theta, np.float64)
    n = r_target.size
    ct, st = np.cos(theta), np.sin(theta)
    atol, bias = setup.zero_atol, setup.contrapoint_bias
    pt = dict(g=np.full(n, np.nan), rho=np.full(n, np.nan), t=np.full(n, np.nan), status=np.full(n, -1, np.int32),
              lambda_max=np.full(n, np.nan), x=np.full((4, n), np.nan), alpha=np.full(n, np.nan), beta=np.full(n, np.nan))
    df = np.zeros(n)

    def step(idx, xr):
        """(point, df, y) of the reference's `step`, for the pairs idx at offsets xr; records point and df."""
        al, be = xr * ct[idx] + setup.alpha0, xr * st[idx] + setup.beta0
        kw = {}
        if height is not None:
            kw["height"] = height[idx]
        if group is not None:
            kw["group"] = group[idx]
        res = prober.dual(al, be, ct[idx][None, :], st[idx][None, :], **kw)
        pt["g"][idx], pt["rho"][idx], pt["t"][idx], pt["status"][idx] = res.g, res.rho, res.x[0], res.status
        pt["lambda_max"][idx], pt["x"][:, idx], pt["alpha"][idx], pt["beta"][idx] = res.lambda_max, res.x, al, be
        df[idx] = res.drho[0]
        return res.rho - r_target[idx]

    if r_min is None:
        r_min = api.inner_radius(prober.m) if getattr(prober, "m", None) is not None else 0.0
    r_min = np.broadcast_to(np.asarray(r_min, np.float64), (n,))
    x = np.maximum(20.0, r_target) if initial_r is None else np.array(np.broadcast_to(initial_r, (n,)), np.float64)
    contra = np.zeros(n)
    allidx = np.arange(n)
    y = step(allidx, x)
    dy = np.zeros(n)
    previous = np.zeros((n, 6))
    it = np.zeros(n, int)
    active = ~(np.abs(y) <= atol)
    while True:
        A = np.nonzero(active & (it <= setup.max_iter))[0]
        if A.size == 0:
            break
        with np.errstate(all="ignore"):
            next_x = x[A] - y[A] / df[A]
        next_y = step(A, next_x)
        pull = (next_x < 0) | ((next_y < 0) & (y[A] > 0))
        contra[A] = np.where(pull, np.maximum(contra[A], next_x), contra[A])
        # the overshoot ended in (or next to) the hole: step back towards the contrapoint instead
        redo = pull & ((next_x < 0) | (pt["rho"][A] < r_min[A] + 1))
        if redo.any():
            R = A[redo]
            next_x[redo] = (contra[R] * bias + x[R]) / (1 + bias)
            next_y[redo] = step(R, next_x[redo])
        with np.errstate(all="ignore"):
            failed = (next_y < 0) & (y[A] < 0) & ((-y[A] / df[A]) < 0)  # "Converge failed": x, y keep their old values
            next_dy = (y[A] - next_y) / y[A]
            cycle = ~failed & (y[A] > 0) & np.any(np.abs(next_dy[:, None] - previous[A]) <= atol * 100, axis=1)
        active[A[failed]] = False
        if cycle.any():  # stuck with Newton-Raphson: finish off by bracketing between the contrapoint and x
            Cy = A[cycle]
            x[Cy] = _bracket_offsets(lambda k, xr: step(Cy[k], xr), contra[Cy], x[Cy], atol)
            good = np.isfinite(x[Cy])
            y[Cy[good]] = step(Cy[good], x[Cy[good]])
            y[Cy[~good]] = np.inf
            active[Cy] = False
        go = ~failed & ~cycle
        G = A[go]
        x[G], dy[G], y[G] = next_x[go], next_dy[go], next_y[go]
        previous[G, it[G] % 6] = dy[G]
        it[G] += 1
        active[G] = ~(np.abs(y[G]) <= atol)
    # exceeded max_iter with a large residual: "Attempting to bracket"
    late = np.nonzero((it >= setup.max_iter) & (y > 10.0) & np.isfinite(x))[0]
    if late.size:
        x[late] = _bracket_offsets(lambda k, xr: step(late[k], xr), contra[late], x[late], atol)
        good = np.isfinite(x[late])
        y[late[good]] = step(late[good], x[late[good]])
        y[late[~good]] = np.inf
    with np.errstate(invalid="ignore"):
        poor = ~np.isfinite(x) | (x < 0) | ~(np.abs(y) <= 1e-4 * r_target)
    return np.where(poor, np.nan, x), pt


def jacobian_ab_gr(prober, alpha, beta, setup: TransferFunctionSetup = TransferFunctionSetup(), thick=False, group=None,
                   disc_inner_radius=None):
    """|∂(ρ, g)/∂(α, β)|⁻¹ (`jacobian_∂αβ_∂gr`, precision-solvers.jl:401-451): one forward-mode trace per point with the two
    partials of (α, β), all points in one launch.  On a thick disc the trace runs against the disc itself under
    `domain_upper_hemisphere`, and g is set to zero inside the disc's inner radius (:431-435), which makes the inverse
    Jacobian infinite: the visibility flag of `_thick_workhorse`."""
    alpha = np.asarray(alpha, np.float64)
    beta = np.asarray(beta, np.float64)
    n = alpha.size
    one, zero = np.ones(n), np.zeros(n)
    kw = {} if group is None else {"group": group}
    if thick:
        kw.update(thick=True, callback=api.domain_upper_hemisphere())
    res = prober.dual(alpha, beta, np.stack([one, zero]), np.stack([zero, one]), **kw)
    dg = res.dg.copy()
    if thick and disc_inner_radius is not None:
        dg[:, res.rho < disc_inner_radius] = 0.0
    with np.errstate(all="ignore"):
        det = res.drho[0] * dg[1] - res.drho[1] * dg[0]
        return np.abs(1.0 / det)


class _Workhorse:
    """`_rear_workhorse` for thin discs (cunningham-transfer-functions.jl:253-272): θ → (g, J, t), batched."""

    def __init__(self, prober, setup):
        self.prober, self.setup = prober, setup

    def __call__(self, r_e, theta, group=None):
        r, pt = find_offset_for_radius(self.prober, r_e, theta, self.setup, group=group, r_min=_r_min_of(self.prober, group))
        if np.any(np.isnan(r)):
            k = int(np.nonzero(np.isnan(r))[0][0])
            raise RuntimeError(f"Transfer function integration failed (rₑ={r_e[k]}, θ={theta[k]}).")
        J = jacobian_ab_gr(self.prober, r * np.cos(theta) + self.setup.alpha0, r * np.sin(theta) + self.setup.beta0, self.setup,
                           group=group)
        return pt["g"], J, pt["t"]


def _r_min_of(prober, group):
    """`inner_radius(m)` of each pair's metric (one metric per cell of a table)."""
    if group is None or not hasattr(prober, "probers"):
        return api.inner_radius(prober.m)
    rm = np.array([api.inner_radius(p.m) for p in prober.probers])
    return rm[np.asarray(group)]


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
        r, pt = find_offset_for_radius(self.prober, r_e, theta, self.setup, height=h, group=group, r_min=_r_min_of(self.prober, group))
        if np.any(np.isnan(r)):
            k = int(np.nonzero(np.isnan(r))[0][0])
            raise RuntimeError(f"Transfer function integration failed (rₑ={r_e[k]}, θ={theta[k]}).")
        alpha = r * np.cos(theta) + self.setup.alpha0
        beta = r * np.sin(theta) + self.setup.beta0
        # the reference re-traces with the default chart and stops at 1.1 λ_max of the datum-plane point: a disc hit
        # later than that is no hit
        gt = self.prober.points(alpha, beta, thick=True, default_chart=True, **gq)
        status = np.where((gt.status == api.StatusCodes.IntersectedWithGeometry) & (gt.lambda_max > 1.1 * pt["lambda_max"]),
                          api.StatusCodes.NoStatus, gt.status)
        dist = np.linalg.norm(pt["x"] - gt.x, axis=0)
        close = dist <= 1e-3 * np.maximum(np.linalg.norm(pt["x"], axis=0), np.linalg.norm(gt.x, axis=0))  # isapprox(rtol = 1e-3)
        ok = (status == pt["status"]) & close
        J = np.full(r.size, np.nan)
        if ok.any():
            inner = self.prober.disc_inner_radius(None if group is None else group[ok])
            J[ok] = jacobian_ab_gr(self.prober, alpha[ok], beta[ok], self.setup, thick=True, group=None if group is None else group[ok],
                                   disc_inner_radius=inner)
        J[~np.isfinite(J)] = np.nan  # `is_visible = isfinite(J)`; invisible samples keep g and t, J = NaN (utils.jl:71-78)
        return pt["g"], J, pt["t"]


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
    setup_keys = {"theta_offset", "zero_atol", "N", "N_extrema", "h", "max_iter", "contrapoint_bias", "alpha0", "beta0"}
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
    ensemble = ensemble if ensemble is not None else api.default_ensemble()
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
