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
    # False: every offset search starts from max(20, rₑ) like the reference's (`initial_r`, precision-solvers.jl:141).
    # True: a golden-section probe starts from the offset its predecessor converged to (the angle moved by < 0.2 rad), which
    # cuts its Newton iteration from ~20 trials to ~4 and with it the number of sequential launches of a table; the roots
    # satisfy the same |ρ − rₑ| ≤ zero_atol, the iteration path -- and hence the last digits of g -- differ.
    warm_start: bool = False
    # 0: a Newton iteration that cannot reach zero_atol runs all `max_iter` trials like the reference's (about 5 % of the
    # pairs sit at the noise floor of ρ(x), |y| ≈ 1e-6: the trace's own error at 1e-9 from r = 1e5, and are accepted by the
    # |y| ≤ 1e-4 rₑ rule in the end).  n > 0: such a pair stops after n trials without a new smallest |y| -- the same
    # acceptance rule, 40 fewer sequential launches for whoever waits for that pair.
    stall_exit: int = 0
    # origin of the polar coordinates on the image plane (`_rθ_to_αβ`, precision-solvers.jl:1-7)
    alpha0: float = 0.0
    beta0: float = 0.0


def theta_samples(setup: TransferFunctionSetup) -> np.ndarray:
    """The N fixed angles, clustered where g has its extrema (cunningham-transfer-functions.jl:366-370)."""
    K = setup.N // 5
    o = 2 * setup.theta_offset
    return np.concatenate([np.linspace(-o, o, K), np.linspace(-math.pi / 2, 3 * math.pi / 2, setup.N - 2 * K),
                           np.linspace(math.pi - o, math.pi + o, K)])


# phases of one (r_target, θ) pair inside `OffsetEngine`
_INIT, _NEWTON, _REDO, _BR_LO, _BR_MID, _FINAL, _DONE = range(7)


class OffsetEngine:
    """`_find_offset_for_radius` (precision-solvers.jl:133-241) as a batch of independent state machines, one per
    (r_target, θ) pair: the image-plane offset x with ρ(x cos θ, x sin θ) = r_target.  The iteration is the reference's,
    pair by pair -- Newton steps x − y / y′ with y′ = dρ/dx read off a dual number pushed through the trace, a contrapoint
    inside the hole that pulls overshoots back (`contrapoint_bias`), the "converge failed" exit, cycle detection on the
    relative decrease with a bracketing finish (`find_zero(f, (contra, x), atol)`), the late bracketing after `max_iter`
    -- but every pair advances on its own: one `round()` evaluates the next trial offset of EVERY unfinished pair in a
    single launch, whatever phase each is in, and pairs can be added between rounds.  A straggler (a pair that sits at
    the noise floor of ρ(x) for all `max_iter` iterations, or bisects for 60 halvings) therefore costs rounds only to
    whoever waits for that very pair."""

    def __init__(self, prober, setup):
        self.prober, self.setup = prober, setup
        self.n = 0
        cap = 1024
        f = lambda fill=0.0, dt=np.float64, shape=(): np.full((cap,) + shape, fill, dt)
        self.a = dict(rt=f(), ct=f(), st=f(), hgt=f(), grp=f(0, np.int64), rmin=f(), x=f(), y=f(), df=f(), contra=f(), dy=f(),
                      prev=f(0.0, np.float64, (6,)), it=f(0, np.int64), phase=f(_DONE, np.int64), trial=f(), lo=f(), hi=f(), flo=f(),
                      fhi=f(), best=f(), root=f(), halvings=f(0, np.int64), ymin=f(np.inf), stalled=f(0, np.int64),
                      g=f(np.nan), rho=f(np.nan), t=f(np.nan), status=f(-1, np.int32), lam=f(np.nan), px=f(np.nan, np.float64, (4,)),
                      alpha=f(np.nan), beta=f(np.nan))
        self.has_height = False
        self.has_group = False
        self.rounds = 0

    def _grow(self, need):
        cap = self.a["rt"].shape[0]
        if need <= cap:
            return
        new = max(need, 2 * cap)
        for k, v in self.a.items():
            w = np.empty((new,) + v.shape[1:], v.dtype)
            w[:cap] = v
            if k == "phase":
                w[cap:] = _DONE
            self.a[k] = w

    def add(self, r_target, theta, height=None, group=None, r_min=0.0, initial_r=None):
        """New pairs; returns their ids."""
        r_target = np.atleast_1d(np.asarray(r_target, np.float64))
        theta = np.atleast_1d(np.asarray(theta, np.float64))
        m = r_target.size
        ids = np.arange(self.n, self.n + m)
        self._grow(self.n + m)
        A = self.a
        A["rt"][ids], A["ct"][ids], A["st"][ids] = r_target, np.cos(theta), np.sin(theta)
        A["rmin"][ids] = r_min
        if height is not None:
            A["hgt"][ids] = height
            self.has_height = True
        if group is not None:
            A["grp"][ids] = group
            self.has_group = True
        A["x"][ids] = np.maximum(20.0, r_target) if initial_r is None else initial_r
        A["trial"][ids] = A["x"][ids]
        for k in ("contra", "dy", "y", "df"):
            A[k][ids] = 0.0
        A["prev"][ids] = 0.0
        A["it"][ids] = 0
        A["halvings"][ids] = 0
        A["ymin"][ids] = np.inf
        A["stalled"][ids] = 0
        A["phase"][ids] = _INIT
        self.n += m
        return ids

    def pending(self):
        return bool(np.any(self.a["phase"][: self.n] != _DONE))

    def _finish_loop(self, ids):
        """After the Newton loop of `ids` ended (converged or out of iterations): the late bracketing, or done."""
        A, s = self.a, self.setup
        late = (A["it"][ids] >= s.max_iter) & (A["y"][ids] > 10.0)
        L, D = ids[late], ids[~late]
        A["lo"][L], A["hi"][L], A["fhi"][L] = A["contra"][L], A["x"][L], A["y"][L]
        A["trial"][L] = A["lo"][L]
        A["phase"][L] = _BR_LO
        A["phase"][D] = _DONE
        return D

    def round(self):
        """One launch: the next trial offset of every unfinished pair.  Returns the ids that finished in this round."""
        A, s = self.a, self.setup
        atol, bias = s.zero_atol, s.contrapoint_bias
        act = np.nonzero(A["phase"][: self.n] != _DONE)[0]
        if act.size == 0:
            return act
        self.rounds += 1
        xr = A["trial"][act]
        al, be = xr * A["ct"][act] + s.alpha0, xr * A["st"][act] + s.beta0
        kw = {}
        if self.has_height:
            kw["height"] = A["hgt"][act]
        if self.has_group:
            kw["group"] = A["grp"][act]
        res = self.prober.dual(al, be, A["ct"][act][None, :], A["st"][act][None, :], **kw)
        A["g"][act], A["rho"][act], A["t"][act], A["status"][act] = res.g, res.rho, res.x[0], res.status
        A["lam"][act], A["px"][act], A["alpha"][act], A["beta"][act] = res.lambda_max, res.x.T, al, be
        fy = res.rho - A["rt"][act]  # y of the reference's `step`
        ph = A["phase"][act].copy()
        done = []

        # ---- first evaluation
        I = act[ph == _INIT]
        if I.size:
            A["y"][I], A["df"][I] = fy[ph == _INIT], res.drho[0][ph == _INIT]
            conv = np.abs(A["y"][I]) <= atol
            A["phase"][I[conv]] = _DONE
            done.append(I[conv])
            G = I[~conv]
            A["phase"][G] = _NEWTON
            with np.errstate(all="ignore"):
                A["trial"][G] = A["x"][G] - A["y"][G] / A["df"][G]

        # ---- a Newton trial (or its contrapoint replacement) came back
        for code in (_NEWTON, _REDO):
            sel = ph == code
            K = act[sel]
            if K.size == 0:
                continue
            next_x, next_y = A["trial"][K].copy(), fy[sel]
            A["df"][K] = res.drho[0][sel]
            proceed = np.ones(K.size, bool)
            if code == _NEWTON:
                pull = (next_x < 0) | ((next_y < 0) & (A["y"][K] > 0))
                A["contra"][K] = np.where(pull, np.maximum(A["contra"][K], next_x), A["contra"][K])
                # the overshoot ended in (or next to) the hole: step back towards the contrapoint instead
                redo = pull & ((next_x < 0) | (res.rho[sel] < A["rmin"][K] + 1))
                R = K[redo]
                A["trial"][R] = (A["contra"][R] * bias + A["x"][R]) / (1 + bias)
                A["phase"][R] = _REDO
                proceed = ~redo
            P, nx, ny = K[proceed], next_x[proceed], next_y[proceed]
            if P.size == 0:
                continue
            with np.errstate(all="ignore"):
                failed = (ny < 0) & (A["y"][P] < 0) & ((-A["y"][P] / A["df"][P]) < 0)  # "Converge failed": x, y keep their old values
                next_dy = (A["y"][P] - ny) / A["y"][P]
                cycle = ~failed & (A["y"][P] > 0) & np.any(np.abs(next_dy[:, None] - A["prev"][P]) <= atol * 100, axis=1)
            F = P[failed]
            done.append(self._finish_loop(F))
            Cy = P[cycle]  # stuck with Newton-Raphson: finish off by bracketing between the contrapoint and x
            A["lo"][Cy], A["hi"][Cy], A["fhi"][Cy] = A["contra"][Cy], A["x"][Cy], A["y"][Cy]
            A["trial"][Cy] = A["lo"][Cy]
            A["phase"][Cy] = _BR_LO
            go = ~failed & ~cycle
            G = P[go]
            A["x"][G], A["dy"][G], A["y"][G] = nx[go], next_dy[go], ny[go]
            A["prev"][G, A["it"][G] % 6] = A["dy"][G]
            A["it"][G] += 1
            stop = (np.abs(A["y"][G]) <= atol) | (A["it"][G] > s.max_iter)
            if s.stall_exit > 0:
                improved = np.abs(A["y"][G]) < A["ymin"][G]
                A["ymin"][G] = np.where(improved, np.abs(A["y"][G]), A["ymin"][G])
                A["stalled"][G] = np.where(improved, 0, A["stalled"][G] + 1)
                stop |= (A["stalled"][G] >= s.stall_exit) & (np.abs(A["y"][G]) <= 1e-4 * A["rt"][G])
            done.append(self._finish_loop(G[stop]))
            C = G[~stop]
            A["phase"][C] = _NEWTON
            with np.errstate(all="ignore"):
                A["trial"][C] = A["x"][C] - A["y"][C] / A["df"][C]

        # ---- bracketing: f(lo) came back
        sel = ph == _BR_LO
        B = act[sel]
        if B.size:
            A["flo"][B] = fy[sel]
            lo_better = np.abs(A["flo"][B]) <= np.abs(A["fhi"][B])
            A["root"][B] = np.where(lo_better, A["lo"][B], A["hi"][B])
            A["best"][B] = np.minimum(np.abs(A["flo"][B]), np.abs(A["fhi"][B]))
            ok = np.sign(A["flo"][B]) * np.sign(A["fhi"][B]) < 0
            good_enough = A["best"][B] <= atol
            bad = ~ok & ~good_enough  # no sign change: Roots.jl raises; the pair is reported as not found
            A["x"][B[bad]], A["y"][B[bad]] = np.nan, np.inf
            A["phase"][B[bad]] = _DONE
            done.append(B[bad])
            fin = ~bad & good_enough
            A["x"][B[fin]] = A["root"][B[fin]]
            A["trial"][B[fin]] = A["root"][B[fin]]
            A["phase"][B[fin]] = _FINAL
            M = B[~bad & ~good_enough]
            A["halvings"][M] = 0
            A["trial"][M] = 0.5 * (A["lo"][M] + A["hi"][M])
            A["phase"][M] = _BR_MID
        sel = ph == _BR_MID
        B = act[sel]
        if B.size:
            fm, mid = fy[sel], A["trial"][B].copy()
            better = np.abs(fm) < A["best"][B]
            A["root"][B[better]], A["best"][B[better]] = mid[better], np.abs(fm[better])
            same = np.sign(fm) == np.sign(A["flo"][B])
            A["lo"][B] = np.where(same, mid, A["lo"][B])
            A["flo"][B] = np.where(same, fm, A["flo"][B])
            A["hi"][B] = np.where(same, A["hi"][B], mid)
            A["halvings"][B] += 1
            more = (np.abs(fm) > atol) & (A["hi"][B] - A["lo"][B] > 4 * np.finfo(float).eps * np.abs(mid)) & (A["halvings"][B] < 80)
            A["trial"][B[more]] = 0.5 * (A["lo"][B[more]] + A["hi"][B[more]])
            E = B[~more]
            A["x"][E] = A["root"][E]
            A["trial"][E] = A["root"][E]
            A["phase"][E] = _FINAL
        # ---- `point, df, y = step(x)` at the bracketed root
        sel = ph == _FINAL
        B = act[sel]
        if B.size:
            A["y"][B], A["df"][B] = fy[sel], res.drho[0][sel]
            A["phase"][B] = _DONE
            done.append(B)
        return np.concatenate(done) if done else np.zeros(0, np.int64)

    def offset(self, ids):
        """The offsets of finished pairs: NaN where the reference returns NaN (negative offset, or |y| > 1e-4 r_target)."""
        A = self.a
        x, y = A["x"][ids], A["y"][ids]
        with np.errstate(invalid="ignore"):
            poor = ~np.isfinite(x) | (x < 0) | ~(np.abs(y) <= 1e-4 * A["rt"][ids])
        return np.where(poor, np.nan, x)

    def point(self, ids):
        """The last trace of each pair (`point` of the reference): g, rho, t = x[0], status, lambda_max, x, alpha, beta."""
        A = self.a
        return dict(g=A["g"][ids], rho=A["rho"][ids], t=A["t"][ids], status=A["status"][ids], lambda_max=A["lam"][ids],
                    x=A["px"][ids].T, alpha=A["alpha"][ids], beta=A["beta"][ids])


def find_offset_for_radius(prober, r_target, theta, setup: TransferFunctionSetup = TransferFunctionSetup(), initial_r=None,
                           height=None, group=None, r_min=None):
    """`find_offset_for_radius` (precision-solvers.jl:243-270) for a batch of (r_target[i], theta[i]) pairs: runs an
    `OffsetEngine` to completion.  Returns (x, point): x is NaN where no offset was found."""
    r_target = np.asarray(r_target, np.float64)
    if r_min is None:
        r_min = api.inner_radius(prober.m) if getattr(prober, "m", None) is not None else 0.0
    eng = OffsetEngine(prober, setup)
    ids = eng.add(r_target, theta, height=height, group=group, r_min=r_min, initial_r=initial_r)
    while eng.pending():
        eng.round()
    return eng.offset(ids), eng.point(ids)


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


def _r_min_of(prober, group):
    """`inner_radius(m)` of each pair's metric (one metric per cell of a table)."""
    if group is None or not hasattr(prober, "probers"):
        return api.inner_radius(prober.m)
    rm = np.array([api.inner_radius(p.m) for p in prober.probers])
    return rm[np.asarray(group)]


class _GoldenChain:
    """Optim.jl's `GoldenSection` univariate minimiser, one evaluation at a time: `theta` is the next abscissa to evaluate
    (None when finished), `feed(f)` hands over the objective there.  `_search_extremal!` runs it with
    `iterations = N_extrema − 1` on g (minimum near θ = 0) and on −g (maximum near θ = π),
    cunningham-transfer-functions.jl:391-426.  Same sequence of abscissae as `_golden_sections`."""

    def __init__(self, lower, upper, iterations, sign, rel_tol=math.sqrt(np.finfo(float).eps), abs_tol=np.finfo(float).eps):
        self.lower, self.upper, self.left, self.sign = float(lower), float(upper), int(iterations), sign
        self.rel_tol, self.abs_tol = rel_tol, abs_tol
        self.xm = self.lower + _GOLDEN * (self.upper - self.lower)
        self.fm = None
        self.theta = self.xm
        self.right = False

    def feed(self, f):
        if self.fm is None:
            self.fm = f
        else:
            xn, better = self.theta, f < self.fm
            if self.right:
                if better:
                    self.lower = self.xm
                else:
                    self.upper = xn
            else:
                if better:
                    self.upper = self.xm
                else:
                    self.lower = xn
            if better:
                self.xm, self.fm = xn, f
        self.theta = None
        if self.left <= 0:
            return
        self.left -= 1
        tolx = self.rel_tol * abs(self.xm) + self.abs_tol
        mid = 0.5 * (self.upper + self.lower)
        if abs(self.xm - mid) <= 2 * tolx - 0.5 * (self.upper - self.lower):
            return
        self.right = (self.upper - self.xm) > (self.xm - self.lower)
        self.theta = self.xm + _GOLDEN * (self.upper - self.xm) if self.right else self.xm - _GOLDEN * (self.xm - self.lower)


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
    setup_keys = {"theta_offset", "zero_atol", "N", "N_extrema", "h", "max_iter", "contrapoint_bias", "alpha0", "beta0", "warm_start", "stall_exit"}
    if setup is None:
        alias = {"θ_offset": "theta_offset", "α₀": "alpha0", "β₀": "beta0"}
        skw = {alias.get(k, k): kwargs.pop(k) for k in list(kwargs) if alias.get(k, k) in setup_keys}
        setup = TransferFunctionSetup(**skw)
    if prober is None:
        prober = DeviceProber(m, x, d, max_time=max_time, chart=chart, ensemble=ensemble, **kwargs)
    radii = np.atleast_1d(np.asarray(radii, np.float64))
    R = radii.size
    groups = None if groups is None else np.asarray(groups, np.int64)  # cell of each radius (CellProber)
    thick = getattr(prober, "thick", None) is not None
    th0 = theta_samples(setup)
    N, NE = th0.size, setup.N_extrema
    M = N + 2 * NE
    thetas = np.full((R, M), np.nan)
    gs = np.full((R, M), np.nan)
    Js = np.full((R, M), np.nan)
    ts = np.full((R, M), np.nan)

    # Every sample (rₑ, θ) is one pair of the offset engine; when it has converged its redshift and arrival time are
    # final and its Jacobian (for a thick disc: after the visibility re-trace) is queued.  The two golden-section chains
    # of a radius only need g of their own last probe to choose the next angle, so every radius walks through its 17
    # probes at its own pace while the 80 fixed angles of all radii converge alongside
    # (`_cunningham_transfer_function!` and `_search_extremal!` do the same evaluations one after the other).
    eng = OffsetEngine(prober, setup)
    info = {}  # pair id -> (radius index, sample slot, golden chain or None)
    rmin_k = np.broadcast_to(np.asarray(_r_min_of(prober, groups), np.float64), (R,))
    h_k = prober.cross_section(radii, groups) if thick else None

    def request(ks, slots, angles, chains, initial_r=None):
        ks = np.asarray(ks)
        ids = eng.add(radii[ks], angles, height=None if h_k is None else h_k[ks], group=None if groups is None else groups[ks],
                      r_min=rmin_k[ks], initial_r=initial_r)
        thetas[ks, slots] = angles
        for i, k, sl, ch in zip(ids, ks, slots, chains):
            info[int(i)] = (int(k), int(sl), ch)

    def nudged(theta):  # "avoid poles", cunningham-transfer-functions.jl:404-406
        return theta + 1e-4 if (abs(theta) < 1e-4 or abs(abs(theta) - math.pi) < 1e-4) else theta

    request(np.repeat(np.arange(R), N), np.tile(np.arange(N), R), np.tile(th0, R), [None] * (R * N))
    off = setup.theta_offset
    cmin = [_GoldenChain(-off, off, NE - 1, 1.0) for _ in range(R)]
    cmax = [_GoldenChain(math.pi - off, math.pi + off, NE - 1, -1.0) for _ in range(R)]
    fill = np.zeros((R, 2), int)
    for base, chains in ((0, cmin), (1, cmax)):
        request(np.arange(R), np.full(R, N + base * NE), [nudged(c.theta) for c in chains], chains)
        fill[:, base] = 1

    vis_q, jac_q = [], []  # (radius index, slot, alpha, beta[, point status, lambda_max, x]) awaiting their traces

    def flush_visibility():
        k = np.array([q[0] for q in vis_q]); sl = np.array([q[1] for q in vis_q])
        al = np.array([q[2] for q in vis_q]); be = np.array([q[3] for q in vis_q])
        st = np.array([q[4] for q in vis_q]); lam = np.array([q[5] for q in vis_q]); px = np.array([q[6] for q in vis_q]).T
        vis_q.clear()
        gq = {} if groups is None else {"group": groups[k]}
        # the reference re-traces with the default chart and stops at 1.1 λ_max of the datum-plane point: a disc hit
        # later than that is no hit (`_thick_workhorse`, cunningham-transfer-functions.jl:301-316)
        gt = prober.points(al, be, thick=True, default_chart=True, **gq)
        status = np.where((gt.status == api.StatusCodes.IntersectedWithGeometry) & (gt.lambda_max > 1.1 * lam), api.StatusCodes.NoStatus, gt.status)
        dist = np.linalg.norm(px - gt.x, axis=0)
        close = dist <= 1e-3 * np.maximum(np.linalg.norm(px, axis=0), np.linalg.norm(gt.x, axis=0))  # isapprox(rtol = 1e-3)
        for j in np.nonzero((status == st) & close)[0]:
            jac_q.append((int(k[j]), int(sl[j]), al[j], be[j]))

    def flush_jacobians():
        k = np.array([q[0] for q in jac_q]); sl = np.array([q[1] for q in jac_q])
        al = np.array([q[2] for q in jac_q]); be = np.array([q[3] for q in jac_q])
        jac_q.clear()
        grp = None if groups is None else groups[k]
        J = jacobian_ab_gr(prober, al, be, setup, thick=thick, group=grp, disc_inner_radius=prober.disc_inner_radius(grp) if thick else None)
        J[~np.isfinite(J)] = np.nan  # `is_visible = isfinite(J)`; invisible samples keep g and t, J = NaN (utils.jl:71-78)
        Js[k, sl] = J

    while eng.pending() or vis_q or jac_q:
        if eng.pending():
            fin = eng.round()
            if fin.size:
                xs, pt = eng.offset(fin), eng.point(fin)
                if np.any(np.isnan(xs)):
                    j = int(np.nonzero(np.isnan(xs))[0][0])
                    k, sl, _ = info[int(fin[j])]
                    raise RuntimeError(f"Transfer function integration failed (rₑ={radii[k]}, θ={thetas[k, sl]}).")
                again = []
                for j, i in enumerate(fin):
                    k, sl, chain = info.pop(int(i))
                    gs[k, sl], ts[k, sl] = pt["g"][j], pt["t"][j]
                    if thick:
                        vis_q.append((k, sl, pt["alpha"][j], pt["beta"][j], pt["status"][j], pt["lambda_max"][j], pt["x"][:, j]))
                    else:
                        jac_q.append((k, sl, pt["alpha"][j], pt["beta"][j]))
                    if chain is not None:
                        chain.feed(chain.sign * pt["g"][j])
                        if chain.theta is not None:
                            base = 0 if chain.sign > 0 else 1
                            again.append((k, N + base * NE + fill[k, base], nudged(chain.theta), chain, xs[j]))
                            fill[k, base] += 1
                if again:
                    request([a_[0] for a_ in again], [a_[1] for a_ in again], [a_[2] for a_ in again], [a_[3] for a_ in again],
                            initial_r=np.array([a_[4] for a_ in again]) if setup.warm_start else None)
        idle = not eng.pending()
        if vis_q and (idle or len(vis_q) >= 4096):
            flush_visibility()
        if jac_q and (idle or len(jac_q) >= 4096):
            flush_jacobians()
    best = np.array([c.fm for c in cmin] + [c.fm for c in cmax], np.float64)  # minima of g and of −g
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
