"""Shared problem builders for the tests: the BASELINE.json configs at arbitrary resolution."""
import math

import numpy as np

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200 import api
from gradus_b200.api import RenderGrid, tracing_configuration


def render_config(m, x, d, lam, w, h, alims=(-60, 60), blims=(-40, 40), ensemble=None, **kw):
    cfg = tracing_configuration(m, x, RenderGrid(w, h, tuple(alims), tuple(blims)), *( [d] if d is not None else [] ), lam,
                                ensemble=ensemble, trajectories=w * h, **kw)
    return cfg


def c1(w=128, h=128, ensemble=None):
    """C1/C2: Kerr a=0.998, observer r=1000 theta=60deg, ThinDisc(0,50), lambda_max=2000."""
    m = gb.KerrMetric(1.0, 0.998)
    x = [0.0, 1000.0, math.radians(60.0), 0.0]
    return m, x, gb.ThinDisc(0.0, 50.0), render_config(m, x, gb.ThinDisc(0.0, 50.0), 2000.0, w, h, ensemble=ensemble)


def c3(nr=64, nth=64, ensemble=None):
    """C3: line-profile plane, Kerr a=0.998 theta=40deg, ThinDisc(0,400), PolarPlane geometric r in [1,250]."""
    m = gb.KerrMetric(1.0, 0.998)
    x = [0.0, 1000.0, math.radians(40.0), 0.0]
    d = gb.ThinDisc(0.0, 400.0)
    plane = gb.PolarPlane(gb.GeometricGrid(), Nr=nr, Ntheta=nth, r_min=1.0, r_max=250.0)
    cfg = tracing_configuration(m, x, plane, d, (0.0, 2000.0), callback=gb.domain_upper_hemisphere(), ensemble=ensemble)
    return m, x, d, plane, cfg


def c5(w=64, h=64, a=0.6, eps3=2.0, ensemble=None, inner=None):
    m = gb.JohannsenPsaltisMetric(1.0, a, eps3)
    x = [0.0, 1000.0, math.radians(60.0), 0.0]
    d = gb.ThinDisc(gb.isco(m) if inner is None else inner, 50.0)
    return m, x, d, render_config(m, x, d, 2000.0, w, h, ensemble=ensemble)


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


class OracleProber(gb.DeviceProber):
    """The transfer-function orchestration's tracer with the CPU oracle in place of the device (tests only)."""

    def evaluate(self, config):
        from oracle import oracle
        p, ic = config.to_c()
        kinds = [f.kind() for f in self.pfs]
        return oracle.render(p, ic, kinds, plunging=None)

    def evaluate_points(self, config):
        from oracle import oracle
        from gradus_b200 import api
        p, ic = config.to_c()
        return api.GeodesicPoints(oracle.trace(p, ic), config.lambda_domain[0])

    def evaluate_dual(self, config, arrays, norm_mode):
        from oracle import oracle
        p, _ = config.to_c()
        return oracle.trace_dual(p, arrays, norm_mode)


def oracle_plunging_table(kind, mp):
    """interpolate_plunging_velocities (src/orbits/orbit-solving.jl:137-167) restated with the oracle's pieces."""
    risco = _oracle().isco(kind, mp)
    g, dr, _ = _oracle().metric(kind, mp, risco, math.pi / 2)
    D = g[0] * g[3] - g[4] ** 2
    gitt, giphph, gitph = g[3] / D, g[0] / D, -g[4] / D
    v = _oracle().circular_fourvelocity(kind, mp, risco)
    E = _oracle().circular_energy(kind, mp, risco)
    ut = -E
    uph = (v[3] - gitph * ut) / giphph  # v^phi = g^tphi u_t + g^phiphi u_phi
    nom = gitt * E * E - 2 * gitph * E * uph + giphph * uph * uph + 1
    vr = -math.sqrt(abs(nom / (-g[1])))
    p = cabi.Problem()
    p.metric_kind = kind
    p.metric_params[:] = list(mp) + [0.0] * (8 - len(mp))
    p.mu, p.abstol, p.reltol, p.lambda_min, p.lambda_max, p.gtol = 1.0, 1e-9, 1e-9, 0.0, 50000.0, 1e-2
    p.chart_inner = (mp[0] + math.sqrt(mp[0] ** 2 - mp[1] ** 2)) * 1.000001
    p.chart_outer = 12000.0
    u0 = np.array([0.0, risco - 1e-8, math.pi / 2, 0.0, v[0], vr, 0.0, v[3]])
    g0, _, _ = _oracle().metric(kind, mp, u0[1], u0[2])
    disc = -g0[0] * g0[1] * vr**2 - g0[0] - (g0[0] * g0[3] - g0[4] ** 2) * v[3] ** 2  # constrain_time, mu = 1
    u0[4] = -(g0[4] * v[3] + math.sqrt(disc)) / g0[0]
    t, dt, ee, u = _oracle().trace_path(p, u0, cap=1 << 20)
    u = np.vstack([u0, u])
    order = np.argsort(u[:, 1], kind="stable")[1:]
    return u[order, 1], u[order, 4], u[order, 5], u[order, 7]


def _oracle():
    from oracle import oracle
    return oracle


def oracle_solver(configs):
    """A list of TracingConfigurations traced by the CPU oracle (tests only): stands in for the device solver."""
    from gradus_b200 import api
    out = []
    for config in configs:
        p, ic = config.to_c()
        out.append(api.GeodesicPoints(_oracle().trace(p, ic), config.lambda_domain[0]))
    return out


def oracle_target_tracer(config, target, d_tol):
    """`api.trace_target` with the CPU oracle in place of the device (tests only)."""
    oracle = _oracle()
    p, ic = config.to_c()
    closest, out = oracle.trace_target(p, ic, target, d_tol)
    return closest, api.GeodesicPoints(out, config.lambda_domain[0])

