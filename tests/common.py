"""Shared problem builders for the tests: the BASELINE.json configs at arbitrary resolution."""
import math

import numpy as np

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
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
