#!/usr/bin/env python
"""Kernel-time table over the BASELINE.json configs on one GPU (fills BASELINE.md section 3)."""
import ctypes as C
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gradus_b200 as gb
from gradus_b200 import _cabi as cabi

ens = gb.EnsembleB200(devices=(0,))
cpf = gb.ConstPointFunctions
RED = [cpf.redshift() @ cpf.filter_intersected(), cpf.radius() @ cpf.filter_intersected()]


def fpa(metric_kind, disc):
    return 6 * (118 if metric_kind == 0 else 163) + 566 + (280 if disc else 0)


def report(name, n, st, metric_kind, disc, peak):
    att = st.steps_accepted + st.steps_rejected
    tf = att * fpa(metric_kind, disc) / (st.kernel_ms * 1e-3) / 1e12
    print(f"{name:62s} rays {n:9d} kernel {st.kernel_ms:8.2f} ms  {n / st.kernel_ms / 1e3:7.2f} Mrays/s  attempts/ray {att / n:6.1f} "
          f"rejected {st.steps_rejected / att:6.3%}  {tf:5.2f} TFLOP/s = {tf / peak:5.1%} of {peak:.1f}", flush=True)


def render(name, m, x, d, w, h, lam=2000.0, pf=RED, **kw):
    best = None
    for _ in range(3):
        args = (m, x, d, lam) if d is not None else (m, x, lam)
        gb.rendergeodesics(*args, pf=pf, image_width=w, image_height=h, ensemble=ens, **kw)
        st = ens.stats()
        if best is None or st.kernel_ms < best.kernel_ms:
            best = st
    report(name, w * h, best, m.kind, d is not None, PEAK)


pk = C.c_double(); cabi.check(cabi.load().gb200_fp64_peak(ens.ctx(0), C.byref(pk))); PEAK = pk.value
x60 = [0.0, 1000.0, math.radians(60), 0.0]
kerr = gb.KerrMetric(1.0, 0.998)
render("C1 Kerr a=0.998 ThinDisc(0,50) 128x128", kerr, x60, gb.ThinDisc(0.0, 50.0), 128, 128)
render("C2 Kerr a=0.998 ThinDisc(0,50) 2048x2048", kerr, x60, gb.ThinDisc(0.0, 50.0), 2048, 2048)
render("   Kerr a=0.998 no disc (shadow pf) 2048x2048", kerr, x60, None, 2048, 2048, pf=[cpf.shadow()])
render("   Kerr a=0.998 ShakuraSunyaev 2048x2048", kerr, x60, gb.ShakuraSunyaev(kerr), 2048, 2048)
jp = gb.JohannsenPsaltisMetric(1.0, 0.6, 2.0)
render("C5 JP a=0.6 eps3=2 ThinDisc(isco,50) 2048x2048", jp, x60, gb.ThinDisc(gb.isco(jp), 50.0), 2048, 2048)
jp2 = gb.JohannsenPsaltisMetric(1.0, 0.8831, 0.4)
render("C5 JP a=0.8831 eps3=0.4 ThinDisc(2,50) 2048x2048 (radius pf)", jp2, x60, gb.ThinDisc(2.0, 50.0), 2048, 2048, pf=[cpf.radius() @ cpf.filter_intersected()])
x40 = [0.0, 1000.0, math.radians(40), 0.0]
bins = np.linspace(0.1, 1.5, 180)
for nr in (1024, 4096):
    plane = gb.PolarPlane(gb.GeometricGrid(), Nr=nr, Ntheta=nr, r_min=1.0, r_max=250.0)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        _, flux = gb.lineprofile(bins, gb.PowerLawEmissivity(3.0), kerr, x40, gb.ThinDisc(0.0, 400.0), gb.BinningMethod(), plane=plane, lambda_max=2000.0, ensemble=ens)
        wall = time.perf_counter() - t0
        st = ens.stats()
        if best is None or st.kernel_ms < best[0].kernel_ms:
            best = (st, wall)
    report(f"C3 lineprofile Kerr theta=40 PolarPlane {nr}x{nr} (wall {best[1]*1e3:.0f} ms)", nr * nr, best[0], 0, True, PEAK)
xs = [0.0, 10.0, 0.01, 0.0]
delta = np.radians(np.linspace(0.01, 179.99, 1000))
vs = np.stack([np.zeros_like(delta), -np.cos(delta), np.sin(delta) / 10.0, np.zeros_like(delta)], axis=1)
t0 = time.perf_counter(); nl = 50
for k in range(nl):
    a = 0.998 * k / (nl - 1)
    gps = gb.tracegeodesics(gb.KerrMetric(1.0, a), xs, vs, gb.ThinDisc(0.0, 1000.0), 10000.0, callback=gb.domain_upper_hemisphere(), ensemble=ens)
wall = time.perf_counter() - t0
st = ens.stats()
print(f"C4 lamp-post-like 1000-ray ensembles x {nl} spins: {wall / nl * 1e3:.2f} ms per ensemble wall (last kernel {st.kernel_ms:.3f} ms) -> {1000 * nl / wall / 1e6:.3f} Mrays/s; hits {int((gps.status == 2).sum())}")

from gradus_b200 import tracegeodesics_batch
from gradus_b200.api import tracing_configuration
cfgs = [tracing_configuration(gb.KerrMetric(1.0, 0.998 * k / (nl - 1)), xs, vs, gb.ThinDisc(0.0, 1000.0), 10000.0, callback=gb.domain_upper_hemisphere(), ensemble=ens) for k in range(nl)]
tracegeodesics_batch(cfgs)
t0 = time.perf_counter(); out = tracegeodesics_batch(cfgs); wall = time.perf_counter() - t0
st = ens.stats()
print(f"C4 batched: {nl} ensembles x 1000 rays in one gb200_trace_batch call: {wall * 1e3:.2f} ms wall ({st.total_ms:.2f} ms device span) -> {1000 * nl / wall / 1e6:.3f} Mrays/s")
