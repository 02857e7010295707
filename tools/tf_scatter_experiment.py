#!/usr/bin/env python
"""How reproducible are the reference's transfer-function literals?  (tests/test_transfer_functions.py relies on this.)

The reference pins `measure_ctf = sum(f .* g✶) / length(f)` of eleven thin-disc transfer functions to atol 1e-3 (rtol 1e-2 for
the large radii; test/smoke-tests/cunningham-transfer-functions.jl:25-39) and `sum(filter(!isnan, f))` of two thick-disc ones
to atol 1e-4 / 1e-2 (test/transfer-functions/test-thick-disc.jl:10-19).  This script computes every one of them with the
reference's own algorithm (dual numbers through the integrator, its Newton iteration, golden sections, tolerance 1e-9;
CPU oracle as the tracer) under the choices its un-vendored, un-versioned ODE packages leave open:
    norm   whether DiffEqBase's error norm sees the partials of a dual-valued state
    pow    the step controller's power: exact or FastPower's Float32
    tol    1e-9 or 0.999e-9
and reports, per literal, the spread over the eight variants and the spread of the "resolved" part of the statistic
(samples with g✶(1 − g✶) > 1e-5 only).  Usage: python tools/tf_scatter_experiment.py > profiles/r02_tf_scatter.log"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gradus_b200 as gb  # noqa: E402
from gradus_b200 import transfer_functions as tf  # noqa: E402
from common import OracleProber  # noqa: E402
from test_transfer_functions import REFERENCE_CTF, THICK_LITERALS, VARIANTS, reference_tolerance, resolved_statistic  # noqa: E402


def main():
    print("variants (norm sees partials = 0 / values only = 1, pow exact = 0 / Float32 = 1, tolerance):", VARIANTS)
    print("\nthin disc: measure_ctf, observer r = 1e5, chart closest_approach = 1.005, N = 80 (+ 2 x 17 golden-section probes)")
    print(f"{'incl':>4} {'r_e':>7} {'literal':>10} {'ref tol':>8} | {'min':>10} {'max':>10} {'spread':>9} | literal - [min, max] | resolved-only mean, spread")
    for a, angle, re, lit, kind in REFERENCE_CTF:
        vals, res = [], []
        for nm, pw, tol in VARIANTS:
            m = gb.KerrMetric(1.0, a)
            x = [0.0, 100_000.0, math.radians(angle), 0.0]
            d = gb.ThinDisc(0.0, float("inf"))
            pr = OracleProber(m, x, d, chart=gb.chart_for_metric(m, 2 * x[1], closest_approach=1.005), abstol=tol, reltol=tol, pow_mode=pw)
            pr.norm_mode = nm
            ctf = tf.cunningham_transfer_function(m, x, d, re, prober=pr, N=80)
            vals.append(tf.measure_ctf(ctf))
            res.append(resolved_statistic(ctf))
        vals, res = np.array(vals), np.array(res)
        rt = reference_tolerance(lit, kind)
        inside = "inside" if vals.min() <= lit <= vals.max() else f"{min(abs(lit - vals.min()), abs(lit - vals.max())):.1e} outside"
        verdict = "reproducible at the reference's tolerance" if np.ptp(vals) <= rt else "NOT reproducible at the reference's tolerance"
        print(f"{angle:4d} {re:7.1f} {lit:10.6f} {rt:8.1e} | {vals.min():10.6f} {vals.max():10.6f} {np.ptp(vals):9.2e} | {inside:>14} | "
              f"{res.mean():.6f} {np.ptp(res):.1e} | {verdict}", flush=True)
    print("\nthick disc (ShakuraSunyaev): sum of the finite f, observer r = 1e4, beta0 = 2")
    for a, angle, kw, re, lit, rt in THICK_LITERALS:
        vals, res = [], []
        for nm, pw, tol in VARIANTS:
            m = gb.KerrMetric(1.0, a)
            x = [0.0, 10_000.0, math.radians(angle), 0.0]
            d = gb.ShakuraSunyaev(m, **kw)
            pr = OracleProber(m, x, d, chart=gb.chart_for_metric(m, 2 * x[1]), abstol=tol, reltol=tol, pow_mode=pw)
            pr.norm_mode = nm
            ctf = tf.cunningham_transfer_function(m, x, d, re, prober=pr, beta0=2.0)
            vals.append(float(np.nansum(ctf.f)))
            ok = np.isfinite(ctf.f) & (ctf.g_star * (1 - ctf.g_star) > 1e-5)
            res.append(float(np.sum(ctf.f[ok])))
        vals, res = np.array(vals), np.array(res)
        print(f"a = {a}, {angle} deg, r_e = {re:.4f}: literal {lit:.5f} (ref atol {rt:g}) | variants {vals.min():.5f} .. {vals.max():.5f} (spread {np.ptp(vals):.2e}) | "
              f"resolved-only sum {res.mean():.5f}, spread {np.ptp(res):.1e}", flush=True)


if __name__ == "__main__":
    main()
