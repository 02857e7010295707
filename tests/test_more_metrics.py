"""Further static, axis-symmetric metrics: Johannsen, Bumblebee, Kerr-Newman (neutral particles),
Morris-Thorne wormhole.

Pins: the reference's `rendergeodesics` smoke matrix (test/smoke-tests/rendergeodesics.jl:32-82 — shadow, thin-disc and
Shakura-Sunyaev fingerprints per metric, default parameters) for the oracle; GPU vs oracle at non-trivial parameters
with the same protocol as the main parity tests (identical termination class outside the oracle-defined grazing band,
disc-hit endpoints to 1e-6, redshift to 1e-6)."""
import math

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200 import api, hostmath
from oracle import oracle

from common import render_config

ORACLE_GEOMETRY_SS_LEGACY_GTOL = 101

# (metric with the reference's defaults, shadow literal, thin-disc literal, Shakura-Sunyaev literal)
SMOKE = [
    (gb.JohannsenMetric(), 9009.448935932085, 38412.08386562321, 34455.344169980635),
    (gb.BumblebeeMetric(), 9009.452384885506, 38412.0832157869, 34455.3441698318),
    (gb.KerrNewmanMetric(), 9009.451384824908, 38412.08517225652, 34455.34416971527),
]
NONTRIVIAL = [
    gb.JohannsenMetric(M=1.0, a=0.7, alpha13=0.35, alpha22=-0.2, alpha52=0.5, eps3=0.8),
    # a = 0.1: the chart's inner boundary 1.01 (M + sqrt(M^2 - a^2)) = 2.015 stays outside this metric's coordinate
    # singularity at r = 2M (for larger a the reference's inner_radius, bumblebee-ad.jl:51, lets rays run into it)
    gb.BumblebeeMetric(M=1.0, a=0.1, l=0.4),
    gb.KerrNewmanMetric(M=1.0, a=0.6, Q=0.5),
]
IDS = ["johannsen", "bumblebee", "kerr_newman"]
# Einstein-Maxwell-Dilaton-Axion (src/metrics/dilaton-axion-ad.jl): no entry in the smoke matrix; pinned on its ISCO literals
DILATON = [gb.DilatonAxion(M=1.0, a=0.5, beta=0.0, b=0.3), gb.DilatonAxion(M=1.0, a=0.6, beta=0.2, b=0.5)]


def _smoke_fixture(m, d):
    return render_config(m, [0.0, 100.0, math.radians(85), 0.0], d, 200.0, 20, 20, (-9.5, 9.5), (-9.5, 9.5))


@pytest.mark.parametrize("m, shadow, thin, ss", SMOKE, ids=IDS)
def test_oracle_reproduces_the_smoke_matrix(m, shadow, thin, ss):
    p, ic = _smoke_fixture(m, None).to_c()
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(shadow, rel=1e-6)
    p, ic = _smoke_fixture(m, gb.ThinDisc(0.0, 40.0)).to_c()
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(thin, rel=1e-6)
    p, ic = _smoke_fixture(m, gb.ShakuraSunyaev(m)).to_c()
    p.geometry_kind = ORACLE_GEOMETRY_SS_LEGACY_GTOL  # the form the literals were recorded with, see test_oracle_kat.py
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(ss, rel=1e-9)


def test_dilaton_axion_isco_literals():
    """test/smoke-tests/special-radii.jl:20-22,34-36: 29.701502242023523 and 6.0 (atol 1e-5 there)."""
    m = gb.DilatonAxion(M=1.0, a=0.6, beta=-0.5, b=0.1)
    assert api.isco(m) == pytest.approx(29.701502242023523, abs=1e-9)
    assert oracle.isco(m.kind, list(m.params())) == pytest.approx(29.701502242023523, abs=1e-9)
    assert api.isco(gb.DilatonAxion(M=1.0, a=0.0, beta=0.0, b=0.0)) == pytest.approx(6.0, abs=1e-9)
    assert api.inner_radius(gb.DilatonAxion(1.0, 0.5, 0.0, 1.0)) == pytest.approx(2.0 + math.sqrt(4.0 - 0.25))
    with pytest.raises(ValueError):
        gb.DilatonAxion(1.0, 0.0, 0.3, 1.0).params()  # β / a


@pytest.mark.parametrize("m", NONTRIVIAL + DILATON, ids=IDS + ["dilaton_axion", "dilaton_axion_beta"])
def test_host_and_oracle_metric_algebra_agree(m):
    mp = list(m.params())
    for r, th in [(3.0, 0.4), (9.0, 1.3), (250.0, 2.2)]:
        g, dr, _ = oracle.metric(m.kind, mp, r, th)
        assert np.allclose(hostmath.metric_components(m, r, th), g, rtol=1e-13, atol=0)
        assert np.allclose(hostmath.metric_dr(m, r, th), dr, rtol=1e-11, atol=1e-300)
    r_isco = api.isco(m)  # library host code (dual-number root find on the generated closed form)
    assert r_isco == pytest.approx(oracle.isco(m.kind, mp), rel=1e-10)
    assert 1.0 < r_isco < 9.0


def test_kerr_limits_of_the_deformed_metrics():
    """Zero deformation = Kerr (Johannsen, Kerr-Newman) / Schwarzschild (Bumblebee at a = 0, l = 0)."""
    k = gb.KerrMetric(1.0, 0.6)
    for m in (gb.JohannsenMetric(1.0, 0.6), gb.KerrNewmanMetric(1.0, 0.6, 0.0)):
        for r, th in [(2.5, 0.7), (30.0, 1.5)]:
            assert np.allclose(hostmath.metric_components(m, r, th), hostmath.metric_components(k, r, th), rtol=1e-13)
        assert api.isco(m) == pytest.approx(api.isco(k), rel=1e-9)
    assert api.isco(gb.BumblebeeMetric()) == pytest.approx(6.0, rel=1e-9)
    assert api.inner_radius(gb.KerrNewmanMetric(1.0, 0.6, 0.5)) == pytest.approx(1.0 + math.sqrt(1 - 0.36 - 0.25))


def test_constructor_checks_follow_the_reference():
    with pytest.raises(ValueError):
        gb.BumblebeeMetric(1.0, 0.0, -1.0)
    with pytest.raises(ValueError):
        gb.BumblebeeMetric(1.0, 0.5, 0.0)
    with pytest.raises(ValueError):
        gb.KerrNewmanMetric(1.0, 0.9, 0.9)


# --------------------------------------------------------------------------- Morris-Thorne wormhole
# test/smoke-tests/rendergeodesics.jl:36,46,60,93: the fourth column of the smoke matrix.  The only rays with a finite
# fingerprint in the shadow image are the four that pass the throat (l <= 0, the chart's inner boundary at
# inner_radius = 0): a discrete callback without a root find, so their lambda is the end of whichever step crossed l = 0.
# The oracle lands within 2e-6 (shadow), 7e-7 (thin disc) of the literals (the reference's own tolerance: rtol 1e-1).
MT_SHADOW, MT_THIN, MT_TORUS = 402.17907632733284, 9375.430228131403, 5104.032822512765


def test_oracle_reproduces_the_morris_thorne_column():
    m = gb.MorrisThorneWormhole()
    assert api.inner_radius(m) == 0.0
    p, ic = _smoke_fixture(m, None).to_c()
    img, ep = oracle.render(p, ic, [cabi.PF_SHADOW], endpoints=True)
    assert np.nansum(img) == pytest.approx(MT_SHADOW, rel=1e-5)
    assert np.count_nonzero(ep.status == cabi.STATUS_WITHIN_INNER_BOUNDARY) == 4 and np.all(ep.x[1, ep.status == cabi.STATUS_WITHIN_INNER_BOUNDARY] <= 0)
    p, ic = _smoke_fixture(m, gb.ThinDisc(0.0, 40.0)).to_c()
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(MT_THIN, rel=1e-5)
    p.geometry_kind = 100  # the torus closure (below), in the form the literal was recorded with
    p.geometry_params[3] = 1.0
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(MT_TORUS, rel=1e-3)
    with pytest.raises(gb.GradusB200Error):
        api.isco(m)  # no ISCO, as in the reference (the Shakura-Sunyaev entry of this column is skipped there, :63-64)


def test_morris_thorne_host_algebra():
    m = gb.MorrisThorneWormhole(b=1.5)
    for l_, th in [(-3.0, 0.4), (0.0, 1.3), (250.0, 2.2)]:
        g, dr, dth = oracle.metric(m.kind, list(m.params()), l_, th)
        assert np.allclose(hostmath.metric_components(m, l_, th), g, rtol=1e-15, atol=0)
        assert np.allclose(g, [-1.0, 1.0, 2.25 + l_ * l_, (2.25 + l_ * l_) * math.sin(th), 0.0])
        assert np.allclose(dr, [0, 0, 2 * l_, 2 * l_ * math.sin(th), 0]) and np.allclose(dth, [0, 0, 0, (2.25 + l_ * l_) * math.cos(th), 0])


@pytest.mark.gpu
def test_device_morris_thorne(ensemble):
    m = gb.MorrisThorneWormhole()
    x = [0.0, 100.0, math.radians(85), 0.0]
    kw = dict(image_width=20, image_height=20, alpha_lims=(-9.5, 9.5), beta_lims=(-9.5, 9.5), ensemble=ensemble)
    for d, literal in [(None, MT_SHADOW), (gb.ThinDisc(0.0, 40.0), MT_THIN)]:
        args = (m, x, 200.0) if d is None else (m, x, d, 200.0)
        _, _, img = gb.rendergeodesics(*args, **kw)
        p, ic = _smoke_fixture(m, d).to_c()
        want = oracle.render(p, ic, [cabi.PF_SHADOW])[0]
        # measured 3.4e-5 (shadow): in this nearly flat space the error estimate is rounding noise, so the step
        # sequences of two right-hand-side formulations drift apart at that level, and lambda of the throat crossing
        # (a step end, no root find) follows the step sequence
        assert np.nansum(img) == pytest.approx(literal, rel=1e-4)
        assert np.array_equal(np.isnan(img.T.reshape(-1)), np.isnan(want))
    # a wider view under the main protocol: same class everywhere (no grazing geometry here), disc hits to 1e-6
    d = gb.ThinDisc(2.0, 30.0)
    cfg = render_config(gb.MorrisThorneWormhole(b=2.0), x, d, 400.0, 96, 96, (-30, 30), (-12, 12), ensemble=ensemble)
    p, ic = cfg.to_c()
    ref = oracle.trace(p, ic)
    band = oracle.band_ratio(p, ic)
    graze = (band > 0) & (band < 1.3)
    got = api.solve_tracing_problem(cfg)
    ok = ~graze
    assert graze.mean() < 0.01 and np.array_equal(np.asarray(got.status)[ok], ref.status[ok])
    hit = ok & (ref.status == cabi.STATUS_INTERSECTED)
    assert hit.sum() > 500
    assert np.max(np.abs(got.x[1:3, hit] - ref.x[1:3, hit])) < 1e-6
    through = ok & (ref.status == cabi.STATUS_WITHIN_INNER_BOUNDARY)
    assert through.sum() > 20 and np.all(got.x[1, through] <= 0)


# --------------------------------------------------------------------------- device
@pytest.mark.gpu
@pytest.mark.parametrize("m, shadow, thin, ss", SMOKE, ids=IDS)
def test_device_reproduces_the_smoke_matrix(m, shadow, thin, ss):
    x = [0.0, 100.0, math.radians(85), 0.0]
    kw = dict(image_width=20, image_height=20, alpha_lims=(-9.5, 9.5), beta_lims=(-9.5, 9.5))
    _, _, img = gb.rendergeodesics(m, x, 200.0, **kw)
    assert np.nansum(img) == pytest.approx(shadow, rel=1e-6)
    _, _, img = gb.rendergeodesics(m, x, gb.ThinDisc(0.0, 40.0), 200.0, **kw)
    assert np.nansum(img) == pytest.approx(thin, rel=1e-6)
    _, _, img = gb.rendergeodesics(m, x, gb.ShakuraSunyaev(m), 200.0, **kw)
    assert np.nansum(img) == pytest.approx(ss, rel=0.1)  # current thick-disc source; the literal's legacy form is oracle-only


@pytest.mark.gpu
@pytest.mark.parametrize("m", NONTRIVIAL + DILATON, ids=IDS + ["dilaton_axion", "dilaton_axion_beta"])
def test_device_matches_oracle_at_nontrivial_parameters(m):
    x = [0.0, 1000.0, math.radians(70), 0.0]
    d = gb.ThinDisc(api.isco(m), 40.0)
    cfg = render_config(m, x, d, 2000.0, 96, 96, (-45, 45), (-30, 30))
    p, ic = cfg.to_c()
    want, ep = oracle.render(p, ic, [cabi.PF_REDSHIFT, cabi.PF_DISC_RADIUS], endpoints=True)
    band = oracle.band_ratio(p, ic)
    ep_l = oracle.trace(p, ic, precision=1)
    graze = ((band > 0) & (band < 1.3)) | (ep.status != ep_l.status)
    pfs = [gb.ConstPointFunctions.redshift(m, x) @ gb.ConstPointFunctions.filter_intersected(),
           gb.ConstPointFunctions.radius() @ gb.ConstPointFunctions.filter_intersected()]
    _, _, imgs = gb.rendergeodesics(m, x, d, 2000.0, pf=pfs, image_width=96, image_height=96, alpha_lims=(-45, 45), beta_lims=(-30, 30))
    got = np.stack([im.T.reshape(-1) for im in imgs])
    gps = api.solve_tracing_problem(cfg)
    ok = ~graze
    assert graze.mean() < 0.02
    assert np.array_equal(np.asarray(gps.status)[ok], ep.status[ok])
    hit = ok & (ep.status == cabi.STATUS_INTERSECTED)
    assert hit.sum() > 500
    assert np.max(np.abs(got[0][hit] - want[0][hit])) < 1e-6
    assert np.max(np.abs(got[1][hit] / want[1][hit] - 1)) < 1e-6
    assert np.array_equal(np.isnan(got[0][ok]), np.isnan(want[0][ok]))


# --------------------------------------------------------------------------- charged test particles in Kerr-Newman
# test/unit/metrics.kerr-newman.jl:6-30: shadow fingerprints of a 40 x 40 render for q = 0, +1, -1 (rtol 1e-3 there)
CHARGED = [(0.0, 428809.9681726607), (1.0, 253280.6794972752), (-1.0, 619335.5670363897)]


def _charged_config(q, w=40, h=40):
    m = gb.KerrNewmanMetric(M=1.0, a=0.6, Q=0.6)
    x = [0.0, 1000.0, math.pi / 2, 0.0]
    return m, x, api.tracing_configuration(m, x, api.RenderGrid(w, h, (-8, 8), (-8, 8)), 2000.0, q=q, trajectories=w * h)


@pytest.mark.parametrize("q, literal", CHARGED)
def test_oracle_reproduces_the_charged_particle_literals(q, literal):
    """Lorentz force q F v (faraday_tensor, src/tracing/utility.jl:89-99; kerr-newman-ad.jl:66-102), dual-number dA."""
    p, ic = _charged_config(q)[2].to_c()
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(literal, rel=1e-6)


def test_charge_needs_an_electromagnetic_potential():
    m = gb.KerrMetric(1.0, 0.5)
    with pytest.raises(ValueError):
        api.tracing_configuration(m, [0.0, 1000.0, 1.0, 0.0], api.RenderGrid(4, 4, (-8, 8), (-8, 8)), 2000.0, q=1.0).to_c()


@pytest.mark.gpu
@pytest.mark.parametrize("q, literal", CHARGED)
def test_device_reproduces_the_charged_particle_literals(q, literal):
    m, x, cfg = _charged_config(q)
    _, _, img = gb.rendergeodesics(m, x, 2000.0, image_width=40, image_height=40, alpha_lims=(-8, 8), beta_lims=(-8, 8), q=q)
    assert np.nansum(img) == pytest.approx(literal, rel=1e-6)
    # ray by ray against the oracle (closed-form Faraday tensor vs dual numbers): same classes, same affine times
    m, x, cfg = _charged_config(q, 64, 64)
    p, ic = cfg.to_c()
    ref = oracle.trace(p, ic)
    ref_l = oracle.trace(p, ic, precision=1)
    got = api.solve_tracing_problem(cfg)
    ok = ref.status == ref_l.status
    assert ok.mean() > 0.99 and np.array_equal(got.status[ok], ref.status[ok])
    esc = ok & (ref.status == cabi.STATUS_NO_STATUS)  # reached lambda_max: a well-defined end state
    assert esc.sum() > 1000
    # 1e-6, widened ray by ray by what the oracle itself moves between double and long double (charged massive particles
    # that orbit the hole before leaving amplify rounding by > 1e3: measured 1.9e-6 on the worst such ray of q = -1)
    scale = np.maximum(np.abs(ref.x[:, esc]), 1.0)
    own = np.abs(ref.x[:, esc] - ref_l.x[:, esc]) / scale
    dev = np.abs(got.x[:, esc] - ref.x[:, esc]) / scale
    assert np.all(dev < 1e-6 + 20.0 * own.max(axis=0)) and np.quantile(dev.max(axis=0), 0.99) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("m, shadow, thin, ss", SMOKE + [(gb.KerrMetric(1.0, 0.0), 9009.452876609641, None, None)], ids=IDS + ["kerr"])
def test_prerendergeodesics_cache_on_the_device(m, shadow, thin, ss):
    """test/smoke-tests/prerendergeodesics.jl: endpoints cached once, point functions applied afterwards."""
    x = [0.0, 100.0, math.radians(85), 0.0]
    _, _, cache = gb.prerendergeodesics(m, x, 200.0, image_width=20, image_height=20, alpha_lims=(-9.5, 9.5), beta_lims=(-9.5, 9.5))
    pf = gb.HostPointFunction(lambda m_, gp, lam: gp.lambda_max) @ gb.HostPointFunction(lambda m_, gp, lam: gp.lambda_max < lam, np.nan)
    img = gb.apply(pf, cache)
    assert np.nansum(img) == pytest.approx(shadow, rel=1e-6)
    _, _, fused = gb.rendergeodesics(m, x, 200.0, image_width=20, image_height=20, alpha_lims=(-9.5, 9.5), beta_lims=(-9.5, 9.5))
    assert np.array_equal(img, fused, equal_nan=True)  # the fused device point function writes the same numbers


def test_charged_circular_orbits_reproduce_the_reference_literals():
    """test/unit/metrics.kerr-newman.jl:38-41: CircularOrbits.fourvelocity(m, 20.0; q = +-1), rtol 1e-5 there."""
    m = gb.KerrNewmanMetric(M=1.0, a=0.6, Q=0.6)
    assert np.allclose(hostmath.circular_fourvelocity(m, 20.0, q=1.0), [1.065341126764724, 0.0, 0.0, 0.007652504287280518], rtol=1e-8, atol=0)
    assert np.allclose(hostmath.circular_fourvelocity(m, 20.0, q=-1.0), [1.099562453625687, 0.0, 0.0, 0.015091823134051219], rtol=1e-8, atol=0)
    # q = 0 falls back to the analytic angular velocity
    v0 = hostmath.circular_fourvelocity(m, 20.0)
    g = hostmath.metric_components(m, 20.0, math.pi / 2)
    assert hostmath.dot(g, v0, v0) == pytest.approx(-1.0, abs=1e-12)


# --------------------------------------------------------------------------- ThickDisc(f): closure cross-sections as tables
# test/smoke-tests/rendergeodesics.jl:8-15,85-96: a torus of unit radius centred on rho = 10; literals last computed 25/08/2023,
# i.e. (like the Shakura-Sunyaev ones) when the thick-disc distance still subtracted gtol |r|.
def _torus(rho):
    if rho < 9.0 or rho > 11.0:
        return -1.0
    return math.sqrt(1 - (rho - 10.0) ** 2)


TORUS = [(gb.KerrMetric(1.0, 0.0), 16918.69258396256), (gb.JohannsenMetric(), 16918.689593279843), (gb.BumblebeeMetric(), 16918.692092917947),
         (gb.KerrNewmanMetric(), 16918.691837255217)]
ORACLE_GEOMETRY_TEST_THICK_DISC = 100  # oracle-only: the closure itself


@pytest.mark.parametrize("m, literal", TORUS, ids=["kerr", "johannsen", "bumblebee", "kerr_newman"])
def test_oracle_torus_literal_and_its_tabulated_form(m, literal):
    p, ic = _smoke_fixture(m, gb.ThinDisc(0.0, 40.0)).to_c()
    p.geometry_kind = ORACLE_GEOMETRY_TEST_THICK_DISC
    exact = np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0])
    assert exact == pytest.approx(literal, rel=0.1)  # the reference's own tolerance, current source (measured -2.3 %)
    p.geometry_params[3] = 1.0  # the form the literal was recorded with: distance - gtol |r|
    assert np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0]) == pytest.approx(literal, rel=1e-4)  # measured 3.2e-5
    # what crosses the ABI is a table of the closure on 4097 Chebyshev nodes: same image to 1e-8
    d = gb.ThickDisc(_torus, (9.0, 11.0))
    p2, ic2 = _smoke_fixture(m, d).to_c()
    assert p2.geometry_kind == cabi.GEOMETRY_THICK_TABLE
    oracle.set_cross_section(d.rho, d.height)
    assert np.nansum(oracle.render(p2, ic2, [cabi.PF_SHADOW])[0]) == pytest.approx(exact, rel=1e-8)


@pytest.mark.gpu
def test_device_thick_disc_table(ensemble):
    """GB200_GEOMETRY_THICK_TABLE on the device: the smoke-matrix torus against the oracle (closure and table), and a
    128 x 128 render under the main parity protocol (same termination class outside the grazing band, hits to 1e-6)."""
    d = gb.ThickDisc(_torus, (9.0, 11.0))
    oracle.set_cross_section(d.rho, d.height)
    x = [0.0, 100.0, math.radians(85), 0.0]
    for m, literal in TORUS + [(gb.MorrisThorneWormhole(), MT_TORUS)]:
        _, _, img = gb.rendergeodesics(m, x, d, 200.0, image_width=20, image_height=20, alpha_lims=(-9.5, 9.5), beta_lims=(-9.5, 9.5), ensemble=ensemble)
        p, ic = _smoke_fixture(m, gb.ThinDisc(0.0, 40.0)).to_c()
        p.geometry_kind = ORACLE_GEOMETRY_TEST_THICK_DISC
        exact = np.nansum(oracle.render(p, ic, [cabi.PF_SHADOW])[0])
        # Morris-Thorne: 1.1e-6 measured (step sequences in nearly flat space follow rounding noise, see above)
        rel = 1e-5 if isinstance(m, gb.MorrisThorneWormhole) else 1e-6
        assert np.nansum(img) == pytest.approx(exact, rel=rel) and np.nansum(img) == pytest.approx(literal, rel=0.1)
    m = gb.KerrMetric(1.0, 0.9)
    cfg = render_config(m, x, d, 200.0, 128, 128, (-14, 14), (-6, 6), ensemble=ensemble)
    p, ic = cfg.to_c()
    ref = oracle.trace(p, ic)
    got = api.solve_tracing_problem(cfg)
    agree = got.status == ref.status
    assert agree.mean() > 0.995
    hit = agree & (ref.status == cabi.STATUS_INTERSECTED)
    assert hit.sum() > 1000
    rel = np.abs(got.x[1:3, hit] - ref.x[1:3, hit]) / np.maximum(np.abs(ref.x[1:3, hit]), 1e-3)
    assert np.quantile(rel.max(axis=0), 0.995) < 1e-6
    # the tabulated height is what the hits sit on
    rho = got.x[1, hit] * np.abs(np.sin(got.x[2, hit]))
    z = got.x[1, hit] * np.abs(np.cos(got.x[2, hit]))
    assert np.quantile(np.abs(z - d.cross_section(rho)), 0.99) < 1e-9


# --------------------------------------------------------------------------- PolishDoughnut (src/geometry/discs/polish-doughnut.jl)
def test_polish_doughnut_cross_section_fingerprint():
    """test/discs/test-polish-doughnut.jl: the cross-section map of the default torus around Kerr a = 0.2 (atol 1e-5 there).
    The host set-up restates the reference's (innermost radius by Newton on dE/dr, isobar by Tsit5 with dtmax = 5e-2 and
    OrdinaryDiffEq's defaults, linear interpolation over the saved steps): the literal is reproduced to 1e-11."""
    d = gb.PolishDoughnut(gb.KerrMetric(1.0, 0.2), rk=12.0, n=0.21)
    h = d.cross_section(np.linspace(10.0, 15.0, 200))
    assert h.sum() == pytest.approx(219.97440610254944, abs=1e-9)
    assert d.inner_radius == pytest.approx(10.0875, abs=1e-4) and 14.7 < d.outer_radius < 14.71
    assert np.all(np.diff(d.rho) > 0) and d.height[0] < 1e-12 and 1.5 < d.height.max() < 1.6
    with pytest.raises(ValueError):
        gb.PolishDoughnut(gb.JohannsenMetric())


@pytest.mark.gpu
def test_device_polish_doughnut(ensemble):
    """The torus as a device geometry: its node table crosses the ABI (GB200_GEOMETRY_THICK_TABLE), so the device interpolates
    the nodes the reference interpolates.  Main parity protocol against the oracle with the same table."""
    m = gb.KerrMetric(1.0, 0.2)
    d = gb.PolishDoughnut(m)
    oracle.set_cross_section(d.rho, d.height)
    x = [0.0, 1000.0, math.radians(75), 0.0]
    cfg = render_config(m, x, d, 2000.0, 96, 96, (-22, 22), (-12, 12), ensemble=ensemble)
    p, ic = cfg.to_c()
    assert p.geometry_kind == cabi.GEOMETRY_THICK_TABLE
    ref = oracle.trace(p, ic)
    ref_l = oracle.trace(p, ic, precision=1)
    band = oracle.band_ratio(p, ic)
    graze = ((band > 0) & (band < 1.3)) | (ref.status != ref_l.status)
    got = api.solve_tracing_problem(cfg)
    ok = ~graze
    assert graze.mean() < 0.02 and np.array_equal(np.asarray(got.status)[ok], ref.status[ok])
    hit = ok & (ref.status == cabi.STATUS_INTERSECTED)
    assert hit.sum() > 1000
    assert np.max(np.abs(got.x[1:3, hit] - ref.x[1:3, hit]) / np.maximum(np.abs(ref.x[1:3, hit]), 1e-3)) < 1e-6
    rho = got.x[1, hit] * np.abs(np.sin(got.x[2, hit]))
    z = got.x[1, hit] * np.abs(np.cos(got.x[2, hit]))
    assert np.quantile(np.abs(z - d.cross_section(rho)), 0.99) < 1e-8  # the hits sit on the torus surface


def test_shakura_sunyaev_surface_tangents():
    """test/discs/test-geometry.jl: unit tangents of the Shakura-Sunyaev surface around Kerr a = 0.9 (atol 1e-5 there)."""
    d = gb.ShakuraSunyaev(gb.KerrMetric(1.0, 0.9))
    assert np.allclose(api.cartesian_tangent_vector(d, 2.6), [0.689693957000099, 0.0, 0.724100991352412], atol=1e-5)
    assert np.allclose(api.cartesian_tangent_vector(d, 6.6), [0.9679192396299138, 0.0, 0.2512615083021063], atol=1e-5)
    assert np.allclose(api.cartesian_tangent_vector(d, 1.0), [1.0, 0.0, 0.0], atol=1e-5)
    n = api.cartesian_surface_normal(d, 6.6)
    assert abs(np.dot(n, api.cartesian_tangent_vector(d, 6.6))) < 1e-15 and n[2] > 0


def test_special_radii_and_circular_orbit_literals():
    """test/smoke-tests/special-radii.jl:27-41 (Johannsen ISCOs, atol 1e-5 there) and test/smoke-tests/circular-orbits.jl:15-24
    (sum of the circular-orbit v^phi over r = 6:0.5:10, atol 1e-6 there; the reference finds each v^phi by minimising the
    radial excursion of a traced orbit -- the closed form of `CircularOrbits` lands on the same numbers)."""
    assert api.isco(gb.JohannsenMetric(1.0, 0.998, alpha13=1.0)) == pytest.approx(2.8482863127671534, abs=1e-9)
    assert api.isco(gb.JohannsenMetric(1.0, 0.998, alpha22=1.0)) == pytest.approx(1.1306596884484472, abs=1e-9)
    assert api.isco(gb.KerrMetric(1.0, 0.0)) == 6.0 and api.isco(gb.KerrMetric(1.0, 1.0)) == pytest.approx(1.0, abs=1e-12)
    rs = np.arange(6.0, 10.01, 0.5)
    for m, literal in [(gb.KerrMetric(1.0, 0.0), 0.5432533297869712), (gb.KerrMetric(1.0, 1.0), 0.5016710246454921),
                       (gb.KerrMetric(1.0, -1.0), 0.5993458160081419), (gb.JohannsenMetric(1.0, 1.0, alpha22=1.0), 0.4980454719932759)]:
        total = sum(hostmath.circular_fourvelocity(m, r)[3] for r in rs)
        assert total == pytest.approx(literal, abs=1e-6)  # measured 2e-10 ... 9.8e-7 (the optimiser's own tolerance)


def test_host_tsit5_on_known_solutions():
    """`hostmath.tsit5_solve` (the set-up integrator behind `PolishDoughnut`): exponential decay and a harmonic oscillator with a
    discrete termination, at OrdinaryDiffEq's default tolerances; dtmax is honoured and every accepted step is saved."""
    sol = hostmath.tsit5_solve(lambda u: -u, [1.0], 2.0, dtmax=0.5)
    assert sol[-1][0] == pytest.approx(math.exp(-2.0), rel=2e-4) and len(sol) >= 5
    ts, sol = hostmath.tsit5_solve(lambda u: np.array([u[1], -u[0]]), [0.0, 1.0], 10.0, dtmax=0.05, terminate=lambda u: u[0] < 0.0,
                                   return_times=True)
    ts, x = np.array(ts), np.array([u[0] for u in sol])
    assert np.all(x[:-1] >= 0.0) and x[-1] < 0.0                  # stops at the first step that ends below zero ...
    assert math.pi < ts[-1] < math.pi + 0.05 + 1e-12              # ... i.e. within one dtmax after t = pi
    assert np.max(np.diff(ts)) <= 0.05 + 1e-15 and np.max(np.abs(x - np.sin(ts))) < 1e-6
