"""Closed-form unit checks of the oracle's building blocks (independent of any golden value)."""
import math

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from oracle import oracle

from common import c1, c3, render_config

METRICS = [(cabi.METRIC_KERR, (1.0, 0.998)), (cabi.METRIC_KERR, (1.0, 0.0)), (cabi.METRIC_JP, (1.0, 0.6, 2.0)),
           (cabi.METRIC_JP, (1.0, 0.8831, 0.4))]


@pytest.mark.parametrize("kind,mp", METRICS)
def test_metric_jacobian_matches_finite_differences(kind, mp):
    for r, th in [(3.0, 0.4), (10.0, 1.3), (1000.0, 1.047), (2.2, 2.5)]:
        g, dr, dth = oracle.metric(kind, mp, r, th)
        h = 1e-5 * r
        gp, _, _ = oracle.metric(kind, mp, r + h, th)
        gm, _, _ = oracle.metric(kind, mp, r - h, th)
        assert np.allclose(dr, (gp - gm) / (2 * h), rtol=1e-7, atol=1e-9 * np.abs(g).max())
        h = 1e-6
        gp, _, _ = oracle.metric(kind, mp, r, th + h)
        gm, _, _ = oracle.metric(kind, mp, r, th - h)
        assert np.allclose(dth, (gp - gm) / (2 * h), rtol=1e-6, atol=1e-9 * np.abs(g).max())


def test_kerr_metric_matches_textbook():
    M, a, r, th = 1.0, 0.7, 5.0, 1.1
    g, _, _ = oracle.metric(cabi.METRIC_KERR, (M, a), r, th)
    S = r * r + a * a * math.cos(th) ** 2
    D = r * r - 2 * M * r + a * a
    s2 = math.sin(th) ** 2
    want = [-(1 - 2 * M * r / S), S / D, S, s2 * (r * r + a * a + 2 * M * r * a * a * s2 / S), -2 * M * r * a * s2 / S]
    assert np.allclose(g, want, rtol=1e-14)


def test_jp_reduces_to_kerr():
    u = np.array([0.0, 6.0, 1.0, 0.3, 1.2, -0.4, 0.05, 0.07])
    assert np.allclose(oracle.rhs(cabi.METRIC_JP, (1.0, 0.9, 0.0), u), oracle.rhs(cabi.METRIC_KERR, (1.0, 0.9), u), rtol=1e-12, atol=1e-15)


def _invariants(kind, mp, x, v):
    g, _, _ = oracle.metric(kind, mp, x[1], x[2])
    E = -(g[0] * v[0] + g[4] * v[3])
    L = g[3] * v[3] + g[4] * v[0]
    norm = g[0] * v[0] ** 2 + g[1] * v[1] ** 2 + g[2] * v[2] ** 2 + g[3] * v[3] ** 2 + 2 * g[4] * v[0] * v[3]
    return E, L, norm


def test_trajectories_conserve_energy_angular_momentum_and_null_norm():
    """Physics check of RHS + integrator that needs no reference value: E, L_z and g(v,v)=0 along every ray."""
    for build in (lambda: c1(24, 24)[3], lambda: c3(24, 24)[4]):
        cfg = build()
        p, ic = cfg.to_c()
        ep = oracle.trace(p, ic)
        mp = tuple(p.metric_params)
        for i in range(0, ic.n, 7):
            E0, L0, n0 = _invariants(p.metric_kind, mp, ep.x_init[:, i], ep.v_init[:, i])
            E1, L1, n1 = _invariants(p.metric_kind, mp, ep.x[:, i], ep.v[:, i])
            assert abs(n0) < 1e-12
            scale = 1.0 + abs(ep.v[0, i])  # v^t blows up towards the horizon
            assert abs(E1 - E0) < 2e-7 * scale and abs(L1 - L0) < 2e-6 * scale * max(1.0, abs(L0))
            assert abs(n1) < 1e-6 * scale**2


def test_initial_velocity_is_null_and_matches_impact_parameters():
    _, _, _, cfg = c1(8, 8)
    p, _ = cfg.to_c()
    for al, be in [(1e-6, 1e-6), (5.0, -3.0), (-60.0, 40.0)]:
        u = oracle.initial_state(p, al, be)
        _, _, norm = _invariants(p.metric_kind, tuple(p.metric_params), u[:4], u[4:])
        assert abs(norm) < 1e-13
        assert u[5] < 0  # ingoing
        # at r = 1000 space is nearly flat: r v^theta / v^r = beta / r and r sin(theta) v^phi / v^r = alpha / r to O(M/r)
        assert 1000.0 * u[6] / u[5] == pytest.approx(be / 1000.0, rel=2e-2, abs=1e-6)
        assert 1000.0 * math.sin(p.observer[2]) * u[7] / u[5] == pytest.approx(al / 1000.0, rel=2e-2, abs=1e-5)


def test_redshift_of_equatorial_keplerian_emitter_far_field():
    """Face-on limit: for an observer on the axis g = 1/u^t(1 - ...) -> sqrt(1 - 3M/r) for Schwarzschild circular orbits."""
    m = gb.KerrMetric(1.0, 0.0)
    x = [0.0, 1e4, 1e-3, 0.0]
    cfg = render_config(m, x, gb.ThinDisc(6.0, 40.0), 2e4, 9, 9, (-30, 30), (-30, 30))
    p, ic = cfg.to_c()
    img = oracle.render(p, ic, [cabi.PF_REDSHIFT, cabi.PF_DISC_RADIUS])
    g, rho = img[0], img[1]
    ok = ~np.isnan(g)
    assert ok.sum() > 20
    assert np.allclose(g[ok], np.sqrt(1 - 3.0 / rho[ok]), rtol=2e-3)


def test_band_ratio_is_consistent_with_statuses():
    _, _, _, cfg = c1(40, 40)
    p, ic = cfg.to_c()
    ep = oracle.trace(p, ic)
    ratio = oracle.band_ratio(p, ic)
    hit = ep.status == cabi.STATUS_INTERSECTED
    assert not np.any(hit & (ratio == 0))       # a detected hit implies the transparent ray enters the disc region
    assert not np.any(~hit & (ratio > 1.3) & (ep.status != cabi.STATUS_WITHIN_INNER_BOUNDARY))  # long passages are never missed


def test_oracle_matches_frozen_vectors():
    """tests/golden/oracle_small.npz (made by tests/golden/make_golden.py once the oracle was pinned to the reference)."""
    import os

    from common import c5

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.npz"))
    for name, cfg in [("c1_24x24", c1(24, 24)[3]), ("c3_20x20", c3(20, 20)[4]), ("c5_20x20", c5(20, 20)[3])]:
        p, ic = cfg.to_c()
        imgs, ep = oracle.render(p, ic, [cabi.PF_REDSHIFT, cabi.PF_DISC_RADIUS], endpoints=True, nthreads=2)
        assert np.array_equal(ep.status, gold[name + "_status"]) and np.array_equal(ep.naccept, gold[name + "_naccept"])
        assert np.allclose(ep.x, gold[name + "_x"], rtol=1e-12, atol=1e-12) and np.allclose(ep.v, gold[name + "_v"], rtol=1e-12, atol=1e-12)
        assert np.allclose(imgs[0], gold[name + "_redshift"], rtol=1e-12, atol=0, equal_nan=True)
        assert np.allclose(imgs[1], gold[name + "_radius"], rtol=1e-12, atol=0, equal_nan=True)
