"""Per-ray datum planes: `GB200_IC_IMPACT_PARAMETERS` with x[2] = height[n] lets one launch probe
`datumplane(d, rₑ)` (src/geometry/discs/datum-plane.jl:6-19) of many emission radii at once, which is what the
thick-disc transfer-function workhorse (`_thick_workhorse`, src/transfer-functions/cunningham-transfer-functions.jl:251-300)
needs when it runs in lock step over radii."""
import math

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200.api import ImpactParameters, solve_tracing_problem, tracing_configuration
from oracle import oracle


def _fixture(n=48, seed=3):
    rng = np.random.default_rng(seed)
    m = gb.KerrMetric(1.0, 0.9)
    x = [0.0, 1000.0, math.radians(70.0), 0.0]
    rad = rng.uniform(3.0, 25.0, n)
    ang = rng.uniform(0.0, 2 * math.pi, n)
    alpha, beta = rad * np.cos(ang), rad * np.sin(ang) * math.cos(x[2])
    height = rng.uniform(0.0, 2.0, n)
    height[::7] = 0.0
    return m, x, alpha, beta, height


def _config(m, x, alpha, beta, height, plane_height):
    return tracing_configuration(m, x, ImpactParameters(alpha, beta, height), gb.DatumPlane(plane_height), 2000.0)


def test_oracle_per_ray_heights_equal_one_plane_per_ray():
    m, x, alpha, beta, height = _fixture()
    p, ic = _config(m, x, alpha, beta, height, 0.0).to_c()
    ref = oracle.trace(p, ic)
    for i in range(0, len(alpha), 5):
        p1, ic1 = _config(m, x, alpha[i:i + 1], beta[i:i + 1], None, float(height[i])).to_c()
        one = oracle.trace(p1, ic1)
        assert one.status[0] == ref.status[i]
        assert np.array_equal(one.x[:, 0], ref.x[:, i]) and np.array_equal(one.v[:, 0], ref.v[:, i])
    hit = ref.status == cabi.STATUS_INTERSECTED
    assert hit.sum() > len(alpha) // 2
    z = ref.x[1, hit] * np.cos(ref.x[2, hit])
    assert np.max(np.abs(z - height[hit])) < 1e-9


def test_heights_need_a_datum_plane():
    m, x, alpha, beta, height = _fixture(4)
    cfg = tracing_configuration(m, x, ImpactParameters(alpha, beta, height), gb.ThinDisc(0.0, 50.0), 2000.0)
    with pytest.raises(ValueError):
        cfg.to_c()


@pytest.mark.gpu
def test_device_per_ray_heights_match_the_oracle():
    m, x, alpha, beta, height = _fixture(4096, seed=11)
    cfg = _config(m, x, alpha, beta, height, 0.5)
    p, ic = cfg.to_c()
    ref = oracle.trace(p, ic)
    gps = solve_tracing_problem(cfg)
    same = gps.status == ref.status
    assert same.mean() > 0.999
    hit = same & (ref.status == cabi.STATUS_INTERSECTED)
    assert hit.sum() > 2000
    ex = np.max(np.abs(gps.x[:, hit] - ref.x[:, hit]) / np.maximum(np.abs(ref.x[:, hit]), 1.0))
    ev = np.max(np.abs(gps.v[:, hit] - ref.v[:, hit]), axis=0) / np.maximum(np.max(np.abs(ref.v[:, hit]), axis=0), 1e-12)
    assert ex < 1e-6 and ev.max() < 1e-6
    z = gps.x[1, hit] * np.cos(gps.x[2, hit])
    assert np.max(np.abs(z - height[hit])) < 1e-8
    # NULL heights keep the old meaning: every ray meets geometry_params[0]
    cfg0 = _config(m, x, alpha, beta, None, 0.5)
    g0 = solve_tracing_problem(cfg0)
    h0 = g0.status == cabi.STATUS_INTERSECTED
    assert np.max(np.abs(g0.x[1, h0] * np.cos(g0.x[2, h0]) - 0.5)) < 1e-8
