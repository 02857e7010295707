"""The generic-scalar integrator on the device (gb200_generic.cuh / gb200_dual.cu): forward-mode traces against the
oracle's independent dual-number trace, against finite differences, against the throughput kernel, and the path recorder
(the N = 0 instantiation) against the ensemble kernel's end points."""
import ctypes as C
import math

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200 import api

pytestmark = pytest.mark.gpu


def seeds(n, nd, theta=None):
    if nd == 1:
        return np.cos(theta)[None, :], np.sin(theta)[None, :]
    one, zero = np.ones(n), np.zeros(n)
    return np.stack([one, zero]), np.stack([zero, one])


def problem(m, x, d, ens, al, be, chart=None, **kw):
    cfg = api.tracing_configuration(m, x, api.ImpactParameters(al, be), d, 2 * x[1], chart=chart or gb.chart_for_metric(m, 2 * x[1]),
                                    ensemble=ens, **kw)
    return cfg


CASES = [
    ("kerr_datum", gb.KerrMetric(1.0, 0.998), [0.0, 1e4, math.radians(30), 0.0], lambda m: gb.DatumPlane(0.0), {}),
    ("kerr_datum_far", gb.KerrMetric(1.0, 0.998), [0.0, 1e5, math.radians(30), 0.0], lambda m: gb.DatumPlane(0.0), {}),
    ("kerr_retro_datum", gb.KerrMetric(1.0, -0.6), [0.0, 1e4, math.radians(75), 0.0], lambda m: gb.DatumPlane(0.0), {}),
    ("kerr_thick", gb.KerrMetric(1.0, 0.998), [0.0, 1e4, math.radians(75), 0.0], lambda m: gb.ShakuraSunyaev(m), {"callback": gb.domain_upper_hemisphere()}),
    ("jp_datum", gb.JohannsenPsaltisMetric(1.0, 0.6, 2.0), [0.0, 1e4, math.radians(60), 0.0], lambda m: gb.DatumPlane(0.0), {}),
]


@pytest.mark.parametrize("name,m,x,geom,kw", CASES, ids=[c[0] for c in CASES])
def test_forward_mode_trace_matches_the_oracle(ensemble, name, m, x, geom, kw):
    from oracle import oracle

    rng = np.random.default_rng(7)
    n = 192
    rr = rng.uniform(2.5, 30.0, n)
    th = rng.uniform(0, 2 * math.pi, n)
    al, be = rr * np.cos(th), rr * np.sin(th)
    d = geom(m)
    cfg = problem(m, x, d, ensemble, al, be, **kw)
    p, _ = cfg.to_c()
    plunging = None
    if not isinstance(m, gb.KerrMetric):
        plunging = api.interpolate_plunging_velocities(m, ensemble)
    for nd in (1, 2):
        da, db = seeds(n, nd, th)
        for norm_mode in (cabi.DUAL_NORM_WITH_PARTIALS, cabi.DUAL_NORM_VALUES_ONLY):
            dev = api.trace_dual(cfg, cabi.DualArrays(al, be, da, db), norm_mode, plunging=plunging)
            orc = oracle.trace_dual(p, cabi.DualArrays(al, be, da, db), norm_mode, plunging=plunging.c if plunging is not None else None)
            same = dev.status == orc.status
            assert same.mean() > 0.98, name
            hit = same & (dev.status == cabi.STATUS_INTERSECTED)
            assert hit.sum() > 40, name
            assert np.array_equal(np.isnan(dev.g[hit]), np.isnan(orc.g[hit]))
            fin = hit & np.isfinite(orc.g)
            # values: the protocol of the main path (1e-6; from r = 1e5 a relative tolerance of 1e-9 is an absolute 1e-4 per
            # step in r, and two correct step sequences land 3e-5 apart on the disc: measured 1.4e-6 relative);
            # partials: 1e-4 of their scale (measured 3e-5)
            assert np.max(np.abs(dev.rho[fin] / orc.rho[fin] - 1)) < (3e-6 if x[1] > 5e4 else 1e-6), name
            assert np.max(np.abs(dev.g[fin] - orc.g[fin])) < 1e-6, name
            for k in range(nd):
                sc_r = np.maximum(np.abs(orc.drho[k][fin]), 1e-2)
                sc_g = np.maximum(np.abs(orc.dg[k][fin]), 1e-3)
                assert np.max(np.abs(dev.drho[k][fin] - orc.drho[k][fin]) / sc_r) < 1e-4, (name, nd, k)
                assert np.max(np.abs(dev.dg[k][fin] - orc.dg[k][fin]) / sc_g) < 1e-4, (name, nd, k)
            assert abs(int(dev.naccept.sum()) - int(orc.naccept.sum())) < 0.01 * orc.naccept.sum()


def test_values_only_norm_retraces_the_plain_render(ensemble):
    """With the partials kept out of the error norm the dual trace takes the plain trace's steps: its values must be
    those of `gb200_render` (the throughput kernel) on the same rays."""
    m, x = gb.KerrMetric(1.0, 0.9), [0.0, 1e4, math.radians(50), 0.0]
    rng = np.random.default_rng(3)
    n = 256
    rr, th = rng.uniform(2.0, 40.0, n), rng.uniform(0, 2 * math.pi, n)
    al, be = rr * np.cos(th), rr * np.sin(th)
    cfg = problem(m, x, gb.DatumPlane(0.0), ensemble, al, be)
    da, db = seeds(n, 2)
    dev = api.trace_dual(cfg, cabi.DualArrays(al, be, da, db), cabi.DUAL_NORM_VALUES_ONLY)
    pfs = [gb.ConstPointFunctions.redshift(m, x) @ gb.ConstPointFunctions.filter_intersected(), gb.ConstPointFunctions.radius(),
           api.PointFunction("status")]
    img = api.apply_point_functions(cfg, pfs)
    assert np.array_equal(dev.status, img[2].astype(np.int32))
    hit = dev.status == cabi.STATUS_INTERSECTED
    assert hit.sum() > 100
    # same steps up to rounding: the two kernels order their arithmetic differently, and a last-bit difference in an error
    # estimate moves the following step sizes (measured 7e-8 in g; the protocol of the main path asks for 1e-6)
    assert np.max(np.abs(dev.g[hit] - img[0][hit])) < 1e-6
    assert np.max(np.abs(dev.rho[hit] / img[1][hit] - 1)) < 1e-6
    assert abs(int(dev.naccept.sum()) - int(api.solve_tracing_problem(cfg).naccept.sum())) <= 0.002 * dev.naccept.sum()


def test_jacobian_against_finite_differences_of_the_throughput_kernel(ensemble):
    m, x = gb.KerrMetric(1.0, 0.998), [0.0, 1e4, math.radians(40), 0.0]
    rng = np.random.default_rng(11)
    n = 128
    rr, th = rng.uniform(3.0, 25.0, n), rng.uniform(0, 2 * math.pi, n)
    al, be = rr * np.cos(th), rr * np.sin(th)
    cfg = problem(m, x, gb.DatumPlane(0.0), ensemble, al, be, abstol=1e-13, reltol=1e-13)
    da, db = seeds(n, 2)
    dev = api.trace_dual(cfg, cabi.DualArrays(al, be, da, db))
    h = 2e-5 * rr
    pfs = [gb.ConstPointFunctions.redshift(m, x) @ gb.ConstPointFunctions.filter_intersected(), gb.ConstPointFunctions.radius()]
    cfd = problem(m, x, gb.DatumPlane(0.0), ensemble, np.concatenate([al + h, al - h, al, al]), np.concatenate([be, be, be + h, be - h]),
                  abstol=1e-13, reltol=1e-13)
    img = api.apply_point_functions(cfd, pfs)
    g, rho = img[0], img[1]
    dga, dgb = (g[:n] - g[n:2 * n]) / (2 * h), (g[2 * n:3 * n] - g[3 * n:]) / (2 * h)
    dra, drb = (rho[:n] - rho[n:2 * n]) / (2 * h), (rho[2 * n:3 * n] - rho[3 * n:]) / (2 * h)
    ok = np.isfinite(dga) & np.isfinite(dgb) & (dev.status == cabi.STATUS_INTERSECTED)
    assert ok.sum() > 60
    det_fd = dra * dgb - drb * dga
    det = dev.drho[0] * dev.dg[1] - dev.drho[1] * dev.dg[0]
    assert np.max(np.abs(det[ok] / det_fd[ok] - 1)) < 2e-5


def test_batch_equals_single_calls(ensemble):
    cells = [(0.998, 30.0), (0.5, 60.0), (0.0, 80.0)]
    rng = np.random.default_rng(5)
    cfgs, arrs, singles = [], [], []
    for a, inc in cells:
        m, x = gb.KerrMetric(1.0, a), [0.0, 1e4, math.radians(inc), 0.0]
        n = int(rng.integers(20, 90))
        rr, th = rng.uniform(3.0, 25.0, n), rng.uniform(0, 2 * math.pi, n)
        al, be = rr * np.cos(th), rr * np.sin(th)
        cfg = problem(m, x, gb.DatumPlane(0.0), ensemble, al, be)
        cfgs.append(cfg)
        arrs.append(cabi.DualArrays(al, be, *seeds(n, 1, th)))
        singles.append(api.trace_dual(cfg, cabi.DualArrays(al, be, *seeds(n, 1, th))))
    api.trace_dual_batch(cfgs, arrs)
    for b, s_ in zip(arrs, singles):
        assert np.array_equal(b.status, s_.status) and np.array_equal(b.g, s_.g, equal_nan=True) and np.array_equal(b.drho, s_.drho)
        assert np.array_equal(b.x, s_.x) and np.array_equal(b.naccept, s_.naccept)


def test_path_recorder_ends_where_the_ensemble_kernel_ends(ensemble):
    """`gb200_trace_path` (N = 0 instantiation of the generic integrator) against `gb200_trace` on the same initial state:
    one integrator semantics in two kernels."""
    lib = cabi.load()
    for m, mu in [(gb.KerrMetric(1.0, 0.9), 0.0), (gb.KerrMetric(1.0, 0.5), 1.0), (gb.JohannsenPsaltisMetric(1.0, 0.6, 1.0), 0.0)]:
        x0 = np.array([0.0, 60.0, 1.1, 0.3])
        for v0 in ([0.0, -1.0, 0.002, 0.0009], [0.0, -0.6, -0.01, 0.004], [0.0, 0.3, 0.01, 0.012]):
            cfg = api.tracing_configuration(m, x0, np.array([v0]), 4000.0, ensemble=ensemble, mu=mu)
            gp = api.solve_tracing_problem(cfg)
            p, _ = cfg.to_c()
            cap = 4096
            lam, u = np.zeros(cap), np.zeros(cap * 8)
            nrows, status = C.c_int32(), C.c_int32()
            u0 = np.concatenate([x0, v0])
            cabi.check(lib.gb200_trace_path(ensemble.ctx(0), C.byref(p), cabi.dptr(u0), cap, cabi.dptr(lam), cabi.dptr(u), C.byref(nrows), C.byref(status)),
                       ensemble.ctx(0))
            rows = nrows.value
            assert 2 < rows <= cap and status.value == gp.status[0]
            assert rows - 1 == gp.naccept[0]
            end = u.reshape(-1, 8)[rows - 1]
            assert abs(lam[rows - 1] - gp.lambda_max[0]) <= 1e-9 * abs(gp.lambda_max[0])
            ref = np.concatenate([gp.x[:, 0], gp.v[:, 0]])
            assert np.max(np.abs(end - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-7
            assert np.all(np.diff(lam[:rows]) > 0)
