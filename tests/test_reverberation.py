"""Lag transfer functions (second half) against the reference's own test, literal for literal
(test/transfer-functions/test-2d.jl): hit counts of both legs, the binned 2-D transfer function, and the
semi-analytic path (sampled-sky emissivity profile × Cunningham transfer functions × time-resolved quadrature)."""
import math

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200 import api, corona
from gradus_b200 import reverberation as rv
from gradus_b200 import tf_integration as ti

import common
from oracle import oracle

REF_OBSERVER_TO_DISC = 337          # test-2d.jl:25
REF_SOURCE_TO_DISC = 58             # test-2d.jl:26
REF_BINFLUX_SUM = 3.9126785201177956  # test-2d.jl:32, atol 1e-2
REF_LAG_ROW_40 = 0.021759503160585468  # test-2d.jl:65, atol 1e-4


def fixture():
    m = gb.KerrMetric(1.0, 0.998)
    x = [0.0, 1e6, math.radians(30), 0.0]
    plane = gb.PolarPlane(gb.GeometricGrid(), Nr=20, Ntheta=20)
    d = gb.ThinDisc(gb.isco(m), 500.0)
    model = corona.LampPostModel(h=10.0, theta=math.radians(0.0001))
    return m, x, plane, d, model


def oracle_evaluator(config, pfs):
    p, ic = config.to_c()
    return oracle.render(p, ic, [f.kind() for f in pfs])


def run_reference_test(solver=None, evaluator=None, prober_cls=None):
    m, x, plane, d, model = fixture()
    sampler = corona.EvenSampler("both", "golden")
    lt = rv.lagtransfer(m, x, d, model, plane=plane, n_samples=100, sampler=sampler, solver=solver, evaluator=evaluator)
    t, E, f = rv.binflux(lt, N_t=100, N_E=100)
    prof = corona.emissivity_profile(m, d, model, n_samples=5000, sampler=sampler, solver=solver)
    radii = ti.inverse_grid(gb.isco(m), 100.0, 5)
    d2 = gb.ThinDisc(0.0, 500.0)
    kw = {} if prober_cls is None else {"prober": prober_cls(m, x, d2)}
    itb = ti.transferfunctions(m, x, d2, radii=radii, **kw)
    flux = ti.integrate_lagtransfer(prof, itb, np.linspace(0.0, 1.5, 100), np.linspace(0.0, 150.0, 100), t0=x[1], n_radii=1000,
                                    rmin=radii.min(), rmax=radii.max())
    return lt, (t, E, f), flux


def check(lt, binned, flux, tol_sum, tol_row):
    t, E, f = binned
    assert lt.observer_to_disc_count == REF_OBSERVER_TO_DISC
    assert lt.coronal_geodesics.geodesic_points["x"].shape[1] == REF_SOURCE_TO_DISC
    assert f.shape == (100, 100) and 0 < t[0] < 100 and 0 < E[0] < E[-1] < 6.4 * 1.6
    assert abs(np.nansum(f) - REF_BINFLUX_SUM) < tol_sum
    assert flux.sum() == pytest.approx(1.0) and np.all(flux[-1] == 0)
    assert abs(flux[39].sum() - REF_LAG_ROW_40) < tol_row


@pytest.fixture
def oracle_plunging_kerr():
    m = gb.KerrMetric(1.0, 0.998)
    key = (type(m).__name__, m.params())
    saved = api._PLUNGING_CACHE.get(key)
    api._PLUNGING_CACHE[key] = api.PlungingInterpolation(*common.oracle_plunging_table(cabi.METRIC_KERR, [1.0, 0.998]))
    yield
    if saved is None:
        api._PLUNGING_CACHE.pop(key, None)
    else:
        api._PLUNGING_CACHE[key] = saved


def test_two_dimensional_simple_bucket():
    tb, eb, td = rv.bin_transfer_function(np.array([0.0, 0.49, 1.0, 1.0]), np.array([2.0, 2.0, 3.0, 2.5]), np.array([1.0, 2.0, 4.0, 8.0]),
                                          N_E=3, N_t=3)
    assert np.allclose(eb, [2.0, 2.5, 3.0]) and np.allclose(tb, [0.0, 0.5, 1.0])
    want = np.full((3, 3), np.nan)
    want[0, 0] = 3.0 / 0.25   # (E=2, t=0) and (E=2, t=0.49) share the lower-edge-labelled cell
    want[2, 2] = 4.0 / 0.25   # values on the last edge go to the last slot
    want[1, 2] = 8.0 / 0.25
    assert np.array_equal(np.isnan(td), np.isnan(want)) and np.allclose(td[~np.isnan(td)], want[~np.isnan(want)])


def test_reference_reverberation_literals_with_the_oracle_tracer(oracle_plunging_kerr):
    lt, binned, flux = run_reference_test(solver=common.oracle_solver, evaluator=oracle_evaluator, prober_cls=common.OracleProber)
    check(lt, binned, flux, tol_sum=1e-4, tol_row=1e-5)  # the reference quotes 1e-2 and 1e-4


@pytest.mark.gpu
def test_reference_reverberation_literals_on_the_device():
    lt, binned, flux = run_reference_test()
    check(lt, binned, flux, tol_sum=1e-4, tol_row=1e-5)


@pytest.mark.gpu
def test_device_bucket2d_equals_the_host_bucket_and_ignores_the_order(ensemble):
    """`gb200_bucket2d` (bin_transfer_function's 2-D `bucket`, transfer-functions-2d.jl:100-122): the host histogram to the
    rounding of a float sum, and bit-identical when the samples are shuffled (fixed-point accumulation)."""
    from gradus_b200 import reverberation as rv

    rng = np.random.default_rng(2)
    n = 400_000
    t = rng.uniform(900.0, 1400.0, n)
    e = 6.4 * rng.uniform(0.2, 1.4, n)
    f = rng.uniform(0.0, 1.0, n) ** 3
    f /= f.sum()
    f[::1000] = np.nan  # samples without a value are skipped
    eb, tb = np.linspace(e.min(), e.max(), 300), np.linspace(t.min(), t.max(), 300)
    ok = np.isfinite(f)
    host = rv.bucket2d(e[ok], t[ok], f[ok], eb, tb)
    dev = rv.bucket2d(e, t, f, eb, tb, ensemble)
    assert np.abs(dev - host).max() < 1e-15 and abs(dev.sum() - host.sum()) < 1e-14
    perm = rng.permutation(n)
    assert np.array_equal(rv.bucket2d(e[perm], t[perm], f[perm], eb, tb, ensemble), dev)
    tb2, eb2, td = rv.bin_transfer_function(t[ok], e[ok], f[ok], ensemble=ensemble)
    _, _, td_host = rv.bin_transfer_function(t[ok], e[ok], f[ok])
    assert np.array_equal(np.isnan(td), np.isnan(td_host)) and np.nanmax(np.abs(td - td_host)) < 1e-12 * np.nanmax(td_host)


# --------------------------------------------------------------------------- test/smoke-tests/reverberation.jl
REF_FREQ_SUM = 2449.8787687490535   # reverberation.jl:44, rtol 1e-2
REF_TAU_132 = 9.322742661315855     # reverberation.jl:45, rtol 1e-2


def run_lag_frequency_smoke(solver=None, prober_cls=None, tracer=None):
    """The reference's lag-frequency smoke test end to end: lamp-post emissivity profile (500 rays), continuum time through
    the target solver, ten Cunningham transfer functions on the inverse grid, time-resolved quadrature, FFT."""
    m = gb.KerrMetric(1.0, 0.998)
    x = [0.0, 10_000.0, math.radians(45), 0.0]
    d = gb.ThinDisc(0.0, float("inf"))
    model = corona.LampPostModel()
    radii = ti.inverse_grid(gb.isco(m), 100.0, 10)
    kw = {} if prober_cls is None else {"prober": prober_cls(m, x, d)}
    itb = ti.transferfunctions(m, x, d, radii=radii, beta0=2.0, **kw)
    prof = corona.emissivity_profile(m, d, model, n_samples=500, solver=solver)
    tkw = {} if tracer is None else {"tracer": tracer, "grid": 17}
    t0 = rv.continuum_time(m, x, model, **tkw)
    bins, tbins = np.linspace(0.0, 1.5, 100), np.linspace(0.0, 100.0, 100)
    flux = ti.integrate_lagtransfer(prof, itb, bins, tbins, t0=t0, n_radii=100, h=1e-8, rmin=radii.min(), rmax=radii.max())
    flux[flux == 0] = np.nan
    return t0, rv.lag_frequency(tbins, flux)


def check_lag_frequency(t0, freq, tau):
    assert freq.sum() == pytest.approx(REF_FREQ_SUM, rel=1e-2)
    assert tau[131] == pytest.approx(REF_TAU_132, rel=1e-2)
    # light travel time from h = 5 on the axis to an observer at r = 10^4, 45 degrees: flat-space distance + Shapiro delay
    flat = math.sqrt(1e8 + 25.0 - 2 * 1e4 * 5.0 * math.cos(math.radians(45)))
    assert flat < t0 < flat + 4.0 * math.log(1e4 / 5.0) + 5.0


def test_lag_frequency_smoke_with_the_oracle_tracer(oracle_plunging_kerr):
    t0, (freq, tau) = run_lag_frequency_smoke(common.oracle_solver, common.OracleProber, common.oracle_target_tracer)
    check_lag_frequency(t0, freq, tau)


@pytest.mark.gpu
def test_lag_frequency_smoke_on_the_device():
    t0, (freq, tau) = run_lag_frequency_smoke()
    check_lag_frequency(t0, freq, tau)
