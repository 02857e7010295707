"""Line profiles through transfer-function quadrature.

Pins: the reference's own test (test/line-profiles/test-cunningham.jl:5-23,26-39: red/blue edges and unit area), and —
much tighter — agreement with the *binned* line profile of the same model (image-plane histogram, a11): two independent
algorithms (root-found rings + Jacobians + quadrature vs a plain image of 10⁵ rays) that share only the tracer."""
import math

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200 import api
from gradus_b200 import tf_integration as ti
from gradus_b200 import transfer_functions as tf

from common import OracleProber
from oracle import oracle


def test_nan_linear_interpolator_follows_the_reference():
    f = ti.NaNLinearInterpolator([0.0, 1.0, 2.0, 3.0], [1.0, 3.0, np.nan, 7.0])
    assert f(0.5) == 2.0 and f(-1.0) == -1.0  # linear extrapolation beyond the ends
    assert ti.NaNLinearInterpolator([0.0, 1.0], [1.0, 3.0])(2.0) == 5.0
    assert f(1.25) == 3.0 and f(1.75) == 0.0  # NaN knot: nearer neighbour, or the default when that one is NaN
    assert f(2.75) == 7.0 and f(2.25) == 0.0
    assert np.allclose(f(np.array([0.0, 1.0, 3.0])), [1.0, 3.0, 7.0])


def test_inverse_grid_and_branch_split():
    r = ti.inverse_grid(2.0, 50.0, 5)
    assert r[0] == pytest.approx(2.0) and r[-1] == pytest.approx(50.0) and np.all(np.diff(r) > 0)
    assert np.allclose(1 / r, np.linspace(1 / 2.0, 1 / 50.0, 5))
    # synthetic closed curve in (g✶, f): upper branch has the larger f
    th = np.linspace(-math.pi / 2, 3 * math.pi / 2, 41)[:-1]
    gs = (1 - np.cos(th - 0.1)) / 2
    f = 1.0 + 0.3 * np.sin(th - 0.1)
    ctf = tf.CunninghamTransferData(f, (gs - gs.min()) / (gs.max() - gs.min()), th.copy(), 0.5, 1.1, 7.0, th)
    br = ti.interpolate_branches(ctf)
    x = np.linspace(0.05, 0.95, 7)
    assert np.all(br.upper_f(x) > br.lower_f(x))
    assert np.allclose(br.upper_f(x) + br.lower_f(x), 2.0, atol=2e-2)
    itb = ti.InterpolatingTransferBranches([br, ti.TransferBranches(br.upper_f, br.lower_f, br.upper_t, br.lower_t, 0.4, 1.2, 9.0)])
    mid = itb(8.0)
    assert mid.gmin == pytest.approx(0.45) and mid.gmax == pytest.approx(1.15) and np.allclose(mid.upper_f(x), br.upper_f(x))


def test_quadrature_of_a_flat_transfer_function_is_analytic():
    """f ≡ 1 on both branches, ε = 1, one annulus: the bin integrals are ∫ 2 g³ / √(g✶(1−g✶)) dg, whose total over
    the whole range has a closed form; also exercises the edge rule (the integrand diverges like 1/√ at both ends)."""
    flat = ti.NaNLinearInterpolator([0.0, 1.0], [1.0, 1.0])
    zero_t = ti.NaNLinearInterpolator([0.0, 1.0], [0.0, 0.0])
    gmin, gmax = 0.5, 1.0
    brs = [ti.TransferBranches(flat, flat, zero_t, zero_t, gmin, gmax, r) for r in (5.0, 6.0)]
    itb = ti.InterpolatingTransferBranches(brs)
    grid = np.linspace(0.4, 1.1, 141)
    def S(g):
        gs = (g - gmin) / (gmax - gmin)
        with np.errstate(invalid="ignore", divide="ignore"):
            return 2 * g**3 / np.sqrt(gs * (1 - gs))

    got = ti._integrate_bins(S, grid[:-1], grid[1:], gmin, gmax, 1e-8, ti._gauss(7))
    # exact: substitute g = gmin + Δ sin²φ: ∫ 2 g³ Δ 2 dφ over φ ∈ [0, π/2]
    phi = np.linspace(0, math.pi / 2, 200001)
    exact = np.trapezoid(4 * (gmax - gmin) * (gmin + (gmax - gmin) * np.sin(phi) ** 2) ** 3, phi)
    assert got[grid[1:] <= gmin].sum() == 0 and got[grid[:-1] >= gmax].sum() == 0
    assert abs(got.sum() / exact - 1) < 2e-2  # 7-point Gauss per bin against the inverse-square-root ends
    prof = ti.integrate_lineprofile(lambda r: 1.0, itb, grid, n_radii=4)
    assert prof.sum() == pytest.approx(1.0) and prof[-1] == 0.0


def _binned_oracle(m, x, d, bins, Nr, Nt):
    cfg = api.tracing_configuration(m, x, gb.PolarPlane(gb.GeometricGrid(), Nr=Nr, Ntheta=Nt, r_max=250.0), d, (0.0, 2000.0),
                                    callback=gb.domain_upper_hemisphere())
    p, ic = cfg.to_c()
    emis = cabi.Emissivity(cabi.EMISSIVITY_POWERLAW, 0, 3.0, None, None)
    return oracle.lineprofile(p, ic, emis, bins, cabi.LineProfileOpts(gb.isco(m) + 1e-2, 50.0, 1, 0))


def _edges(bins, y):
    nz = np.nonzero(y > 0)[0]
    return bins[nz[0]], bins[nz[-1]]


def test_transfer_function_line_profile_with_the_oracle_tracer():
    m = gb.KerrMetric(1.0, 0.6)
    x = [0.0, 1000.0, math.radians(60), 0.0]
    d = gb.ThinDisc(0.0, 250.0)
    bins = np.linspace(0.1, 1.3, 100)
    _, y = ti.lineprofile_transfer_functions(bins, lambda r: r**-3.0, m, x, d, N=40, num_re=24, prober=OracleProber(m, x, d))
    g_low, g_high = _edges(bins, y)
    assert abs(g_low - 0.355) < 0.05 and abs(g_high - 1.2) < 0.05 and y.sum() == pytest.approx(1.0)
    yb = _binned_oracle(m, x, d, bins, 160, 320)
    assert np.abs(y - yb).sum() < 0.04


def test_thick_disc_line_profile_with_the_oracle_tracer():
    """ShakuraSunyaev disc: transfer functions on per-radius datum planes + visibility re-traces + Jacobians on the disc
    surface, against the binned image-plane histogram of the same thick disc.  The thin-disc profile of the same
    black hole is 0.28 away in L1, so the agreement is a test of the thick-disc machinery, not of the spacetime."""
    m = gb.KerrMetric(1.0, 0.6)
    x = [0.0, 1000.0, math.radians(60), 0.0]
    d = gb.ShakuraSunyaev(m, eddington_ratio=0.3)
    bins = np.linspace(0.1, 1.3, 100)
    pr = OracleProber(m, x, d, chart=gb.chart_for_metric(m, 2 * x[1]))
    _, y = ti.lineprofile_transfer_functions(bins, lambda r: r**-3.0, m, x, d, N=40, num_re=24, prober=pr, min_re=gb.isco(m) + 1e-2)
    assert y.sum() == pytest.approx(1.0)
    yb = _binned_oracle(m, x, d, bins, 160, 320)
    y_thin = _binned_oracle(m, x, gb.ThinDisc(0.0, 250.0), bins, 160, 320)
    assert np.abs(y - yb).sum() < 0.05
    assert np.abs(yb - y_thin).sum() > 0.2


# --------------------------------------------------------------------------- device
@pytest.mark.gpu
def test_thick_disc_line_profile_on_the_device():
    m = gb.KerrMetric(1.0, 0.6)
    x = [0.0, 1000.0, math.radians(60), 0.0]
    d = gb.ShakuraSunyaev(m, eddington_ratio=0.3)
    bins = np.linspace(0.1, 1.3, 100)
    _, y = ti.lineprofile_transfer_functions(bins, lambda r: r**-3.0, m, x, d, N=40, num_re=30, min_re=gb.isco(m) + 1e-2)
    _, yb = gb.lineprofile(bins, gb.PowerLawEmissivity(3.0), m, x, d, gb.BinningMethod(), min_re=gb.isco(m) + 1e-2, max_re=50.0)
    assert y.sum() == pytest.approx(1.0) and np.abs(y - yb).sum() < 0.04


@pytest.mark.gpu
@pytest.mark.parametrize("m, g_low_ref", [(gb.KerrMetric(1.0, 0.6), 0.355), (gb.JohannsenPsaltisMetric(1.0, 0.6, 2.0), 0.27)],
                         ids=["kerr", "johannsen_psaltis"])
def test_transfer_function_line_profile_on_the_device(m, g_low_ref):
    """test/line-profiles/test-cunningham.jl, reference settings (N = 40, numrₑ = 30), and the cross-check against
    the device's own binned line profile at full default resolution (450 × 1300 rays)."""
    x = [0.0, 1000.0, math.radians(60), 0.0]
    d = gb.ThinDisc(0.0, 250.0)
    bins = np.linspace(0.1, 1.3, 100)
    _, y = ti.lineprofile_transfer_functions(bins, lambda r: r**-3.0, m, x, d, N=40, num_re=30)
    g_low, g_high = _edges(bins, y)
    assert abs(g_low - g_low_ref) < 0.05 and abs(g_high - 1.2) < 0.05 and y.sum() == pytest.approx(1.0)
    _, yb = gb.lineprofile(bins, gb.PowerLawEmissivity(3.0), m, x, d, gb.BinningMethod(), min_re=gb.isco(m) + 1e-2, max_re=50.0)
    assert np.abs(y - yb).sum() < 0.03
