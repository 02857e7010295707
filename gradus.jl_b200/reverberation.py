"""Lag (2-D) transfer functions (rest of SURVEY 8 row f3): `lagtransfer` / `binflux`
(src/transfer-functions/transfer-functions-2d.jl:100-242) and `AnalyticRadialDiscProfile` (src/corona/analytic.jl:1-40).

Both legs are endpoint traces on the device: source → disc through `tracecorona` (explicit ICs), observer → disc through
one fused `gb200_render` over the image plane that returns (status, g, ρ, t) per ray, so no endpoint arrays cross the
bus.  Binning is the reference's `Buckets.Simple` in two dimensions (slot i takes bins[i] ≤ v < bins[i+1], clamped)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _cabi as cabi
from . import api, corona
from .tf_integration import NaNLinearInterpolator


@dataclass
class AnalyticRadialDiscProfile:
    """ε(r) given as a function, arrival time interpolated from the coronal geodesics (corona/analytic.jl:1-40)."""

    eps: Callable
    t: NaNLinearInterpolator

    @classmethod
    def from_corona(cls, emissivity, cg: corona.CoronaGeodesics):
        x = cg.geodesic_points["x"]
        rho = x[1] * np.sin(x[2])
        J = np.argsort(rho, kind="stable")
        return cls(emissivity, NaNLinearInterpolator(rho[J], x[0][J]))

    def emissivity_at(self, r):
        return self.eps(np.asarray(r))

    def coordtime_at(self, r):
        return self.t(np.clip(r, self.t.t[0], self.t.t[-1]))


@dataclass
class LagTransferFunction:
    """transfer-functions/types.jl `LagTransferFunction`; the observer→disc leg is kept as the per-ray quantities
    `binflux` needs (g, ρ, t) instead of full geodesic points."""

    max_t: float
    x: np.ndarray
    image_plane_areas: np.ndarray
    coronal_geodesics: corona.CoronaGeodesics
    g: np.ndarray
    rho: np.ndarray
    t: np.ndarray

    @property
    def observer_to_disc_count(self):
        return len(self.g)


def unnormalized_areas(plane: api.PolarPlane):
    """image-planes/planes.jl:108-115: r² per ray, r the fastest index."""
    if not isinstance(plane, api.PolarPlane):
        raise ValueError("lagtransfer is implemented for PolarPlane image planes")
    if isinstance(plane.grid, api.GeometricGrid):
        rs = plane.r_min * (plane.r_max / plane.r_min) ** (np.arange(plane.Nr) / (plane.Nr - 1))
    elif isinstance(plane.grid, api.InverseGrid):
        rs = (1.0 / np.linspace(1.0 / plane.r_max, 1.0 / plane.r_min, plane.Nr))[::-1]
    else:
        rs = np.linspace(plane.r_min, plane.r_max, plane.Nr)
    return np.tile(rs**2, plane.Ntheta)


def _device_evaluator(config, pfs):
    return api.apply_point_functions(config, pfs)


def lagtransfer(m, x, d, model, *, plane=None, max_t=None, n_samples=10_000, sampler=None, callback="default",
                ensemble=None, solver=None, evaluator=None, **solver_opts) -> LagTransferFunction:
    """`lagtransfer(m, u, d, model; plane, max_t, n_samples, sampler, ...)` (transfer-functions-2d.jl:163-218).
    `solver` / `evaluator` replace the device tracer in the CPU tests (oracle)."""
    x = np.asarray(x, np.float64)
    if plane is None:
        plane = api.PolarPlane(api.GeometricGrid(), Nr=800, Ntheta=800, r_max=50.0)
    if max_t is None:
        max_t = 2 * x[1]
    if sampler is None:
        sampler = corona.EvenSampler("both", "golden")  # the reference's default draws from its RNG
    if callback == "default":
        callback = api.domain_upper_hemisphere()
    ce = corona.tracecorona(m, d, model, lambda_max=max_t, n_samples=n_samples, sampler=sampler, callback=callback,
                            ensemble=ensemble, solver=solver, **solver_opts)
    config = api.tracing_configuration(m, x, plane, d, (0.0, max_t), chart=api.chart_for_metric(m, 1.1 * x[1]),
                                       callback=callback, ensemble=ensemble, **solver_opts)
    C = api.ConstPointFunctions
    pfs = [api.PointFunction("status"), C.redshift(m, x) @ C.filter_intersected(), C.radius() @ C.filter_intersected(),
           C.coordinate_time() @ C.filter_intersected()]
    status, g, rho, t = (evaluator or _device_evaluator)(config, pfs)
    I = status == cabi.STATUS_INTERSECTED
    return LagTransferFunction(float(max_t), x, unnormalized_areas(plane)[I], ce, g[I], rho[I], t[I])


def _simple_bucket_index(bins, v):
    return np.clip(np.searchsorted(bins, v, side="right") - 1, 0, len(bins) - 1)


def bucket2d(x, y, w, xbins, ybins, ensemble=None):
    """`bucket(x, y, w, xbins, ybins; reduction = sum)` (Buckets.Simple in both axes).  With an `ensemble` the histogram is
    accumulated on the device (`gb200_bucket2d`, order-independent fixed-point sums); without, on the host."""
    xbins = np.ascontiguousarray(xbins, np.float64)
    ybins = np.ascontiguousarray(ybins, np.float64)
    if ensemble is None:
        out = np.zeros((len(xbins), len(ybins)))
        np.add.at(out, (_simple_bucket_index(xbins, x), _simple_bucket_index(ybins, y)), w)
        return out
    import ctypes as C

    x, y, w = (np.ascontiguousarray(a, np.float64) for a in (x, y, w))
    out = np.zeros((len(xbins), len(ybins)))
    ctx = ensemble.ctx(ensemble.devices[0])
    cabi.check(cabi.load().gb200_bucket2d(ctx, len(x), cabi.dptr(x), cabi.dptr(y), cabi.dptr(w), cabi.dptr(xbins), len(xbins), cabi.dptr(ybins),
                                          len(ybins), cabi.dptr(out.reshape(-1))), ctx)
    return out


def bin_transfer_function(time_delays, energy, flux, *, N_E=300, N_t=300, energy_lims=None, time_lims=None, ensemble=None):
    """transfer-functions-2d.jl:100-122 → (time_bins, energy_bins, tf[N_E, N_t]) with empty cells NaN."""
    energy_lims = (np.min(energy), np.max(energy)) if energy_lims is None else energy_lims
    time_lims = (np.min(time_delays), np.max(time_delays)) if time_lims is None else time_lims
    eb = np.linspace(energy_lims[0], energy_lims[1], N_E)
    tb = np.linspace(time_lims[0], time_lims[1], N_t)
    de, dt = eb[1] - eb[0], tb[1] - tb[0]
    out = bucket2d(energy, time_delays, flux, eb, tb, ensemble)
    out /= de * dt
    out[out == 0.0] = np.nan
    return tb, eb, out


def binflux(tf: LagTransferFunction, profile=None, *, E0=6.4, t0=None, **kwargs):
    """`binflux(tf, profile; E₀, t0, N_E, N_t, ...)` (transfer-functions-2d.jl:220-242)."""
    if profile is None:
        profile = AnalyticRadialDiscProfile.from_corona(lambda r: r**-3.0, tf.coronal_geodesics)
    if t0 is None:
        t0 = tf.x[1]
    t = profile.coordtime_at(tf.rho) + tf.t
    eps = profile.emissivity_at(tf.rho)
    f = tf.g**3 * eps * tf.image_plane_areas
    F = f / f.sum()
    tb, eb, td = bin_transfer_function(t, tf.g * E0, F, **kwargs)
    return tb - t0, eb, td


def continuum_time(m, x, model, **kwargs):
    """`continuum_time(m, x, model)` (src/reverberation.jl:81-93): coordinate time of the direct corona-to-observer light
    path, found backwards as the geodesic from the observer that reaches the corona's position (`optimize_for_target`
    with the upper-hemisphere callback and a chart of twice the observer's radius)."""
    pos, _ = corona.sample_position_velocity(m, model)
    kwargs.setdefault("chart", api.chart_for_metric(m, 2.0 * float(x[1])))
    kwargs.setdefault("callback", api.domain_upper_hemisphere())
    _, _, gp, _ = api.optimize_for_target(pos[1:], m, x, **kwargs)
    return float(gp["x"][0])



def lag_frequency(t, f, *, flo=5e-5, R=1.0):
    """`lag_frequency(t, f::AbstractMatrix; flo)` (src/reverberation.jl:29-49): the impulse response ψ(t) = Σ_E f[E, t] (NaNs as
    zeros) is padded with zeros out to 1 / flo on the same time step, Fourier transformed, and the phase
    atan(Im ψ̂ / (1 + Re ψ̂)) of the positive frequencies becomes a time lag φ / (2π ν).  Returns (frequencies, −lag); the
    zero-frequency entry is 0 / 0 as in the reference.  Host FFT of a ~20 000-point array."""
    t = np.asarray(t, np.float64)
    f = np.asarray(f, np.float64)
    psi = np.nansum(f, axis=0) if f.ndim == 2 else np.where(np.isnan(f), 0.0, f)
    dt = t[1] - t[0]
    t_ext = np.arange(t.min(), 1.0 / flo + 0.5 * dt * 1e-9, dt)  # range(minimum(t), 1 / flo, step = Δt)
    psi_ext = np.zeros(t_ext.size)
    psi_ext[: psi.size] = psi
    freq = np.fft.fftfreq(t_ext.size, dt)
    F = R * np.fft.fft(psi_ext)
    n = t_ext.size // 2
    phase = np.arctan(F[:n].imag / (1.0 + F[:n].real))
    with np.errstate(divide="ignore", invalid="ignore"):
        lag = phase / (2.0 * np.pi * freq[:n])
    return freq[:n], -lag
