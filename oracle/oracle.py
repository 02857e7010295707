"""ctypes wrapper of the CPU oracle (oracle/gradus_oracle.cpp).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  It reuses the POD struct layouts of include/gradus_b200.h (via the
package's ctypes mirror) so the very same problem description can be handed to the
oracle and to the CUDA library."""
import ctypes as C
import os
import subprocess

import numpy as np

import gradus_b200._cabi as cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "gradus_oracle.cpp")
    if force or not os.path.exists(LIB) or not os.path.exists(os.path.join(_HERE, "liboracle_fast.so")) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
    return _lib


LIB_FAST = os.path.join(_HERE, "liboracle_fast.so")
_lib_fast = None


def lib_fast():
    """The timing variant (-O3, FMA, closed-form Kerr right-hand side): bench.py's "port_optimised" baseline only."""
    global _lib_fast
    if _lib_fast is None:
        if not os.path.exists(LIB_FAST):
            build(force=True)
        _lib_fast = C.CDLL(LIB_FAST)
        assert _lib_fast.oracle_is_fast_variant() == 1
    return _lib_fast


def render_fast(problem, ic, pointfns, rng=None, nthreads=0):
    """`render` through the timing variant (no end points, no plunging table)."""
    rng = _range(ic, rng)
    pfs = np.asarray(pointfns, np.int32)
    imgs = np.zeros((len(pfs), rng.count))
    ptrs = (cabi._dp * len(pfs))(*[cabi.dptr(imgs[k]) for k in range(len(pfs))])
    rc = lib_fast().oracle_render(C.byref(problem), C.byref(ic), C.byref(rng), nthreads, 0, cabi.iptr(pfs), len(pfs), None, ptrs, None)
    assert rc == 0
    return imgs


def rhs_fast(kind, params, u):
    u = np.ascontiguousarray(u, np.float64)
    du = np.zeros(8)
    mp = _mp(params)
    lib_fast().oracle_rhs(kind, cabi.dptr(mp), cabi.dptr(u), cabi.dptr(du))
    return du


def max_threads():
    return int(lib().oracle_max_threads())


def _range(ic, rng):
    if rng is None:
        return cabi.Range(0, ic.n, 1)
    return rng


def trace(problem, ic, rng=None, nthreads=0, precision=0):
    rng = _range(ic, rng)
    out = cabi.EndpointArrays(rng.count)
    out.margin = np.zeros(rng.count)  # oracle-only: distance of the closest discrete decision from its threshold
    rc = lib().oracle_trace(C.byref(problem), C.byref(ic), C.byref(rng), nthreads, precision, C.byref(out.c),
                            cabi.dptr(out.margin))
    assert rc == 0
    return out


def trace_target(problem, ic, target, d_tol, rng=None, nthreads=0):
    """optimize_for_target's objective for every ray: (closest approach per ray, endpoints)."""
    rng = _range(ic, rng)
    out = cabi.EndpointArrays(rng.count)
    closest = np.zeros(rng.count)
    tgt = np.ascontiguousarray(target, np.float64)
    rc = lib().oracle_trace_target(C.byref(problem), C.byref(ic), C.byref(rng), nthreads, cabi.dptr(tgt), C.c_double(d_tol),
                                   C.byref(out.c), cabi.dptr(closest))
    assert rc == 0
    return closest, out


def render(problem, ic, pointfns, rng=None, nthreads=0, precision=0, plunging=None, endpoints=False):
    rng = _range(ic, rng)
    pfs = np.asarray(pointfns, np.int32)
    imgs = np.zeros((len(pfs), rng.count))
    ptrs = (cabi._dp * len(pfs))(*[cabi.dptr(imgs[k]) for k in range(len(pfs))])
    out = cabi.EndpointArrays(rng.count) if endpoints else None
    rc = lib().oracle_render(C.byref(problem), C.byref(ic), C.byref(rng), nthreads, precision, cabi.iptr(pfs), len(pfs),
                             C.byref(plunging) if plunging is not None else None, ptrs,
                             C.byref(out.c) if out is not None else None)
    assert rc == 0
    return (imgs, out) if endpoints else imgs


def lineprofile(problem, ic, emis, bins, opts, rng=None, nthreads=0, precision=0, plunging=None, endpoints=False):
    rng = _range(ic, rng)
    bins = np.ascontiguousarray(bins, np.float64)
    flux = np.zeros(len(bins))
    out = cabi.EndpointArrays(rng.count) if endpoints else None
    rc = lib().oracle_lineprofile(C.byref(problem), C.byref(ic), C.byref(rng), nthreads, precision, C.byref(emis),
                                  C.byref(plunging) if plunging is not None else None, cabi.dptr(bins), len(bins),
                                  C.byref(opts), cabi.dptr(flux), C.byref(out.c) if out is not None else None)
    assert rc == 0
    return (flux, out) if endpoints else flux


def trace_dual(problem, arrays, norm_mode=0, plunging=None, nthreads=0):
    """Forward-mode trace of `arrays` (a gradus_b200._cabi.DualArrays): fills its outputs and returns it."""
    rc = lib().oracle_trace_dual(C.byref(problem), C.byref(arrays.ic), norm_mode, C.byref(plunging) if plunging is not None else None,
                                 C.byref(arrays.out), nthreads)
    assert rc == 0
    return arrays


def set_cross_section(rho, h):
    """Install the cross-section table of GB200_GEOMETRY_THICK_TABLE (process-global; tests only)."""
    rho = np.ascontiguousarray(rho, np.float64)
    h = np.ascontiguousarray(h, np.float64)
    lib().oracle_set_cross_section(cabi.dptr(rho), cabi.dptr(h), len(rho))


def _mp(params):
    a = np.zeros(8)
    a[: len(params)] = params
    return a


def isco(kind, params, generic=False):
    out = C.c_double()
    mp = _mp(params)
    fn = lib().oracle_generic_isco if generic else lib().oracle_isco
    fn(kind, cabi.dptr(mp), C.byref(out))
    return out.value


def circular_energy(kind, params, r):
    out = C.c_double()
    mp = _mp(params)
    lib().oracle_circular_energy(kind, cabi.dptr(mp), C.c_double(r), C.byref(out))
    return out.value


def circular_fourvelocity(kind, params, r):
    v = np.zeros(4)
    mp = _mp(params)
    lib().oracle_circular_fourvelocity(kind, cabi.dptr(mp), C.c_double(r), cabi.dptr(v))
    return v


def metric(kind, params, r, th):
    g, dr, dth = np.zeros(5), np.zeros(5), np.zeros(5)
    mp = _mp(params)
    lib().oracle_metric(kind, cabi.dptr(mp), C.c_double(r), C.c_double(th), cabi.dptr(g), cabi.dptr(dr), cabi.dptr(dth))
    return g, dr, dth


def rhs(kind, params, u):
    u = np.ascontiguousarray(u, np.float64)
    du = np.zeros(8)
    mp = _mp(params)
    lib().oracle_rhs(kind, cabi.dptr(mp), cabi.dptr(u), cabi.dptr(du))
    return du


def lnrbasis(kind, params, r, th):
    b, f = np.zeros(16), np.zeros(16)
    mp = _mp(params)
    lib().oracle_lnrbasis(kind, cabi.dptr(mp), C.c_double(r), C.c_double(th), cabi.dptr(b), cabi.dptr(f))
    return b.reshape(4, 4), f.reshape(4, 4)


def initial_state(problem, alpha, beta):
    u = np.zeros(8)
    lib().oracle_initial_velocity(C.byref(problem), C.c_double(alpha), C.c_double(beta), cabi.dptr(u))
    return u


def trace_path(problem, u0, precision=0, cap=100000):
    """One ray, every accepted step: returns (t, dt, EEst, u[n,8])."""
    u0 = np.ascontiguousarray(u0, np.float64)
    t, dt, ee, u = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros((cap, 8))
    n = lib().oracle_trace_path(C.byref(problem), cabi.dptr(u0), precision, cap, cabi.dptr(t), cabi.dptr(dt), cabi.dptr(ee),
                                cabi.dptr(u.reshape(-1)))
    n = min(n, cap)
    return t[:n], dt[:n], ee[:n], u[:n]


def band_ratio(problem, ic, rng=None, nthreads=0):
    """Per ray: affine length of the first passage through {disc condition < 0} of the TRANSPARENT trajectory
    divided by the event sampler's spacing dt/7 (0: never inside).  0 < ratio < ~1.3 is the grazing band."""
    rng = _range(ic, rng)
    out = np.zeros(rng.count)
    rc = lib().oracle_band(C.byref(problem), C.byref(ic), C.byref(rng), nthreads, cabi.dptr(out))
    assert rc == 0
    return out


def trace_to(problem, u0, lam_end, nthreads=0):
    """Integrate each row of u0 (n x 8) exactly to lam_end[i] with no termination conditions; returns n x 8 states."""
    u0 = np.ascontiguousarray(u0, np.float64)
    lam_end = np.ascontiguousarray(lam_end, np.float64)
    out = np.zeros_like(u0)
    rc = lib().oracle_trace_to(C.byref(problem), C.c_int64(len(lam_end)), cabi.dptr(u0.reshape(-1)), cabi.dptr(lam_end), nthreads,
                               cabi.dptr(out.reshape(-1)))
    assert rc == 0
    return out
