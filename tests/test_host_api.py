"""Host-side logic and the C-ABI boundary, without a GPU: the library loads and exports every symbol
include/gradus_b200.h declares, struct layouts match, configuration errors surface like the reference's."""
import ctypes as C
import math
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
from gradus_b200.api import RenderGrid, tracing_configuration

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gradus_b200.h")


def test_library_exports_every_declared_symbol():
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", text))
    assert declared == set(cabi.EXPORTED_SYMBOLS)
    lib = cabi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gb200_version() == 100
    # and nothing from the oracle is linked into the product library
    nm = subprocess.run(["nm", "-D", cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in nm


def test_ctypes_struct_layout_matches_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "gradus_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(gb200_problem), sizeof(gb200_ic), sizeof(gb200_range), sizeof(gb200_endpoints),
         sizeof(gb200_emissivity), sizeof(gb200_plunging_table), sizeof(gb200_lineprofile_opts), sizeof(gb200_stats),
         offsetof(gb200_problem, maxiters));
  printf("%zu %zu %zu %zu\n", offsetof(gb200_ic, x), offsetof(gb200_ic, n), offsetof(gb200_endpoints, naccept), offsetof(gb200_problem, gtol));
  printf("%zu %zu %zu %zu %zu\n", sizeof(gb200_dual_ic), sizeof(gb200_dual_out), offsetof(gb200_dual_ic, alpha), offsetof(gb200_dual_out, g),
         offsetof(gb200_dual_out, flags));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        cfile = os.path.join(d, "l.c")
        open(cfile, "w").write(src)
        exe = os.path.join(d, "l")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    sizes = [int(v) for v in out]
    want = [C.sizeof(cabi.Problem), C.sizeof(cabi.IC), C.sizeof(cabi.Range), C.sizeof(cabi.Endpoints), C.sizeof(cabi.Emissivity),
            C.sizeof(cabi.PlungingTable), C.sizeof(cabi.LineProfileOpts), C.sizeof(cabi.Stats), cabi.Problem.maxiters.offset,
            cabi.IC.x.offset, cabi.IC.n.offset, cabi.Endpoints.naccept.offset, cabi.Problem.gtol.offset,
            C.sizeof(cabi.DualIC), C.sizeof(cabi.DualOut), cabi.DualIC.alpha.offset, cabi.DualOut.g.offset, cabi.DualOut.flags.offset]
    assert sizes == want


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = cabi.load().gb200_init(0, C.byref(h))
    assert rc == cabi.ERR_NO_DEVICE
    assert b"no CPU fallback" in cabi.load().gb200_last_error(None)
    with pytest.raises(gb.GradusB200Error):
        gb.rendergeodesics(gb.KerrMetric(1.0, 0.5), [0, 1000.0, 1.0, 0], 2000.0, image_width=4, image_height=4)


def test_special_radii_host_functions():
    assert gb.isco(gb.KerrMetric(1.0, 0.998)) == pytest.approx(1.2369706551751847, abs=1e-12)  # test/smoke-tests/special-radii.jl:27
    assert gb.isco(gb.KerrMetric(1.0, -0.998)) == pytest.approx(8.99437445480357, abs=1e-12)
    assert gb.isco(gb.KerrMetric(1.0, 0.0)) == pytest.approx(6.0, abs=1e-12)
    # generic root find (JP with eps3 = 0 is Kerr) reproduces the analytic value
    assert gb.isco(gb.JohannsenPsaltisMetric(1.0, 0.998, 0.0)) == pytest.approx(1.2369706551751847, abs=1e-9)
    from oracle import oracle

    assert gb.isco(gb.JohannsenPsaltisMetric(1.0, 0.6, 2.0)) == pytest.approx(oracle.isco(cabi.METRIC_JP, (1.0, 0.6, 2.0)), abs=1e-10)
    with pytest.raises(gb.GradusB200Error):  # the reference errors too: "does not have an ISCO solution" (special-radii.jl:16-19)
        gb.isco(gb.JohannsenPsaltisMetric(1.0, 0.8831, 0.4))
    ss = gb.ShakuraSunyaev(gb.KerrMetric(1.0, 0.0))
    assert ss.inner_radius == pytest.approx(6.0) and ss.inv_eta == pytest.approx(1.0 / (1.0 - math.sqrt(8.0 / 9.0)), rel=1e-12)
    assert gb.inner_radius(gb.KerrMetric(1.0, 0.998)) == pytest.approx(1.0 + math.sqrt(1 - 0.998**2))


def test_configuration_defaults_follow_the_reference():
    m = gb.KerrMetric(1.0, 0.998)
    cfg = tracing_configuration(m, [0.0, 1000.0, 1.0, 0.0], RenderGrid(8, 4, (-60, 60), (-40, 40)), gb.ThinDisc(0.0, 50.0), 2000.0, trajectories=32)
    p, ic = cfg.to_c()
    assert (p.abstol, p.reltol, p.gtol) == (1e-9, 1e-9, 1e-2)  # configuration.jl:1,101-102; bootstrap.jl:8
    assert p.chart_inner == pytest.approx(1.01 * gb.inner_radius(m)) and p.chart_outer == 12000.0  # charts.jl:51-58
    assert (p.lambda_min, p.lambda_max) == (0.0, 2000.0)
    assert p.geometry_kind == cabi.GEOMETRY_THIN_DISC and list(p.geometry_params)[:2] == [0.0, 50.0]
    assert ic.kind == cabi.IC_RENDER_GRID and ic.n == 32 and (ic.width, ic.height) == (8, 4)
    plane = gb.PolarPlane(gb.GeometricGrid(), Nr=5, Ntheta=7)
    cfg = tracing_configuration(m, [0.0, 1000.0, 1.0, 0.0], plane, (0.0, 2000.0), callback=gb.domain_upper_hemisphere())
    p, ic = cfg.to_c()
    assert ic.kind == cabi.IC_POLAR_PLANE and ic.n == 35 and p.callback_kind == cabi.CALLBACK_UPPER_HEMISPHERE and p.callback_delta == 1e-4


def test_errors_mirror_the_reference():
    m = gb.KerrMetric(1.0, 0.5)
    x = [0.0, 1000.0, 1.0, 0.0]
    with pytest.raises(TypeError):  # kwargshandle = KeywordArgError (tracing.jl:106,146,215)
        tracing_configuration(m, x, RenderGrid(4, 4, (-1, 1), (-1, 1)), 2000.0, trajectories=16, dense=True)
    with pytest.raises(ValueError):  # tracing.jl:159-161
        tracing_configuration(m, x, RenderGrid(4, 4, (-1, 1), (-1, 1)), 2000.0, trajectories=16, save_on=True)
    with pytest.raises(ValueError):  # no CPU ensemble in this build
        tracing_configuration(m, x, RenderGrid(4, 4, (-1, 1), (-1, 1)), 2000.0, trajectories=16, ensemble=gb.EnsembleEndpointThreads())
    with pytest.raises(AssertionError):  # rendering.jl:148-149
        tracing_configuration(m, x, RenderGrid(4, 4, (1, -1), (-1, 1)), 2000.0, trajectories=16).to_c()
    with pytest.raises(ValueError):  # configuration.jl:47-49
        tracing_configuration(m, x, lambda i: [0, -1, 0, 0], 2000.0).to_c()
    with pytest.raises(ValueError):
        tracing_configuration(object(), x, RenderGrid(4, 4, (-1, 1), (-1, 1)), 2000.0, trajectories=16).to_c()
    with pytest.raises(ValueError):
        tracing_configuration(m, x, RenderGrid(4, 4, (-1, 1), (-1, 1)), object(), 2000.0, trajectories=16).to_c()
    lib = cabi.load()
    p, ic = tracing_configuration(m, x, RenderGrid(4, 4, (-1, 1), (-1, 1)), 2000.0, trajectories=16).to_c()
    p.metric_kind = 7
    assert lib.gb200_validate(C.byref(p), C.byref(ic)) == cabi.ERR_UNSUPPORTED
    p.metric_kind = 0
    p.reltol = 0.0
    assert lib.gb200_validate(C.byref(p), C.byref(ic)) == cabi.ERR_INVALID_ARGUMENT
    p.reltol = 1e-9
    p.metric_params[1] = 1.5  # |a| > M
    assert lib.gb200_validate(C.byref(p), C.byref(ic)) == cabi.ERR_INVALID_ARGUMENT


def test_explicit_initial_conditions_are_flattened_to_soa():
    m = gb.KerrMetric(1.0, 0.9)
    vs = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, -1.0, 0.01, 0.0], [0.0, -1.0, 0.0, 0.002]])
    cfg = tracing_configuration(m, [0.0, 100.0, 1.2, 0.0], vs, gb.ThinDisc(0.0, 20.0), 300.0)
    p, ic = cfg.to_c()
    assert ic.kind == cabi.IC_EXPLICIT and ic.n == 3
    assert np.ctypeslib.as_array(ic.v[2], shape=(3,)).tolist() == [0.0, 0.01, 0.0]
    cfg = tracing_configuration(m, [0.0, 100.0, 1.2, 0.0], lambda i: [0.0, -1.0, 1e-3 * i, 0.0], 300.0, trajectories=4)
    p, ic = cfg.to_c()
    assert ic.n == 4 and np.ctypeslib.as_array(ic.v[2], shape=(4,)).tolist() == pytest.approx([1e-3, 2e-3, 3e-3, 4e-3])


def test_point_function_composition():
    cpf = gb.ConstPointFunctions
    assert (cpf.redshift() @ cpf.filter_intersected()).kind() == cabi.PF_REDSHIFT
    assert cpf.shadow().kind() == cabi.PF_SHADOW
    with pytest.raises(ValueError):
        (cpf.redshift() @ cpf.filter_early_term()).kind()
    a, b = gb.impact_axes(5, 3, (-60, 60), (-40, 40))
    assert a.tolist() == [-60, -30, 0, 30, 60] and b.tolist() == [-40, 0, 40]


def test_apply_over_an_endpoint_cache_reproduces_the_prerender_literal():
    """test/smoke-tests/prerendergeodesics.jl:3-21,33-43: `apply(pf, cache)` with a user point function ∘ filter over the
    cached endpoints of a 20 x 20 render; the cache is filled by the oracle here (the device fills it in the GPU test)."""
    import math

    from oracle import oracle

    import common

    m = gb.KerrMetric(1.0, 0.0)
    cfg = common.render_config(m, [0.0, 100.0, math.radians(85), 0.0], None, 200.0, 20, 20, (-9.5, 9.5), (-9.5, 9.5))
    p, ic = cfg.to_c()
    cache = gb.EndpointCache(m, 200.0, 20, 20, gb.api.GeodesicPoints(oracle.trace(p, ic), 0.0))
    pf = gb.HostPointFunction(lambda m_, gp, lam: gp.lambda_max) @ gb.HostPointFunction(lambda m_, gp, lam: gp.lambda_max < lam, np.nan)
    img = gb.apply(pf, cache)
    assert img.shape == (20, 20)
    assert np.nansum(img) == pytest.approx(9009.452876609641, rel=1e-6)
    assert np.array_equal(img, gb.apply(gb.ConstPointFunctions.shadow(), cache), equal_nan=True)
    with pytest.raises(ValueError):
        gb.apply(gb.ConstPointFunctions.redshift(m, None) @ gb.ConstPointFunctions.filter_intersected(), cache)


def test_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference: one JSON line on stdout (everything else goes to stderr), with the contract's keys."""
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-stride", "1024"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in j, key
    assert j["impl"] == "reference" and j["unit"] == "rays/s" and j["value"] > 0 and j["cpu_baseline"]["kind"] == "port"
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in j["config"]


def _build_c_program(tmpdir):
    exe = os.path.join(tmpdir, "cabi_render")
    csrc = os.path.join(ROOT, "gradus.jl_b200", "csrc")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "cabi_render.c"),
                           "-L", csrc, "-lgradus_b200", f"-Wl,-rpath,{csrc}", "-lm", "-o", exe])
    return exe


def test_plain_c_program_links_against_the_boundary():
    """tests/c/cabi_render.c: only the header and the shared library, no Python in between.  Without a device it must stop
    at gb200_init with GB200_ERR_NO_DEVICE (exit 77), after validating its problem description on the host."""
    import torch

    with tempfile.TemporaryDirectory() as d:
        exe = _build_c_program(d)
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "c1_128x128_redshift.f64")], capture_output=True, text=True, cwd=ROOT)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_plain_c_program_renders_config_1_with_parity():
    """The same program on the GPU: gb200_render driven from C reproduces the frozen oracle image of BASELINE config 1."""
    with tempfile.TemporaryDirectory() as d:
        exe = _build_c_program(d)
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "c1_128x128_redshift.f64")], capture_output=True, text=True, cwd=ROOT)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NaN-mask mismatches" in r.stdout
