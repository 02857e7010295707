#!/usr/bin/env python
"""Exploratory GPU run: parity statistics vs the oracle + raw timings. Output -> stdout."""
import ctypes as C
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gradus_b200 as gb  # noqa: E402
from gradus_b200 import _cabi as cabi  # noqa: E402
from gradus_b200.api import solve_tracing_problem  # noqa: E402
from oracle import oracle  # noqa: E402
import common  # noqa: E402


def compare(name, cfg, ens):
    p, ic = cfg.to_c()
    t = time.time()
    ref = oracle.trace(p, ic)
    t_or = time.time() - t
    gps = solve_tracing_problem(cfg)
    st = ens.stats()
    same = gps.status == ref.status
    cnt_same = (gps.naccept == ref.naccept) & (gps.nreject == ref.nreject)
    xerr = np.max(np.abs(gps.x - ref.x) / np.maximum(np.abs(ref.x), 1.0), axis=0)
    verr = np.max(np.abs(gps.v - ref.v) / np.maximum(np.abs(ref.v), 1e-3), axis=0)
    lerr = np.abs(gps.lambda_max - ref.lambda_max) / np.maximum(np.abs(ref.lambda_max), 1.0)
    ok = same
    print(f"[{name}] n={len(same)} status-mismatch={np.sum(~same)} stepcount-mismatch={np.sum(~cnt_same)} "
          f"status hist gpu={np.bincount(gps.status, minlength=4)} ref={np.bincount(ref.status, minlength=4)}")
    for code, nm in enumerate(["OutOfDomain", "Inner", "Intersected", "NoStatus"]):
        sel = ok & (ref.status == code)
        if sel.any():
            print(f"   {nm:12s} n={sel.sum():7d} max xerr={xerr[sel].max():.2e} verr={verr[sel].max():.2e} lam={lerr[sel].max():.2e} "
                  f"| same-stepcount subset: xerr={xerr[sel & cnt_same].max() if (sel & cnt_same).any() else 0:.2e}")
    print(f"   attempts/ray gpu={(gps.naccept + gps.nreject).mean():.2f} ref={(ref.naccept + ref.nreject).mean():.2f} "
          f"reject frac={gps.nreject.sum() / (gps.naccept.sum() + gps.nreject.sum()):.4f} flagged={st.flagged} "
          f"kernel_ms={st.kernel_ms:.3f} oracle_s={t_or:.2f} ({oracle.max_threads()} threads)")
    worst = np.argsort(-xerr)[:5]
    for i in worst:
        print(f"   worst ray {i}: status {gps.status[i]}/{ref.status[i]} acc {gps.naccept[i]}/{ref.naccept[i]} rej {gps.nreject[i]}/{ref.nreject[i]} "
              f"x_gpu={gps.x[:, i]} x_ref={ref.x[:, i]}")


def main():
    ens = gb.EnsembleB200(devices=(0,))
    peak = C.c_double()
    cabi.check(cabi.load().gb200_fp64_peak(ens.ctx(0), C.byref(peak)))
    print(f"measured FP64 DFMA peak: {peak.value:.2f} TFLOP/s")
    import __graft_entry__
    __graft_entry__.smoke()
    _, _, _, cfg = common.c1(128, 128, ensemble=ens)
    compare("C1 128x128 Kerr thin disc", cfg, ens)
    m = gb.KerrMetric(1.0, 0.0)
    x = [0.0, 100.0, math.radians(85), 0.0]
    cfg = common.render_config(m, x, None, 200.0, 20, 20, (-9.5, 9.5), (-9.5, 9.5), ensemble=ens)
    compare("KAT shadow 20x20", cfg, ens)
    _, _, img = gb.rendergeodesics(m, x, 200.0, image_width=20, image_height=20, αlims=(-9.5, 9.5), βlims=(-9.5, 9.5), ensemble=ens)
    print("   KAT shadow fingerprint gpu:", np.nansum(img), "reference literal 9009.452876609641")
    _, _, img = gb.rendergeodesics(m, x, gb.ThinDisc(0.0, 40.0), 200.0, image_width=20, image_height=20, αlims=(-9.5, 9.5), βlims=(-9.5, 9.5), ensemble=ens)
    print("   KAT thin-disc fingerprint gpu:", np.nansum(img), "reference literal 38412.08347901267")
    _, _, _, _, cfg = common.c3(96, 96, ensemble=ens)
    compare("C3 plane 96x96", cfg, ens)
    _, _, _, cfg = common.c5(96, 96, ensemble=ens)
    compare("C5 JP a=0.6 eps3=2 96x96", cfg, ens)
    # timings at scale (no oracle)
    for (w, h) in [(512, 512), (2048, 2048)]:
        m, x, d, cfg = common.c1(w, h, ensemble=ens)
        pfs = [gb.ConstPointFunctions.redshift() @ gb.ConstPointFunctions.filter_intersected(),
               gb.ConstPointFunctions.radius() @ gb.ConstPointFunctions.filter_intersected()]
        for rep in range(3):
            t = time.time()
            _, _, imgs = gb.rendergeodesics(m, x, d, 2000.0, pf=pfs, image_width=w, image_height=h, ensemble=ens)
            wall = time.time() - t
            st = ens.stats()
            att = st.steps_accepted + st.steps_rejected
            print(f"[render {w}x{h}] kernel {st.kernel_ms:.2f} ms total {st.total_ms:.2f} ms wall {wall*1e3:.1f} ms -> {w*h/st.kernel_ms/1e3:.3f} Mrays/s "
                  f"attempts/ray {att/(w*h):.1f} hits {np.sum(~np.isnan(imgs[0]))}")


if __name__ == "__main__":
    main()
