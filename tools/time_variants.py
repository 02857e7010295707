#!/usr/bin/env python
"""Time the C2 (Kerr) or C5 (Johannsen-Psaltis) render kernel for the library named by GB200_LIB (tuning aid).
usage: GB200_LIB=variants/lib....so tools/time_variants.py [size] [kerr|jp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gradus_b200 as gb
import common
ens = gb.EnsembleB200(devices=(0,))
w = h = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
which = sys.argv[2] if len(sys.argv) > 2 else "kerr"
m, x, d, cfg = (common.c5 if which == "jp" else common.c1)(w, h, ensemble=ens)
pfs = [gb.ConstPointFunctions.redshift() @ gb.ConstPointFunctions.filter_intersected(), gb.ConstPointFunctions.radius() @ gb.ConstPointFunctions.filter_intersected()]
best = 1e9
for rep in range(4):
    _, _, imgs = gb.rendergeodesics(m, x, d, 2000.0, pf=pfs, image_width=w, image_height=h, ensemble=ens)
    st = ens.stats(); best = min(best, st.kernel_ms)
print(f"{os.environ.get('GB200_LIB','default'):40s} {which} {w}x{h} kernel {best:.2f} ms -> {w*h/best/1e3:.2f} Mrays/s checksum {np.nansum(imgs[0]):.9f} hits {np.sum(~np.isnan(imgs[0]))}")
