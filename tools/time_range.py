#!/usr/bin/env python
"""Kernel time of the C2 render for a (first, count, stride) sub-range of a 2048 x (2048*N) image (coherence experiment)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
ens = gb.EnsembleB200(devices=(0,))
lib = cabi.load(); ctx = ens.ctx(0)
for world in (1, 2, 4, 8):
    cfg, w, h = bench.build_workload(world, ensemble=ens)
    p, ic = cfg.to_c()
    from gradus_b200 import distributed as gd
    rng = gd.strip_interleaved_range(ic, 0, world) if os.environ.get("STRIPS") else cabi.Range(0, ic.n // world, world)
    pfs = np.array([cabi.PF_REDSHIFT, cabi.PF_DISC_RADIUS], np.int32)
    imgs = np.zeros((2, rng.count)); ptrs = (cabi._dp * 2)(cabi.dptr(imgs[0]), cabi.dptr(imgs[1]))
    best = 1e9
    for _ in range(3):
        cabi.check(lib.gb200_render(ctx, C.byref(p), C.byref(ic), C.byref(rng), cabi.iptr(pfs), 2, None, ptrs), ctx)
        best = min(best, ens.stats().kernel_ms)
    print(f"world {world}: rank-0 shard (stride {world}) of {w}x{h}: kernel {best:.2f} ms")
