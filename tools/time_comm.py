#!/usr/bin/env python
"""One process, all visible GPUs, through the library's own communicator (gb200_comm_*: what the Julia extension uses):
the C2 render and the C3 line profile against the single-GPU calls -- wall clock, equality of the results."""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402  (device count; loads its NCCL first, which the library then shares)

import gradus_b200 as gb  # noqa: E402
import common  # noqa: E402


def best_of(fn, n=3):
    out, best = None, 1e30
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return out, best


def main():
    ndev = torch.cuda.device_count()
    one = gb.EnsembleB200(devices=(0,))
    many = gb.EnsembleB200(devices=tuple(range(ndev)))
    m, x, d, _ = common.c1(8, 8)
    pfs = [gb.ConstPointFunctions.redshift(m, x) @ gb.ConstPointFunctions.filter_intersected(),
           gb.ConstPointFunctions.radius() @ gb.ConstPointFunctions.filter_intersected()]
    render = lambda ens: gb.rendergeodesics(m, x, d, 2000.0, pf=pfs, image_width=2048, image_height=2048, ensemble=ens)[2]
    render(one); render(many)
    a, t1 = best_of(lambda: render(one))
    b, tn = best_of(lambda: render(many))
    same = all(np.array_equal(u, v, equal_nan=True) for u, v in zip(a, b))
    print(f"C2 render 2048x2048, host images in ray order: 1 GPU {t1 * 1e3:.1f} ms, {ndev} GPUs in one process {tn * 1e3:.1f} ms "
          f"(speed-up {t1 / tn:.2f}, efficiency {t1 / tn / ndev:.3f}); images identical: {same}")
    m3 = gb.KerrMetric(1.0, 0.998)
    x3 = [0.0, 1000.0, math.radians(40.0), 0.0]
    plane = gb.PolarPlane(gb.GeometricGrid(), Nr=4096, Ntheta=4096, r_min=1.0, r_max=250.0)
    bins = np.linspace(0.1, 1.5, 180)
    lp = lambda ens: gb.lineprofile(bins, gb.PowerLawEmissivity(3.0), m3, x3, gb.ThinDisc(0.0, 400.0), gb.BinningMethod(), plane=plane,
                                    lambda_max=2000.0, ensemble=ens)[1]
    lp(one); lp(many)
    f1, t1 = best_of(lambda: lp(one))
    fn_, tn = best_of(lambda: lp(many))
    print(f"C3 line profile 4096x4096, histogram NCCL-reduced inside the library: 1 GPU {t1 * 1e3:.1f} ms, {ndev} GPUs {tn * 1e3:.1f} ms "
          f"(speed-up {t1 / tn:.2f}, efficiency {t1 / tn / ndev:.3f}); L1(flux_N - flux_1) = {np.abs(fn_ - f1).sum():.2e}, max bin {np.abs(fn_ - f1).max():.2e}")


if __name__ == "__main__":
    main()
