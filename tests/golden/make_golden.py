#!/usr/bin/env python
"""Freeze small oracle outputs as regression fixtures (run from the repo root: python tests/golden/make_golden.py).

The reference is Julia and cannot run in this image, so these vectors come from the CPU oracle AFTER it was pinned
to the reference's own golden literals (tests/test_oracle_kat.py).  They guard the oracle (and through it every
parity test) against accidental change."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from gradus_b200 import _cabi as cabi  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    if "--c-only" in sys.argv:
        p, ic = common.c1(128, 128)[3].to_c()
        img = oracle.render(p, ic, [cabi.PF_REDSHIFT], nthreads=1)[0]
        img.astype("<f8").tofile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_128x128_redshift.f64"))
        print("wrote c1_128x128_redshift.f64:", img.size, "doubles,", int(np.sum(~np.isnan(img))), "hits")
        return
    out = {}
    for name, cfg in [("c1_24x24", common.c1(24, 24)[3]), ("c3_20x20", common.c3(20, 20)[4]), ("c5_20x20", common.c5(20, 20)[3])]:
        p, ic = cfg.to_c()
        imgs, ep = oracle.render(p, ic, [cabi.PF_REDSHIFT, cabi.PF_DISC_RADIUS], endpoints=True, nthreads=1)
        out[name + "_status"] = ep.status
        out[name + "_lambda"] = ep.lambda_max
        out[name + "_x"] = ep.x
        out[name + "_v"] = ep.v
        out[name + "_naccept"] = ep.naccept
        out[name + "_redshift"] = imgs[0]
        out[name + "_radius"] = imgs[1]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print("wrote", len(out), "arrays")
    # the C1 redshift image as raw little-endian doubles in ray order, for the plain-C boundary test (tests/c/cabi_render.c)
    p, ic = common.c1(128, 128)[3].to_c()
    img = oracle.render(p, ic, [cabi.PF_REDSHIFT], nthreads=1)[0]
    img.astype("<f8").tofile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_128x128_redshift.f64"))
    print("wrote c1_128x128_redshift.f64:", img.size, "doubles,", int(np.sum(~np.isnan(img))), "hits")


if __name__ == "__main__":
    main()
