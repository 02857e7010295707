"""C2 kernel time with the exact (double log/exp) and the Float32-accuracy (FastPower-style) step-size controller."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
import common
os.environ["GB200_NO_PIPELINE"] = "1"
ens = gb.EnsembleB200(devices=(0,))
m = gb.KerrMetric(1.0, 0.998); x = [0.0, 1000.0, np.deg2rad(60.0), 0.0]; d = gb.ThinDisc(0.0, 50.0)
pfs = [gb.ConstPointFunctions.redshift(m, x) @ gb.ConstPointFunctions.filter_intersected(), gb.ConstPointFunctions.radius() @ gb.ConstPointFunctions.filter_intersected()]
out = {}
for name, mode in (("exact", cabi.POW_EXACT), ("fast32", cabi.POW_FAST32)):
    best = 1e9
    for _ in range(3):
        _, _, imgs = gb.rendergeodesics(m, x, d, 2000.0, pf=pfs, image_width=2048, image_height=2048, ensemble=ens, pow_mode=mode)
        st = ens.stats(); best = min(best, st.kernel_ms)
    out[name] = imgs[0]
    print(f"{name:7s}: kernel {best:.2f} ms, attempts {st.steps_accepted + st.steps_rejected}, checksum {np.nansum(imgs[0]):.6f}, hits {np.sum(~np.isnan(imgs[0]))}")
both = ~np.isnan(out["exact"]) & ~np.isnan(out["fast32"])
print(f"max |g_exact - g_fast32| over common hits: {np.abs(out['exact'][both] - out['fast32'][both]).max():.3e}; NaN-mask differences: {np.sum(np.isnan(out['exact']) != np.isnan(out['fast32']))}")
