"""Wall-clock of Cunningham transfer functions through the device tracer: one radius, and the lock-step batch
over many radii (the loop `interpolated_transfer_branches` runs); optional oracle-tracer timing beside it."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gradus_b200 as gb  # noqa: E402
from gradus_b200 import transfer_functions as tf  # noqa: E402


def main():
    m = gb.KerrMetric(1.0, 0.998)
    x = [0.0, 1e5, math.radians(30), 0.0]
    d = gb.ThinDisc(0.0, float("inf"))
    chart = gb.chart_for_metric(m, 2e5, closest_approach=1.005)
    pr = gb.DeviceProber(m, x, d, chart=chart)
    tf.cunningham_transfer_function(m, x, d, 7.0, prober=pr)  # warm-up
    for mode, kw in (("exact", {}), ("fast", {"warm_start": True, "stall_exit": 6})):
        for nr in (1, 16, 150):
            radii = np.geomspace(gb.isco(m) + 1e-2, 1000.0, nr) if nr > 1 else [7.0]
            pr.launches = pr.rays = 0
            t0 = time.perf_counter()
            out = tf.cunningham_transfer_functions(m, x, d, radii, prober=pr, **kw)
            dt = time.perf_counter() - t0
            print(f"device ({mode}): {nr:4d} radii  {dt*1e3:9.1f} ms  {pr.launches} launches  {pr.rays} rays  "
                  f"{dt/nr*1e3:8.2f} ms/radius  {dt/pr.launches*1e3:6.2f} ms/launch  measure[0]={tf.measure_ctf(out[0]):.6f}", flush=True)
    if "--oracle" in sys.argv:
        from common import OracleProber
        po = OracleProber(m, x, d, chart=chart)
        radii = np.geomspace(gb.isco(m) + 1e-2, 1000.0, 16)
        t0 = time.perf_counter()
        tf.cunningham_transfer_functions(m, x, d, radii, prober=po)
        dt = time.perf_counter() - t0
        print(f"oracle tracer (all host threads): 16 radii {dt*1e3:.1f} ms  {dt/16*1e3:.2f} ms/radius  {po.rays} rays")


if __name__ == "__main__":
    main()
