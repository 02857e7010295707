#!/usr/bin/env python
"""Wall clock of the target solver (optimize_for_target) and of continuum_time on the device."""
import math
import sys
import time

sys.path.insert(0, ".")
import gradus_b200 as gb  # noqa: E402
from gradus_b200 import api  # noqa: E402

m = gb.KerrMetric(1.0, 1.0)
x = [0.0, 1000.0, math.pi / 2, 0.0]
for target in [(10.0, 0.005, 0.0), (10.0, math.radians(40), -math.pi / 4)]:
    api.optimize_for_target(target, m, x)
    t0 = time.perf_counter()
    a, b, gp, acc = api.optimize_for_target(target, m, x)
    print(f"optimize_for_target{target}: alpha {a:.6f} beta {b:.6f} accuracy {acc:.2e} in {(time.perf_counter() - t0) * 1e3:.1f} ms")
m = gb.KerrMetric(1.0, 0.998)
t0 = time.perf_counter()
t = gb.reverberation.continuum_time(m, [0.0, 10_000.0, math.radians(45), 0.0], gb.corona.LampPostModel())
print(f"continuum_time: {t:.4f} in {(time.perf_counter() - t0) * 1e3:.1f} ms")
