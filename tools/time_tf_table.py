"""Wall-clock of a transfer-function table over an (a, theta) grid (BASELINE config 4, make_transfer_function_table):
all cells in lock step through gb200_render_batch against cell-by-cell (the reference's loop order)."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gradus_b200 as gb  # noqa: E402
from gradus_b200 import transfer_functions as tf  # noqa: E402


def main():
    na, nth, nr = (int(v) for v in (sys.argv[1:4] + ["4", "4", "40"])[:3])
    spins = np.linspace(0.0, 0.998, na)
    angles = np.linspace(10.0, 80.0, nth)
    cells = [(a, th) for a in spins for th in angles]
    metrics = [gb.KerrMetric(1.0, a) for a, _ in cells]
    observers = [[0.0, 10_000.0, math.radians(th), 0.0] for _, th in cells]
    d = gb.ThinDisc(0.0, float("inf"))
    radii_of = lambda m: 1.0 / np.linspace(1.0 / 500.0, 1.0 / (gb.isco(m) + 1e-2), nr)[::-1]  # Grids._inverse_grid
    ens = gb.EnsembleB200(devices=(0,))
    fast = tf.TransferFunctionSetup(warm_start=True, stall_exit=6)
    tf.transfer_function_table(metrics[:2], observers[:2], d, radii_of, ensemble=ens, setup=fast)  # warm-up
    t0 = time.perf_counter()
    table = tf.transfer_function_table(metrics, observers, d, radii_of, ensemble=ens, setup=fast)
    dt_table = time.perf_counter() - t0
    nctf = sum(len(r) for r in table)
    print(f"lock-step table: {len(cells)} cells x {nr} radii = {nctf} transfer functions in {dt_table:.2f} s "
          f"({dt_table / nctf * 1e3:.2f} ms per transfer function)", flush=True)
    ncmp = min(len(cells), 4)
    t0 = time.perf_counter()
    for m, x in zip(metrics[:ncmp], observers[:ncmp]):
        tf.cunningham_transfer_functions(m, x, d, radii_of(m), ensemble=ens, setup=fast)
    dt_cell = (time.perf_counter() - t0) / ncmp
    print(f"cell by cell: {dt_cell:.2f} s per cell -> {dt_cell * len(cells):.2f} s for the table "
          f"({dt_cell / nr * 1e3:.2f} ms per transfer function); lock step is {dt_cell * len(cells) / dt_table:.1f}x faster", flush=True)


if __name__ == "__main__":
    main()
