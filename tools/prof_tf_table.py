import cProfile, pstats, sys, math
sys.path.insert(0,'.')
import numpy as np
import gradus_b200 as gb
from gradus_b200 import transfer_functions as tf
spins = np.linspace(0.0, 0.998, 8); angles = np.linspace(10.0, 80.0, 8)
cells = [(a, th) for a in spins for th in angles]
metrics = [gb.KerrMetric(1.0, a) for a, _ in cells]
observers = [[0.0, 10_000.0, math.radians(th), 0.0] for _, th in cells]
d = gb.ThinDisc(0.0, float("inf"))
radii_of = lambda m: 1.0 / np.linspace(1.0 / 500.0, 1.0 / (gb.isco(m) + 1e-2), 50)[::-1]
ens = gb.EnsembleB200(devices=(0,))
tf.transfer_function_table(metrics[:2], observers[:2], d, radii_of, ensemble=ens)
pr = cProfile.Profile(); pr.enable()
tf.transfer_function_table(metrics, observers, d, radii_of, ensemble=ens)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
