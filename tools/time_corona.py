"""Wall-clock of lamp-post emissivity profiles: one model, and a 20 x 20 (spin, height) grid fused into one batch
(SURVEY config C4: 1000 rays per model)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gradus_b200 as gb  # noqa: E402
from gradus_b200 import corona  # noqa: E402


def main():
    d = gb.ThinDisc(0.0, 1000.0)
    ens = gb.EnsembleB200([0])
    one = [(gb.KerrMetric(1.0, 0.998), d, corona.LampPostModel(h=10.0))]
    corona.emissivity_profiles(one, n_samples=1000, ensemble=ens)
    t0 = time.perf_counter()
    for _ in range(10):
        corona.emissivity_profiles(one, n_samples=1000, ensemble=ens)
    print(f"1 model x 1000 rays: {(time.perf_counter()-t0)/10*1e3:.2f} ms per profile")
    grid = [(gb.KerrMetric(1.0, a), d, corona.LampPostModel(h=h)) for a in np.linspace(0.0, 0.998, 20) for h in np.geomspace(2.5, 50.0, 20)]
    corona.emissivity_profiles(grid[:8], n_samples=1000, ensemble=ens)
    t0 = time.perf_counter()
    profs = corona.emissivity_profiles(grid, n_samples=1000, ensemble=ens)
    dt = time.perf_counter() - t0
    print(f"400 models x 1000 rays fused: {dt*1e3:.1f} ms total, {dt/400*1e3:.3f} ms per profile, "
          f"{400*1000/dt:.3e} rays/s; eps(r=10) range {min(p.emissivity_at(10.0) for p in profs):.3e}..{max(p.emissivity_at(10.0) for p in profs):.3e}")
    # the device share of that: the 400 ensembles alone, configurations prebuilt
    from gradus_b200 import api
    prepared = [corona._point_source_job(m, dd, model, 0.01, 179.99, 1000, 10_000.0, api.domain_upper_hemisphere(), {}) for (m, dd, model) in grid]
    configs = [api.tracing_configuration(m, xs.T, vs.T, dd, lam, ensemble=ens, **kw) for (_, _, (m, xs, vs, dd, lam, kw)) in prepared]
    api.tracegeodesics_batch(configs)
    t0 = time.perf_counter()
    api.tracegeodesics_batch(configs)
    dt = time.perf_counter() - t0
    print(f"gb200_trace_batch of 400 x 1000 rays (host buffers in/out): {dt*1e3:.1f} ms, {400*1000/dt:.3e} rays/s")


if __name__ == "__main__":
    main()
