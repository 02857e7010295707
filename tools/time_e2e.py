"""e2e wall clock of gb200_render (host buffers, pinned and pageable) with and without the chunk pipeline."""
import ctypes as C, os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import gradus_b200 as gb
from gradus_b200 import _cabi as cabi
import common
ens = gb.EnsembleB200(devices=(0,))
lib = cabi.load(); ctx = ens.ctx(0)
m, x, d, cfg = common.c1(2048, 2048, ensemble=ens)
p, ic = cfg.to_c()
pfs = np.array([cabi.PF_REDSHIFT, cabi.PF_DISC_RADIUS], np.int32)
rng = cabi.Range(0, ic.n, 1)
for name, mk in (("pinned", lambda: torch.empty(ic.n, dtype=torch.float64).pin_memory().numpy()), ("pageable", lambda: np.empty(ic.n))):
    imgs = [mk(), mk()]
    ptrs = (cabi._dp * 2)(cabi.dptr(imgs[0]), cabi.dptr(imgs[1]))
    for _ in range(2):
        cabi.check(lib.gb200_render(ctx, C.byref(p), C.byref(ic), C.byref(rng), cabi.iptr(pfs), 2, None, ptrs), ctx)
    t0 = time.perf_counter()
    for _ in range(5):
        cabi.check(lib.gb200_render(ctx, C.byref(p), C.byref(ic), C.byref(rng), cabi.iptr(pfs), 2, None, ptrs), ctx)
    dt = (time.perf_counter() - t0) / 5
    st = ens.stats()
    print(f"{os.environ.get('GB200_NO_PIPELINE','pipeline'):>10s} {name:9s}: e2e {dt*1e3:.2f} ms, kernel span {st.kernel_ms:.2f} ms, checksum {np.nansum(imgs[0]):.6f} launches {st.launches}")
