"""world_size-2 gloo test of the N>1 host logic: ray sharding + histogram all-reduce + normalisation.
The per-rank compute is done by the oracle here (CPU); on GPUs the same code path feeds NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import gradus_b200 as gb
    from gradus_b200 import _cabi as cabi
    from gradus_b200 import distributed as gd
    from oracle import oracle
    import common

    r, w, _ = gd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    m, x, d, plane, cfg = common.c3(24, 24)
    p, ic = cfg.to_c()
    bins = np.linspace(0.1, 1.5, 40)
    emis = cabi.Emissivity(cabi.EMISSIVITY_POWERLAW, 0, 3.0, None, None)
    opts = cabi.LineProfileOpts(gb.isco(m), 50.0, 0, 0)
    rng = gd.interleaved_range(ic.n, rank, world) if mode == "interleaved" else gd.block_range(ic.n, rank, world)
    partial = torch.from_numpy(oracle.lineprofile(p, ic, emis, bins, opts, rng=rng, nthreads=2))
    flux = gd.allreduce_histogram(partial)
    total = gd.sum_over_ranks(float(rng.count))
    mx = gd.max_over_ranks(float(rank))
    gd.barrier()
    if rank == 0:
        q.put((flux.numpy(), total, mx))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["interleaved", "block"])
def test_two_rank_line_profile_equals_single_rank(mode):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gradus_b200 as gb
    from gradus_b200 import _cabi as cabi
    from oracle import oracle
    import common

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + (0 if mode == "interleaved" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    [p.start() for p in procs]
    flux2, total, mx = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    m, x, d, plane, cfg = common.c3(24, 24)
    p, ic = cfg.to_c()
    bins = np.linspace(0.1, 1.5, 40)
    emis = cabi.Emissivity(cabi.EMISSIVITY_POWERLAW, 0, 3.0, None, None)
    opts = cabi.LineProfileOpts(gb.isco(m), 50.0, 1, 0)
    flux1 = oracle.lineprofile(p, ic, emis, bins, opts)
    assert total == ic.n and mx == 1.0
    assert flux2.sum() == pytest.approx(1.0, abs=1e-12)
    assert np.abs(flux2 - flux1).sum() < 1e-12  # invariant to the rank count up to summation order


def test_range_helpers_partition_the_rays():
    from gradus_b200 import distributed as gd

    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a = gd.interleaved_range(n, r, world)
                seen += [a.first + k * a.stride for k in range(a.count)]
            assert sorted(seen) == list(range(n))
            seen = []
            for r in range(world):
                b = gd.block_range(n, r, world)
                seen += list(range(b.first, b.first + b.count))
            assert seen == list(range(n))
    # strips of 4 columns of a 16 x 24 image (H = 16): every ray exactly once, whole strips per rank
    import gradus_b200 as gb
    from gradus_b200.api import RenderGrid, tracing_configuration

    cfg = tracing_configuration(gb.KerrMetric(), [0.0, 100.0, 1.0, 0.0], RenderGrid(24, 16, (-5, 5), (-5, 5)), 200.0, trajectories=384)
    _, ic = cfg.to_c()
    for world in (1, 2, 3, 4):
        seen = np.concatenate([gd.strip_interleaved_range(ic, r, world).indices() for r in range(world)])
        assert sorted(seen.tolist()) == list(range(384))
        assert all(gd.strip_interleaved_range(ic, r, world).block == 64 for r in range(world))
    # ray lists have no strips (ic.width = 0 must not divide), Cartesian planes use their mirrored column height
    from gradus_b200 import api

    al = np.linspace(-3, 3, 37)
    _, ic = tracing_configuration(gb.KerrMetric(), [0.0, 100.0, 1.0, 0.0], api.ImpactParameters(al, al), gb.DatumPlane(0.0), 200.0).to_c()
    for world in (1, 2, 5):
        seen = np.concatenate([gd.strip_interleaved_range(ic, r, world).indices() for r in range(world)])
        assert sorted(seen.tolist()) == list(range(37))
    plane = gb.CartesianPlane(gb.LinearGrid(), Nx=18, Ny=34, x_max=20.0, y_max=20.0)  # 17 columns of 33 rays
    _, ic = tracing_configuration(gb.KerrMetric(), [0.0, 100.0, 1.0, 0.0], plane, (0.0, 200.0)).to_c()
    assert ic.n == 33 * 17
    for world in (1, 2, 3):
        rngs = [gd.strip_interleaved_range(ic, r, world) for r in range(world)]
        seen = np.concatenate([r_.indices() for r_ in rngs])
        assert sorted(seen.tolist()) == list(range(ic.n))
