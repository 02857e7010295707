"""Multi-GPU plumbing: one process per GPU (`torch.distributed`), rays sharded with no data-path
collective; the only exchange on this path is the sum of line-profile histograms
(`ncclAllReduce(sum)` of `nbins` doubles over NVLink), SURVEY 8e.

Works with backend "nccl" on GPUs and "gloo" on CPU (used by the world_size-2 tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _cabi as cabi


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def interleaved_range(n_total, rank, world) -> cabi.Range:
    """Rays rank, rank+world, ...: balances the expensive photon-ring region across ranks."""
    count = (n_total - rank + world - 1) // world if n_total > rank else 0
    return cabi.Range(rank, count, world)


def strip_interleaved_range(ic, rank, world, columns=4) -> cabi.Range:
    """Whole strips of `columns` image columns (render grid) / theta-rows (polar plane), strip r, r+world, ... for rank r.
    Balances like ray interleaving but keeps neighbouring rays on the same GPU and in the same warp, which is worth
    up to 15 % of kernel time at 8 ranks (profiles/r01_tuning_log.md).  Falls back to ray interleaving when the image
    does not divide into strips."""
    if ic.kind in (cabi.IC_EXPLICIT, cabi.IC_IMPACT_PARAMETERS):
        return interleaved_range(ic.n, rank, world)  # ray lists have no image structure
    if ic.kind == cabi.IC_RENDER_GRID:
        h = ic.height
    elif ic.kind == cabi.IC_CARTESIAN_PLANE:
        h = 2 * (ic.height // 2) - 1  # fastest-varying extent of the mirrored grid (fill_params, gb200_api.cu)
    else:
        h = ic.width
    strip = columns * h
    if strip <= 0 or ic.n % strip != 0:
        return interleaved_range(ic.n, rank, world)
    nstrips = ic.n // strip
    mine = (nstrips - rank + world - 1) // world if nstrips > rank else 0
    return cabi.Range(rank * strip, mine * strip, world, strip)


def block_range(n_total, rank, world) -> cabi.Range:
    """Contiguous slab of rays (for images: a block of columns of the column-major (H, W) array)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return cabi.Range(first, base + (1 if rank < rem else 0), 1)


def allreduce_histogram(partial: torch.Tensor) -> torch.Tensor:
    """Sum raw line-profile partial sums over ranks in place (NCCL over NVLink on GPUs) and return
    `flux ./ sum(flux)` (src/line-profiles.jl:197), identical on every rank."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial / partial.sum()


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_floats(value: float, device=None):
    """List of `value` from every rank (on every rank)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [value]
    t = torch.tensor([value], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]
