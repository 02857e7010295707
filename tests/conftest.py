import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _no_cuda_device():
    """True only when this box has no CUDA device at all.  A box WITH a GPU on which libgradus_b200 fails to load or to
    initialise is not a reason to skip: there the gpu tests run and fail loudly."""
    try:
        import torch

        return not torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a usable device, so a plain `pytest tests` is meaningful
    everywhere; on the GPU box nothing is skipped and a missing library fails loudly in the tests themselves."""
    if not any("gpu" in it.keywords for it in items):
        return
    if not _no_cuda_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (the product path has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared libraries are built in-tree once per session (cross-compiles without a GPU)."""
    import __graft_entry__

    csrc = os.path.join(ROOT, "gradus.jl_b200", "csrc", "libgradus_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(csrc) and os.path.exists(orc)):
        __graft_entry__.build()
    yield


@pytest.fixture(scope="session")
def ensemble():
    import gradus_b200 as gb

    ens = gb.EnsembleB200(devices=(0,))
    yield ens
    ens.close()
