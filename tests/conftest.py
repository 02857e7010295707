import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


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
