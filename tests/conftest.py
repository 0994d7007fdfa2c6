import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port_oracle():
    from oracles import PortOracle, build_oracles
    build_oracles()
    return PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracles import RefOracle, build_oracles
    build_oracles()
    if not RefOracle.available():
        pytest.skip("compiled reference oracle not available")
    return RefOracle()
