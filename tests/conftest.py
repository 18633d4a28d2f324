import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device: the gpu-marked tests are skipped instead of failing in dm_create.
    (On the GPU box a missing device or a missing libdeepmimic_b200.so must still fail loudly: the skip only looks at the device.)"""
    if not any("gpu" in it.keywords for it in items):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (gpu-marked tests run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def asset_root():
    from deepmimic_b200.assets import asset_root as ar
    # always test against the committed archive so the CPU and GPU boxes see identical inputs
    return ar(prefer_archive=True)


@pytest.fixture(scope="session")
def oracle_lib():
    from tests.oracle_binding import load_oracle
    return load_oracle()
