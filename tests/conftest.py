import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def asset_root():
    from deepmimic_b200.assets import asset_root as ar
    # always test against the committed archive so the CPU and GPU boxes see identical inputs
    return ar(prefer_archive=True)


@pytest.fixture(scope="session")
def oracle_lib():
    from tests.oracle_binding import load_oracle
    return load_oracle()
