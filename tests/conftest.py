import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import hmoracle

    hmoracle.build()
    hmoracle.lib()
    return hmoracle

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_collection_modifyitems(config, items):
    # a wedged kernel must not eat the whole GPU lease: hard per-test limit (thread method = os._exit on expiry,
    # because a blocked cudaStreamSynchronize cannot be interrupted by a signal)
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(180, method="thread"))
