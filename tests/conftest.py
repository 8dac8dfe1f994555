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
    from oracle import capi
    capi.build()
    return capi


@pytest.fixture(scope="session")
def dclient():
    import blaze_b200
    dc = blaze_b200.DriverClient("0", blaze_b200.DriverConfig.driver_client_cfg(blaze_b200.CardType.B200))
    yield dc
    dc.close()
