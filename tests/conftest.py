import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the CUDA library; GPU tests fail loudly if it cannot be loaded."""
    from sus_net_b200 import build as B

    B.build()
    import sus_net_b200

    return sus_net_b200.lib()
