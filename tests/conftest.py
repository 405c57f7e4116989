import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tiny_cfg():
    from infinisst_b200 import tiny_config
    return tiny_config()


@pytest.fixture(scope="session")
def tiny_sd(tiny_cfg):
    from infinisst_b200.synthetic import make_state_dict
    return make_state_dict(tiny_cfg, seed=0)
