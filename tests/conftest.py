import os
import sys

import pytest

os.environ.setdefault("SPP_RL_SYNTHETIC_ENVS", "1")      # the suite opts in to the shape-only MuJoCo stand-ins (spp_rl_b200/envs.py)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    import spp_rl_b200

    return spp_rl_b200.load_library()
