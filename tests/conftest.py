import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def keep_mod():
    import keep_b200
    keep_b200.build()
    return keep_b200


@pytest.fixture(scope="session")
def lib(keep_mod):
    return keep_mod.keep_net.load_library()


@pytest.fixture(scope="session")
def state_dict():
    from oracle import weights
    return weights.make_state_dict(seed=0)


@pytest.fixture(scope="session")
def state_dict_asian():
    """Seeded synthetic weights with the 'Asian' config's key set (CFT at 32/64/128/256, modules/utils.py:58-73)."""
    from oracle import weights
    return weights.make_state_dict(seed=0, config="Asian")
