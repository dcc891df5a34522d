import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Parity numbers measured by the GPU tier: every test appends lines through `parity_report(...)`; they go to
# gpurun_out/parity_report.txt AND are printed in the terminal summary, so the run's log carries them (the driver keeps the
# log, not gpurun_out/).
_PARITY_LINES = []


def parity_report(tag, **kw):
    line = tag + " " + " ".join("%s=%s" % (k, v) for k, v in kw.items())
    _PARITY_LINES.append(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "a") as f:
        f.write(line + "\n")


def pytest_terminal_summary(terminalreporter):
    if _PARITY_LINES:
        terminalreporter.section("parity report (engine vs oracle / reference fixtures)")
        for line in _PARITY_LINES:
            terminalreporter.write_line(line)


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
