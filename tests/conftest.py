import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_events():
    from tests.common import load_golden

    return load_golden("events.npz")


@pytest.fixture(scope="session")
def golden_misc():
    from tests.common import load_golden

    return load_golden("misc.npz")


@pytest.fixture(scope="session")
def golden_traj():
    from tests.common import load_golden

    return load_golden("trajectories.npz")
