import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def berlin52():
    import oracle as O
    ids, x, y = O.read_tsplib_coords(os.path.join(GOLDEN, "berlin52.tsp"))
    return ids, x, y


@pytest.fixture(scope="session")
def att532():
    import oracle as O
    return O.read_tsplib_coords(os.path.join(GOLDEN, "att532.tsp"))
