import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the C-ABI library and the checker are build artefacts; (re)build them when stale so `pytest` works from a clean tree
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def example_vxa():
    with open(os.path.join(ROOT, "tests", "data", "example.vxa")) as f:
        return f.read()
