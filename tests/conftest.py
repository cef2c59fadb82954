import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cv():
    import covid19abm_b200 as m
    if not os.path.exists(m.LIB_PATH):          # clean checkout: the .so files are git-ignored -> build them (nvcc cross-compiles without a GPU)
        import __graft_entry__
        __graft_entry__.build()
    return m


@pytest.fixture(scope="session")
def O():
    import oracle_api
    oracle_api.lib()
    return oracle_api
