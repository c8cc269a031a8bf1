import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_api():
    from oracle import api
    api.build(ref=True)
    return api


@pytest.fixture(scope="session")
def rrx():
    import roaringregex_b200
    roaringregex_b200.lib()
    return roaringregex_b200


def cpu_corpus(cfg, offset, length):
    from oracle import api
    return api.corpus(cfg, offset, length)
