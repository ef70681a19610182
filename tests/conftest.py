import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ora():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as r
    if os.path.isdir("/root/reference/ProbQA"):
        r.build()
    if not r.available():
        pytest.skip("oracle/_ref/libpqa_ref.so not built (needs /root/reference)")
    return r
