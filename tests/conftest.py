import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CPR_HMM = os.path.join(GOLDEN, "cpr_43_markers.hmm")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def cpr_oracle(oracle):
    return oracle.HmmFile(CPR_HMM)


@pytest.fixture(scope="session")
def engine():
    from checkm_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def cpr_models(engine):
    m = engine.load_models(CPR_HMM)
    yield m
    m.close()
