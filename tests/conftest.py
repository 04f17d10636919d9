import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def qp():
    """The product package with the CUDA library built and loaded (fails loudly if it is missing)."""
    from qpsk_modulator_demodulator_b200 import build as _b
    _b.build()
    import qpsk_modulator_demodulator_b200 as q
    q._native.lib()
    return q


@pytest.fixture(scope="session")
def gpu(qp):
    if qp.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need a B200 (there is no CPU fallback)")
    qp.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    return qp
