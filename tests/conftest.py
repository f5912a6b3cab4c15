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
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def engine():
    import vadc_b200
    e = vadc_b200.Engine(max_streams=64)
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine_exact():
    """Engine with the bit-faithful STFT for every bin (SILERO_B200_STFT_EXACT)."""
    import vadc_b200
    e = vadc_b200.Engine(max_streams=8, stft_mode=vadc_b200.api.STFT_EXACT)
    yield e
    e.close()
