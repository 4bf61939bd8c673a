import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def olib():
    """CPU oracle library (the checker)."""
    import oracle_engine
    return oracle_engine.load_oracle_library()


@pytest.fixture(scope="session")
def clib():
    """CUDA library (the product). Fails loudly if it has not been built.
    MCS_TEST_LOGIC_ONLY=1 substitutes the oracle so the TEST LOGIC of the gpu-marked tests can be debugged on a
    box without a GPU; such a run proves nothing about the kernel and is never what the driver runs."""
    if os.environ.get("MCS_TEST_LOGIC_ONLY") == "1":
        import oracle_engine
        return oracle_engine.load_oracle_library()
    import mcs_b200
    return mcs_b200.load_cuda_library()
