import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        from gbnns_dim_red_b200 import capi

        return capi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_index_factory():
    from gbnns_dim_red_b200 import capi

    if not _has_gpu():
        pytest.fail("GPU test selected but no CUDA device / libgbdr.so: the CUDA path has no fallback")
    made = []

    def make(device=0):
        ix = capi.Index(device)
        made.append(ix)
        return ix

    yield make
    for ix in made:
        ix.close()
