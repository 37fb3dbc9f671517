import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # the oracle evaluates prior draws the reference itself turns into NaN (log10 of 0, 0/0)
    config.addinivalue_line("filterwarnings", "ignore::RuntimeWarning")
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0
