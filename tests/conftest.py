import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, loaded strictly: a GPU test must never pass on a fallback."""
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    from video_restore_b200 import _lib

    return _lib.load()
