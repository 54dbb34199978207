import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ckpt_ra1e5():
    from rbc_gym_b200.h5lite import load_checkpoint_2d
    return load_checkpoint_2d(ROOT / "data/checkpoints/train/ckpt_ra100000.h5")


@pytest.fixture(scope="session")
def ckpt_ra1e4():
    from rbc_gym_b200.h5lite import load_checkpoint_2d
    return load_checkpoint_2d(ROOT / "data/checkpoints/test/ckpt_ra10000.h5")
