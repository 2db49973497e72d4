import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def state_dict():
    from chimeralm_b200.weights import make_state_dict, perturb_norms

    return perturb_norms(make_state_dict(0), seed=1)


@pytest.fixture(scope="session")
def engine(state_dict):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from chimeralm_b200.engine import Engine

    eng = Engine(state_dict, device=0, max_batch=4, max_tokens=1024)
    yield eng
    eng.close()
