import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden" / "frontend_golden.npz"
PARAFORMER = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6)

# Stated tolerances: log-mel max-abs <= 1e-3 against the float32 reference on broadband input;
# after CMVN the bound scales with the largest Rescale entry.
LOGMEL_ATOL = 1e-3


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="session")
def cmvn(golden):
    return golden["cmvn"]
