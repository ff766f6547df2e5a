import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _strict_fp32_reference():
    """The parity tests compare the bf16 sm_100a hot path with fp32 oracle code and, where `encoder_autocast = False`, run the
    cuDNN encoder in fp32 to isolate the hot path: TF32 must be off for both, independent of which test module runs first."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture(scope="session")
def golden_stages():
    import numpy as np
    return np.load(GOLDEN / "stages.npz")


@pytest.fixture(scope="session")
def state_shapes():
    import json
    return {m: {k: tuple(v) for k, v in json.loads((GOLDEN / f"state_keys_{m}.json").read_text()).items()}
            for m in ("psmnet", "stereodpnet")}
