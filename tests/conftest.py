import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def demo_inputs():
    import numpy as np
    d = np.load(os.path.join(GOLDEN, "demo_inputs.npz"))
    return torch.from_numpy(d["ecg"]), torch.from_numpy(d["demo"])


@pytest.fixture(scope="session")
def expected_probs():
    import json
    with open(os.path.join(GOLDEN, "expected_probs.json")) as f:
        return json.load(f)


def load_ckpt(name):
    return torch.load(os.path.join(GOLDEN, "ckpts", name), map_location="cpu")["model_state"]
