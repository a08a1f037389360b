import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_block_cases():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "block_cases.pt"), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden_gnn_case():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "gnn_shipped_weights.pt"), map_location="cpu",
                      weights_only=False)
