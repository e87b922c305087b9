"""Test configuration: registers the ``gpu`` marker and common fixtures.

``-m "not gpu"`` covers the oracle against the golden vectors, the host logic
and the C-ABI symbol table; ``-m gpu`` tests are the parity tests proper and
call the CUDA kernels through the C-ABI.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def assembly_golden():
    return np.load(GOLDEN / "assembly_f64.npz")


@pytest.fixture(scope="session")
def example_inputs():
    out = {}
    for p in sorted((GOLDEN / "inputs").glob("*.json")):
        with open(p) as f:
            out[p.stem] = json.load(f)
    return out
