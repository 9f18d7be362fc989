"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu`` and call the
product through its C-ABI; everything else runs on CPU."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))


@pytest.fixture(scope="session")
def cases():
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        return json.load(f)


@pytest.fixture
def random_signal():
    # same generator as the reference's tests/conftest.py:12-16
    return np.random.default_rng(42).standard_normal(22050).astype(np.float32)


@pytest.fixture
def batch_signals():
    return np.random.default_rng(42).standard_normal((4, 22050)).astype(np.float32)


def chirp_noise(n, sr=22050, seed=42):
    """benchmarks/utils.py:92-115 style synthetic clip."""
    t = np.arange(n) / sr
    x = np.sin(2 * np.pi * (100 + 1000 * t) * t)
    return (x + 0.1 * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)
