import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden_pure():
    return {p: load_golden(f"pure_ref{p}.npz") for p in (32, 64)}


@pytest.fixture(scope="session")
def golden_states():
    out = {}
    for p in (32, 64):
        flat = load_golden(f"states_ref{p}.npz")
        cfgs = {}
        for k, v in flat.items():
            cfg, field = k.split(".", 1)
            cfgs.setdefault(cfg, {})[field] = v
        out[p] = cfgs
    return out


def mask_to_idx(mask):
    mask = int(mask)
    return [b for b in range(64) if mask >> b & 1]
