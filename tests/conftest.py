import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_cases():
    return np.load(os.path.join(GOLDEN, "golden_cases.npz"))


@pytest.fixture(scope="session")
def watersurface_u8():
    return np.asfortranarray(np.load(os.path.join(GOLDEN, "watersurface_u8.npz"))["ImData"])


@pytest.fixture(scope="session")
def highway_fixture():
    return np.load(os.path.join(GOLDEN, "highway_half_u8.npz"))


def crop_D(cube_u8, crop):
    """LSD() pre-processing (normalise, mean-subtract, F-reshape) on a crop -- oracle side."""
    from oracle import alm_oracle as O
    r0, r1, c0, c1, t0, t1 = [int(v) for v in crop]
    D, _x, _mean = O.normalize_and_center(cube_u8[r0:r1, c0:c1, t0:t1])
    return D, (r1 - r0, c1 - c0, t1 - t0)


def rel_fro(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) /
                 max(np.linalg.norm(np.asarray(b, dtype=np.float64)), 1e-300))
