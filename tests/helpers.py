"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from tests.conftest import GOLDEN


def golden(pattern):
    return sorted(glob.glob(os.path.join(GOLDEN, pattern)))


def load(path):
    return np.load(path)


def maxdiff(a, b):
    """max |a-b| over finite entries; inf if the NaN positions differ."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    if (np.isnan(a) != np.isnan(b)).any():
        return float("inf")
    m = ~np.isnan(a)
    return float(np.abs(a[m] - b[m]).max()) if m.any() else 0.0


def name(path):
    return os.path.splitext(os.path.basename(path))[0]
