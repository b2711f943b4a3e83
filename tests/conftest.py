import json
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bits_to_f32(a):
    return np.asarray(a, dtype=np.uint32).view(np.float32)


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by oracle/ref_crosscheck.cpp from the reference's own headers."""
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as f:
        return json.load(f)


def ulp_diff(a, b):
    """Distance in units in the last place between float32 arrays (NaNs must match)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)


def rel_rmse(a, b):
    """sqrt(mean((a-b)^2) / mean(b^2)) over RGB: the RMS of the per-pixel error relative to the RMS of the reference image
    (one bright outlier pixel dominates it, which is the point of the gate).  The other reading of north_star's
    "per-pixel relative RMSE" -- the RMS over pixels of |a-b| / |b| -- is per_pixel_rel_rmse below; the full-size C2 test
    gates both."""
    a = np.asarray(a, dtype=np.float64)[..., :3]
    b = np.asarray(b, dtype=np.float64)[..., :3]
    return float(np.sqrt(np.mean((a - b) ** 2) / max(np.mean(b**2), 1e-30)))


def per_pixel_rel_rmse(a, b, floor=1e-6):
    """sqrt(mean_pixels((|a-b| / max(|b|, floor))^2)) over RGB vectors"""
    a = np.asarray(a, dtype=np.float64)[..., :3]
    b = np.asarray(b, dtype=np.float64)[..., :3]
    num = np.linalg.norm(a - b, axis=-1)
    den = np.maximum(np.linalg.norm(b, axis=-1), floor)
    return float(np.sqrt(np.mean((num / den) ** 2)))


@pytest.fixture(scope="session")
def gpu_render():
    """A Render (eCompute backend) on cuda:0, created once per session for the -m gpu tests."""
    from strelka_b200 import RenderFactory, RenderType

    r = RenderFactory.createRender(RenderType.eCompute)
    r.init()
    yield r
    r.destroy()
