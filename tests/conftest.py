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
def ref_fp32():
    return dict(np.load(os.path.join(GOLDEN, "ref_fp32.npz")))


@pytest.fixture(scope="session")
def ref_channel():
    return dict(np.load(os.path.join(GOLDEN, "ref_channel.npz")))


@pytest.fixture(scope="session")
def rtl_vectors():
    import json
    return json.load(open(os.path.join(GOLDEN, "rtl_generator_vectors.json")))


@pytest.fixture(scope="session")
def rtl_critic_vectors():
    import json
    return json.load(open(os.path.join(GOLDEN, "rtl_critic_vectors.json")))


def assert_close(got, ref, rtol=1e-5, name=""):
    """|got-ref| <= rtol * max(|ref|, scale) with scale = max|ref| over the tensor: the 1e-5 relative bound of
    BASELINE.json's north_star, taken relative to the tensor's scale so exact zeros do not demand exactness."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), scale)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, f"{name}: max rel-to-scale err {worst:.3e} > {rtol:.1e} (scale {scale:.3e})"
    return worst
