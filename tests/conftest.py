import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def import_reference():
    """Make the unmodified reference (pip --target install under baseline/_ref) importable, with the
    scipy>=1.15 sph_harm shim.  Returns False when it is not available."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "kymatio")):
        return False
    import scipy.special
    if not hasattr(scipy.special, "sph_harm"):
        scipy.special.sph_harm = lambda m, n, az, pol: scipy.special.sph_harm_y(n, m, pol, az)
    if ref not in sys.path:
        sys.path.insert(0, ref)
    return True
