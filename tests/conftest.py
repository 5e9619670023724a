"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"` runs in a container with no GPU: the oracle against the golden
vectors, the host logic, the loader, the C-ABI export check and the
world-size-2 gloo tests.  `-m gpu` runs the parity tests proper on a B200,
through the C ABI of libb2pt.so.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_cuda() -> bool:
    try:
        from mygpuraytracer_b200 import api

        return api.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must FAIL LOUDLY on a GPU box whose CUDA library is missing;
    # they are only skipped where there is no CUDA device at all.
    have = None
    for item in items:
        if "gpu" in item.keywords:
            if have is None:
                have = _have_cuda()
            if not have and not os.environ.get("B2PT_REQUIRE_GPU"):
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
