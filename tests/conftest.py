import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the reference checkout at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(os.path.join(os.environ.get("REINLIFE_REF", "/root/reference"), "ReinLife"))
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    try:
        import pytest_timeout  # noqa: F401
        have_timeout = True
    except Exception:
        have_timeout = False
    for item in items:
        # a hung GPU kernel (spin-wait mbarrier pipelines) must fail one test, not stall the whole run
        if have_timeout and "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(300))
        if "ref" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference checkout not present"))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
