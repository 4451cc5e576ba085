import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "first_hw_run(reason): GPU test of code that has not run on hardware yet; it is "
                            "collected LAST and reported as xfail/xpass (non-strict) until a GPU run has confirmed it")


def pytest_collection_modifyitems(config, items):
    import torch
    # tests of not-yet-hardware-verified code run after everything that is verified, and cannot turn the suite red
    pending = [i for i in items if i.get_closest_marker("first_hw_run")]
    if pending:
        items[:] = [i for i in items if not i.get_closest_marker("first_hw_run")] + pending
        for i in pending:
            m = i.get_closest_marker("first_hw_run")
            i.add_marker(pytest.mark.xfail(reason=m.kwargs.get("reason", "first hardware run pending"), strict=False))
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
