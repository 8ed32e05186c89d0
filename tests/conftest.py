import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # no test may hang the GPU box: pytest-timeout's per-test limit unless the command line or a marker sets one
    if config.pluginmanager.hasplugin("timeout") and not getattr(config.option, "timeout", None):
        config.option.timeout = float(os.environ.get("NFST_TEST_TIMEOUT", "600"))


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; on a box without CUDA they are skipped rather
    # than failed so that a plain `pytest tests/` stays green in the build container.
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
