import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_harness

    have_ref = ref_harness.reference_available()
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="reference tree not present"))
        if "gpu" in it.keywords and not have_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
