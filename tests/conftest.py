import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) where there is no CUDA device.  On a GPU box they always run: a missing
    libsgrace_b200.so must fail loudly there, not skip."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="gpu test: no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_terminal_summary(terminalreporter):
    """Plain element-wise relative errors of the float comparisons (tests/util.py:assert_close_f32), worst first."""
    try:
        from tests import util as U
    except Exception:
        return
    if not U.ELEMENTWISE_LOG:
        return
    worst = sorted(U.ELEMENTWISE_LOG, key=lambda t: -t[1])[:8]
    terminalreporter.write_line("plain element-wise relative error (elements >= 1e-3 of the row max), worst of "
                                f"{len(U.ELEMENTWISE_LOG)} float comparisons:")
    for what, rel, n in worst:
        terminalreporter.write_line(f"  {rel:.3e}  over {n} elements  {what}")
