import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    from unitex_b200 import _lib, build
    if not _lib.lib_path().exists():
        build.build()
    return _lib.load()


_PARITY_LINES = []


@pytest.fixture
def parity_log():
    """Tests hand measured parity figures (dB, counts) to the terminal summary, so they show in the run's tail."""
    return _PARITY_LINES.append


def pytest_terminal_summary(terminalreporter):
    if _PARITY_LINES:
        terminalreporter.write_sep("-", "parity figures")
        for ln in _PARITY_LINES:
            terminalreporter.write_line("parity: " + ln)
