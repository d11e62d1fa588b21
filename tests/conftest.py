import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / 'faster-gaussian-splatting_b200'):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def pytest_collection_modifyitems(config, items):
    """A per-test time limit (pytest-timeout, if installed): the emulated-kernel tests run hundreds of host threads with
    barriers -- a deadlock must fail the test, not hang the suite."""
    if not config.pluginmanager.hasplugin('timeout'):
        return
    for item in items:
        if item.get_closest_marker('timeout') is None:
            # method 'thread': the main thread sits inside a ctypes call while the emulator runs, where a signal handler
            # would never get to run
            item.add_marker(pytest.mark.timeout(1800, method='thread'))
