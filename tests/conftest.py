import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def oracle():
    """The CPU oracle (compiled on first use)."""
    import oracle as oracle_module
    oracle_module.host.build()
    return oracle_module.host


@pytest.fixture(scope='session')
def gpu():
    """(context, command_queue) on CUDA device 0 through the C ABI; fails (not skips)
    when the library or a device is missing -- there is no CPU fallback."""
    from katsdpimager_b200 import accel
    context = accel.Context(int(os.environ.get('KIB_DEVICE', '0')))
    return context, context.create_command_queue()
