import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle():
    """the CPU oracle (test infrastructure): builds oracle/libnfx_oracle.so on first use"""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope='session')
def gpu():
    """the product module; fails loudly when the CUDA library or a device is missing"""
    import torch
    assert torch.cuda.is_available(), 'a CUDA device is required for -m gpu tests'
    from nemoflux_b200 import nemoflux_gpu
    return nemoflux_gpu


@pytest.fixture(autouse=True)
def _guard_bands(request):
    """NFX_DEBUG_GUARDS=1 (the stand-in for compute-sanitizer memcheck, which this pool refuses): after every GPU test
    the guard bands around every device buffer the library owns must be intact"""
    yield
    if os.environ.get('NFX_DEBUG_GUARDS', '0') not in ('', '0') and request.node.get_closest_marker('gpu') is not None:
        from nemoflux_b200 import _lib
        _lib.check_guards()        # raises when a band is damaged; buffers freed during the test were checked at release
