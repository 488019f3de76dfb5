import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed_global_rng(request):
    """Module constructors draw their initial weights from torch's global RNG: seed it per test
    (from the test id, not from Python's salted hash) so that every run -- here and on the GPU box --
    builds the same models.  Without this each process tested a different random instance, and the
    three-layer encoder test, whose deep gradients are sensitive to samples sitting next to pixel
    boundaries, passed or failed depending on the draw."""
    import zlib
    import torch
    torch.manual_seed(zlib.crc32(request.node.nodeid.encode()) & 0x7fffffff)
    yield
