import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN_DIR, name + '.pt'), weights_only=False)


def golden_names(prefix=''):
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith('.pt') and f.startswith(prefix))
