"""pytest configuration: the ``gpu`` marker and the import paths.

* ``multimodal-baselines_b200/`` is put on ``sys.path`` so that the drop-in modules are
  imported by the reference's own top-level names (``import sif_functions``, ``import
  losses`` ...), exactly as a user of the reference would.
* ``oracle/`` is importable as the ``oracle`` package from the repo root (tests only).
* Tests marked ``gpu`` are skipped automatically where there is no CUDA device.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'multimodal-baselines_b200')
for p in (os.path.join(ROOT, 'tests', 'golden'), ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
