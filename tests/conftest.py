import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class GoldenScenes:
    """tests/golden/geometry.npz: outputs of the reference itself (oracle/make_golden.py)."""

    def __init__(self, fname='geometry.npz'):
        self.z = np.load(os.path.join(GOLDEN, fname))
        self.names = [str(n) for n in self.z['names']]

    def scene(self, name):
        z = self.z
        g = lambda k: z[f'{name}/{k}']
        counts = g('counts')
        centers = [g('centers')[c, :counts[c]] for c in range(3)]
        Ks = [g('Ks')[c] for c in range(3)]
        RTs = [g('RTs')[c] for c in range(3)]
        ref = {k: g(f'ref_{k}') for k in ('F', 'cost', 'idx', 'X', 'reproj', 'boxes', 'centroids')}
        return dict(name=name, Ks=Ks, RTs=RTs, Ks_arr=g('Ks'), RTs_arr=g('RTs'), boxes=g('boxes'),
                    centers_arr=g('centers'), counts=counts, centers=centers, ref=ref)


@pytest.fixture(scope='session')
def golden_scenes():
    return GoldenScenes()


@pytest.fixture(scope='session')
def golden_skew():
    """tests/golden/skew.npz: reference outputs for skewed / general intrinsics (np.linalg.inv is a real LU there)."""
    return GoldenScenes('skew.npz')


@pytest.fixture(scope='session')
def golden_detect():
    """tests/golden/detect.npz: the reference's PoseEstimator._detect run unbound on a fake detector."""
    return np.load(os.path.join(GOLDEN, 'detect.npz'))


@pytest.fixture(scope='session')
def golden_crops():
    return np.load(os.path.join(GOLDEN, 'crops.npz'))


@pytest.fixture(scope='session')
def golden_bop():
    return np.load(os.path.join(GOLDEN, 'bop_scene.npz'))


@pytest.fixture(scope='session')
def golden_train():
    return np.load(os.path.join(GOLDEN, 'train_crops.npz'))
