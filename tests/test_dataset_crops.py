"""Training-side crop transform: BOPSingleObjDataset.__getitem__, data_utils.py:233-298.

Golden: tests/golden/train_crops.npz, recorded from the reference's dataset class itself (oracle/make_golden.py).
"""
import random

import numpy as np
import pytest
import torch

from oracle import crop as ocrop
from tests.gpu_util import to_dev


def test_oracle_windows_and_canvases_match_reference(golden_train):
    g = golden_train
    image, T = g['image'], int(g['T'])
    H, W = image.shape[:2]
    for r, bbox in enumerate(g['bbox']):
        x1, y1, x2, y2 = ocrop.dataset_window(bbox, W, H)
        assert (y2 - y1, x2 - x1) == tuple(g['orig_hw'][r])
        assert np.array_equal(ocrop.crop_u8_ref(image, (x1, y1, x2, y2), T), g['orig_canvas'][r])
        x1, y1, x2, y2 = ocrop.dataset_window(bbox, W, H, g['scale'][r], g['shift'][r])
        assert (y2 - y1, x2 - x1) == tuple(g['aug_hw'][r]), r
        assert np.array_equal(ocrop.crop_u8_ref(image, (x1, y1, x2, y2), T), g['aug_canvas'][r]), r
    for r in range(len(g['orig_t'])):
        want = ocrop.dataset_tensor_ref(image, ocrop.dataset_window(g['bbox'][r], W, H), T)
        assert np.array_equal(want.view(np.uint32), g['orig_t'][r].view(np.uint32))


def test_jitter_draws_follow_the_reference_stream(golden_train):
    """draw_crop_jitter consumes Python's `random` exactly like data_utils.py:257-263."""
    from bpc_baseline_b200.utils.data_utils import draw_crop_jitter
    g = golden_train
    for r, bbox in enumerate(g['bbox']):
        random.seed(1000 + r)
        scale, shift = draw_crop_jitter([bbox])
        assert scale[0] == g['scale'][r] and tuple(shift[0]) == tuple(g['shift'][r])
    rnd = random.Random(7)
    a = draw_crop_jitter(g['bbox'][:5], rnd=rnd)
    rnd = random.Random(7)
    b = [draw_crop_jitter([bb], rnd=rnd) for bb in g['bbox'][:5]]
    assert np.array_equal(a[0], np.concatenate([x[0] for x in b])) and np.array_equal(a[1], np.concatenate([x[1] for x in b]))


@pytest.mark.gpu
def test_dataset_crops_match_reference(golden_train):
    from bpc_baseline_b200 import batched
    from bpc_baseline_b200.utils.data_utils import dataset_crops
    g = golden_train
    image, T = g['image'], int(g['T'])
    H, W = image.shape[:2]
    got = dataset_crops(image[None], g['bbox'], target_size=T, as_uint8=True).cpu().numpy()
    assert np.array_equal(got, g['orig_canvas'])
    got = dataset_crops(image[None], g['bbox'], target_size=T, jitter=(g['scale'], g['shift']), as_uint8=True).cpu().numpy()
    assert np.array_equal(got, g['aug_canvas'])
    t = dataset_crops(image[None], g['bbox'][:4], target_size=T).cpu().numpy()
    assert np.array_equal(t.view(np.uint32), g['orig_t'].view(np.uint32))
    # the ROI records themselves, incl. a fuzz of the clamps against the oracle restatement
    rng = np.random.default_rng(3)
    n = 4000
    xywh = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(1, 400, n), rng.integers(1, 400, n)], 1).astype(np.int32)
    scale = 1.0 + 0.2 * rng.random(n)
    scale[:50] = (np.arange(50) * 2 + 1) / (2.0 * xywh[:50, 2])               # w * scale lands on .5: half-to-even
    shift = rng.integers(-40, 41, (n, 2)).astype(np.int32)
    rois = batched.train_rois(to_dev(xywh), W, H, scale=to_dev(scale), shift=to_dev(shift)).cpu().numpy()
    plain = batched.train_rois(to_dev(xywh), W, H).cpu().numpy()
    for r in range(n):
        assert tuple(rois[r, 1:]) == ocrop.dataset_window(xywh[r], W, H, scale[r], shift[r]), r
        assert tuple(plain[r, 1:]) == ocrop.dataset_window(xywh[r], W, H), r


@pytest.mark.gpu
def test_dataset_crops_empty_window_raises(golden_train):
    from bpc_baseline_b200.utils.data_utils import dataset_crops
    image = golden_train['image']
    H, W = image.shape[:2]
    with pytest.raises(RuntimeError, match='Empty crop'):
        dataset_crops(image[None], [[W, 10, 20, 20]], target_size=64)
