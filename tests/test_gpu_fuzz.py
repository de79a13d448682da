"""Property / fuzz tests of the CUDA path against the oracle (hypothesis): ragged N != M != P, N*M < P, exact
ties, sentinel costs, thin / tiny / huge boxes in all three INTER_AREA regimes."""
import os

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import crop as ocrop
from oracle import geometry as og
from tests.gpu_util import to_dev

pytestmark = pytest.mark.gpu
# BPC_FUZZ_EXAMPLES=2000 widens the search (and draws fresh examples) for a one-off soak run
_N = int(os.environ.get('BPC_FUZZ_EXAMPLES', '40'))
FUZZ = settings(max_examples=_N, deadline=None, suppress_health_check=list(HealthCheck), derandomize=_N == 40)


@FUZZ
@given(st.integers(1, 6), st.integers(1, 6), st.integers(1, 9), st.integers(0, 3), st.integers(0, 2 ** 31 - 1))
def test_match_objects_any_shape(N, M, P, mode, seed):
    """Explicit cost tensors of any shape, incl. N*M < P (no transposition inside SciPy) and heavy ties."""
    from bpc_baseline_b200 import batched
    rng = np.random.default_rng(seed)
    if mode == 0:
        cost = (rng.random((N, M, P)) * 60).astype(np.float32)
    elif mode == 1:
        cost = rng.integers(0, 3, (N, M, P)).astype(np.float32) * 10
    elif mode == 2:
        cost = rng.integers(0, 4, (N, M, P)).astype(np.float32)
        cost[rng.random((N, M, P)) < 0.4] = 9999
    else:
        cost = np.tile((rng.random((1, M, P)) * 40).astype(np.float32), (N, 1, 1))      # duplicate first-camera detections
    idx, n = batched.match_objects(to_dev(cost[None]), 30)
    n = int(n.cpu()[0])
    assert [tuple(r) for r in idx.cpu().numpy()[0, :n]] == og.match_objects(cost, 30)


@FUZZ
@given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 9), st.floats(0.0, 0.4), st.integers(0, 2), st.integers(0, 10 ** 6))
def test_full_path_ragged_scenes(n1, n2, n3, p_drop, n_dup, seed):
    """Ragged scenes with duplicates: indices exact, costs bit-equal, X within tolerance."""
    from bpc_baseline_b200 import batched, synth
    D = max(n1, n2, n3)
    batch = synth.make_scenes(1, D, seed=seed, p_drop=p_drop, n_dup=n_dup)
    batch.counts[0] = np.minimum(batch.counts[0], [n1, n2, n3])
    res = batched.match_triangulate(to_dev(batch.Ks), to_dev(batch.RTs), to_dev(batch.centers), to_dev(batch.counts), 30)
    n = int(res.n.cpu()[0])
    Ks, RTs = batch.capture_arrays(0)
    cen = [batch.centers[0, c, :batch.counts[0, c]] for c in range(3)]
    want = og.match_scene(Ks, RTs, cen, 30)
    assert n == len(want['idx'])
    assert np.array_equal(res.idx.cpu().numpy()[0, :n], want['idx'])
    if n:
        assert np.array_equal(res.cost.cpu().numpy()[0, :n].view(np.uint32), want['cost'].view(np.uint32))
        err = np.linalg.norm(res.X.cpu().numpy()[0, :n] - want['X'], axis=1) / np.linalg.norm(want['X'], axis=1)
        assert err.max() < 1e-4


_IMG = {}


def _image(width):
    """width 1500: image pitch 4500 B (1-D bulk-copy staging); 1504: 4512 B, a multiple of 16 (2-D tensor-map staging)."""
    if width not in _IMG:
        from bpc_baseline_b200 import synth
        img = synth.make_images(1, seed=77, width=width, height=1100)
        _IMG[width] = (img, to_dev(img))
    return _IMG[width]


@FUZZ
@given(st.integers(1, 1400), st.integers(1, 1000), st.sampled_from([224, 256, 96, 33, 20]), st.integers(0, 10 ** 6),
       st.sampled_from([1500, 1504]))
def test_crop_any_box(w, h, T, seed, width):
    """Any box (1 px .. nearly the whole image) at several target sizes: uint8 result identical to cv2's, or the
    ROI is rejected exactly when the reference would fail (a resized side of 0)."""
    from bpc_baseline_b200 import batched
    img, dev = _image(width)
    rng = np.random.default_rng(seed)
    x1 = int(rng.integers(0, width - w + 1)); y1 = int(rng.integers(0, 1100 - h + 1))
    roi = np.array([[0, x1, y1, x1 + w, y1 + h]], np.int32)
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    got = batched.roi_crop_u8(dev, to_dev(roi), T=T, status=status).cpu().numpy()[0]
    _, nw, nh, _, _ = ocrop.letterbox_geometry(h, w, T)
    if nw < 1 or nh < 1:
        assert int(status[0]) == 1 and np.all(got == 255)
        return
    assert int(status[0]) == 0
    want = ocrop.crop_u8_ref(img[0], (x1, y1, x1 + w, y1 + h), T)
    assert np.array_equal(got, want), (w, h, T, x1, y1, width, int(np.abs(got.astype(int) - want).max()))
