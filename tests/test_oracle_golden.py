"""The oracle against outputs of the reference itself (tests/golden, made by oracle/make_golden.py).

Integer results (match indices, uint8 crops) must be identical.  Float64 intermediates are the
same NumPy calls as the reference, so on the machine that made the fixtures they are bit-equal;
on another host BLAS may pick other kernels (different FMA order), hence the 1e-12 allowance.
"""
import numpy as np
import pytest

from oracle import crop as ocrop
from oracle import geometry as og


def _names():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'geometry.npz'))
    return [str(n) for n in z['names']]


@pytest.mark.parametrize('name', _names())
def test_geometry_scene(golden_scenes, name):
    sc = golden_scenes.scene(name)
    ref = sc['ref']
    res = og.match_scene(sc['Ks'], sc['RTs'], sc['centers'], threshold=30)
    np.testing.assert_allclose(res['F'], ref['F'], rtol=1e-12, atol=0)
    assert np.array_equal(res['idx'], ref['idx'])
    if len(ref['idx']):
        F12, F13, F23 = res['F']
        cost = og.cost_tensor(*sc['centers'], F12, F13, F23)
        assert cost.dtype == np.float32 and cost.shape == ref['cost'].shape
        # f32 cost: equal up to 1 ulp on foreign hosts, bit-equal where the fixtures were made
        np.testing.assert_allclose(cost, ref['cost'], rtol=2e-7, atol=0)
        np.testing.assert_allclose(res['X'], ref['X'], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(res['reproj'], ref['reproj'], rtol=1e-6, atol=1e-9)
        np.testing.assert_array_equal(res['cost'], cost[tuple(res['idx'].T)])


def test_skewed_intrinsics(golden_skew):
    """K with skew / a general 3x3: np.linalg.inv computes in float64 and rounds the result to float32."""
    for name in golden_skew.names:
        sc = golden_skew.scene(name)
        ref = sc['ref']
        res = og.match_scene(sc['Ks'], sc['RTs'], sc['centers'], threshold=30)
        np.testing.assert_allclose(res['F'], ref['F'], rtol=1e-12, atol=0)
        assert np.array_equal(res['idx'], ref['idx'])
        np.testing.assert_allclose(res['X'], ref['X'], rtol=1e-9, atol=1e-9)
        Kinv = golden_skew.z[f'{name}/ref_Kinv']
        for c in range(3):
            exact = np.linalg.inv(sc['Ks'][c].astype(np.float64)).astype(np.float32)
            assert np.array_equal(exact.view(np.uint32), Kinv[c].view(np.uint32))


def test_detect_restatement(golden_detect):
    """oracle.detect against PoseEstimator._detect itself (process_pose.py:113-142, run unbound on a fake detector)."""
    from oracle import detect as odet
    g = golden_detect
    S = g['nraw'].shape[0]
    kept = 0
    for s in range(S):
        for c in range(3):
            n = int(g['nraw'][s, c])
            got = odet.detections_from_yolo(g['xyxy'][s, c, :n], g['conf'][s, c, :n], g['cls'][s, c, :n], float(g['thresh']))
            assert len(got) == len(g[f'bbox_{s}_{c}'])
            for d, det in enumerate(got):
                assert det['bbox'] == tuple(int(v) for v in g[f'bbox_{s}_{c}'][d])
                assert det['bb_center'] == tuple(float(v) for v in g[f'center_{s}_{c}'][d])
            kept += len(got)
    assert kept > 50 and len(g['bbox_2_0']) == 0 and len(g['bbox_3_0']) == 0
    assert tuple(g['bbox_0_0'][0]) == (0, -1, 10, 20) or g['cls'][0, 0, 0] != 0     # int() truncates toward zero
    assert len(g['bbox_0_0']) >= int((g['cls'][0, 0, :6] == 0).sum())                # conf == thresh is kept


def test_cost_tensor_equals_loop(golden_scenes):
    """the separable restatement is bit-identical to the reference's triple loop."""
    for name in ('clean10_0', 'drop12_1', 'dup8_0', 'tiny3_3'):
        sc = golden_scenes.scene(name)
        F12, F13, F23 = og.scene_fundamentals(sc['Ks'], sc['RTs'])
        a = og.cost_tensor(*sc['centers'], F12, F13, F23)
        b = og.cost_tensor_loop(*sc['centers'], F12, F13, F23)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_bop_scene(golden_bop):
    """Config 1 (single IPD-style scene via Capture.from_dir): dtype flow f32 K, f64 RT."""
    g = golden_bop
    assert g['Ks'].dtype == np.float32 and g['RTs'].dtype == np.float64
    counts = g['counts']
    centers = [g['centers'][c, :counts[c]] for c in range(3)]
    res = og.match_scene(list(g['Ks']), list(g['RTs']), centers, threshold=30)
    assert np.array_equal(res['idx'], g['ref_idx'])
    assert len(res['idx']) == 10
    np.testing.assert_allclose(res['X'], g['ref_X'], rtol=1e-9, atol=1e-9)


def test_sentinel_and_degenerate_lines():
    """norm <= 1e-8 -> distance 9999 (epipolar_matching.py:25-26)."""
    F = np.zeros((3, 3))
    assert og.epipolar_error((1.0, 2.0), (3.0, 4.0), F) == 9999.0
    F = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])     # pure x-translation: l = (0, -1, y)
    e = og.epipolar_error((10.0, 20.0), (30.0, 26.0), F)
    assert e == 6.0


@pytest.mark.parametrize('T', [224, 256, 64])
def test_letterbox_u8(golden_crops, T):
    g = golden_crops
    boxes = g['boxes'] if T != 256 else g['boxes'][::2]
    for b, canvas, geom in zip(boxes, g[f'canvas_T{T}'], g[f'geom_T{T}']):
        x1, y1, x2, y2 = [int(v) for v in b]
        crop = g['image'][y1:y2, x1:x2]
        for fn in (ocrop.letterbox_ref, ocrop.letterbox_spec):
            got, scale, dx, dy = fn(crop, target_size=T)
            assert np.array_equal(got, canvas), (fn.__name__, b)
            assert (scale, dx, dy) == tuple(geom)


def test_normalise_lut_and_tensor(golden_crops):
    g = golden_crops
    lut = ocrop.normalise_lut()
    assert np.array_equal(lut.view(np.uint32), g['lut'].view(np.uint32))
    for b, want in zip(g['boxes'][:4], g['tensor_T64']):
        for fn in (ocrop.crop_tensor_ref, ocrop.crop_tensor_spec):
            got = fn(g['image'], b, target_size=64, swap_rb=True)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), fn.__name__
    # white padding constants quoted in SURVEY.md section 8 (a12)
    np.testing.assert_allclose(lut[:, 255], [2.2489083, 2.4285715, 2.6400001], rtol=1e-7)
