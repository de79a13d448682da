"""The reference's Python call surface (bpc_baseline_b200.inference / .utils) against the reference's own
outputs recorded in tests/golden -- these read like tests the reference could have shipped."""
import types
from types import SimpleNamespace

import numpy as np
import pytest

from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _dets(sc):
    out = {}
    for c in range(3):
        n = int(sc['counts'][c])
        out[c] = [{'bbox': tuple(int(v) for v in sc['boxes'][c, d]),
                   'bb_center': (float(sc['centers_arr'][c, d, 0]), float(sc['centers_arr'][c, d, 1]))} for d in range(n)]
    return out


def test_compute_fundamental_matrix(golden_scenes):
    from bpc_baseline_b200.inference.utils.camera_utils import compute_fundamental_matrix
    sc = golden_scenes.scene('clean10_0')
    K, RT = sc['Ks'], sc['RTs']
    pairs = [(0, 1), (0, 2), (1, 2)]
    for p, (a, b) in enumerate(pairs):
        F = compute_fundamental_matrix(K[a], RT[a][:3, :3], RT[a][:3, 3], K[b], RT[b][:3, :3], RT[b][:3, 3])
        assert F.shape == (3, 3) and F.dtype == np.float64
        np.testing.assert_allclose(F, sc['ref']['F'][p], rtol=1e-12, atol=0)
    with pytest.raises(TypeError):
        compute_fundamental_matrix(K[0].astype(np.float64), RT[0][:3, :3], RT[0][:3, 3], K[1], RT[1][:3, :3], RT[1][:3, 3])


def test_epipolar_error_and_full(golden_scenes):
    from bpc_baseline_b200.inference.epipolar_matching import epipolar_error, epipolar_error_full
    from oracle import geometry as og
    sc = golden_scenes.scene('drop12_0')
    F12, F13, F23 = sc['ref']['F']
    c1, c2, c3 = sc['centers']
    for i in range(3):
        for j in range(3):
            e = epipolar_error(tuple(c1[i]), tuple(c2[j]), F12)
            assert isinstance(e, float) and e == og.epipolar_error(c1[i], c2[j], F12)
            full = epipolar_error_full(tuple(c1[i]), tuple(c2[j]), tuple(c3[i]), F12, F13, F23)
            assert np.float32(full) == sc['ref']['cost'][i, j, i]
    assert epipolar_error((1.0, 2.0), (3.0, 4.0), np.zeros((3, 3))) == 9999.0          # degenerate lines -> sentinel
    with pytest.raises(NameError):                                                       # the reference's broken viz branch
        epipolar_error((1.0, 2.0), (3.0, 4.0), F12, img1=np.zeros((4, 4, 3)), img2=np.zeros((4, 4, 3)))


@pytest.mark.parametrize('name', ['clean10_1', 'drop12_3', 'dup8_2', 'false9_1', 'tiny3_0', 'tiny3_3', 'tiny3_4'])
def test_cost_matrix_match_objects_sort(golden_scenes, name):
    from bpc_baseline_b200.inference.epipolar_matching import compute_cost_matrix, match_objects
    sc = golden_scenes.scene(name)
    d = _dets(sc)
    F12, F13, F23 = sc['ref']['F']
    cost = compute_cost_matrix(d[0], d[1], d[2], F12, F13, F23)
    assert cost.dtype == np.float32 and cost.shape == sc['ref']['cost'].shape
    assert np.array_equal(cost.view(np.uint32), sc['ref']['cost'].view(np.uint32))
    matches = match_objects(cost, 30)
    assert all(isinstance(m, tuple) and len(m) == 3 for m in matches)
    matches_sorted = sorted(matches, key=lambda t: cost[t[0], t[1], t[2]])           # process_pose.py:183
    assert [tuple(int(v) for v in m) for m in matches_sorted] == [tuple(r) for r in sc['ref']['idx']]
    assert match_objects(np.zeros((0, 3, 2), np.float32), 30) == []
    with pytest.raises(ValueError):
        match_objects(np.full((2, 2, 2), np.nan, np.float32), 30)


def test_triangulate_and_reprojection(golden_scenes):
    from bpc_baseline_b200.inference.epipolar_matching import triangulate_multi_view
    from bpc_baseline_b200.inference.utils import triangulation as tri
    from oracle import geometry as og
    sc = golden_scenes.scene('clean20_1')
    Ps = og.projection_matrices(sc['Ks'], sc['RTs'])
    ref = sc['ref']
    for m in range(5):
        X = triangulate_multi_view(Ps, ref['centroids'][m])
        assert X.shape == (3,) and rel_err(X[None], ref['X'][m][None])[0] < 1e-9
        X2 = tri.triangulate_multi_view(Ps[:2], ref['centroids'][m][:2])              # two views
        assert rel_err(X2[None], og.triangulate_multi_view(Ps[:2], ref['centroids'][m][:2])[None])[0] < 1e-7
        for v in range(3):
            e = tri.compute_reprojection_error(Ps[v], ref['X'][m], ref['centroids'][m, v])
            assert abs(e - ref['reproj'][m, v]) < 1e-7


@pytest.mark.parametrize('name', ['clean10_0', 'clean20_0', 'drop12_4', 'dup8_0', 'false9_3', 'tiny3_5', 'dense40_0'])
def test_pose_estimator_match(golden_scenes, name):
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams, PosePrediction
    sc = golden_scenes.scene(name)
    est = PoseEstimator(PoseEstimatorParams())
    capture = SimpleNamespace(images=[None] * 3, Ks=sc['Ks'], RTs=sc['RTs'])
    preds = est._match(capture, _dets(sc))
    ref = sc['ref']
    assert len(preds) == len(ref['idx'])
    for p, idx, boxes, cen, X in zip(preds, ref['idx'], ref['boxes'], ref['centroids'], ref['X']):
        assert isinstance(p, PosePrediction) and p.match == tuple(idx)
        assert p.boxes.shape == (3, 4) and np.array_equal(p.boxes, boxes)
        assert p.centroids.dtype == np.float64 and np.array_equal(p.centroids, cen)
        assert rel_err(p.t[None], X[None])[0] < 1e-9
    if len(preds):
        again = PosePrediction([{'bbox': tuple(b), 'bb_center': tuple(c)} for b, c in zip(preds[0].boxes, preds[0].centroids)], capture)
        assert rel_err(again.t[None], ref['X'][0][None])[0] < 1e-9


@pytest.mark.parametrize('T', [224, 256])
def test_letterbox_and_crop_inputs(golden_crops, T):
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams
    from bpc_baseline_b200.utils.data_utils import letterbox_preserving_aspect_ratio
    from oracle import crop as ocrop
    g = golden_crops
    image = g['image']
    boxes = g['boxes'] if T != 256 else g['boxes'][::2]
    for b, canvas, geom in list(zip(boxes, g[f'canvas_T{T}'], g[f'geom_T{T}']))[:8]:
        x1, y1, x2, y2 = [int(v) for v in b]
        got, scale, dx, dy = letterbox_preserving_aspect_ratio(image[y1:y2, x1:x2], target_size=T)     # a strided view, as the reference passes
        assert got.dtype == np.uint8 and np.array_equal(got, canvas) and (scale, dx, dy) == tuple(geom)
    with pytest.raises(ZeroDivisionError):                          # max(h, w) == 0, data_utils.py:36
        letterbox_preserving_aspect_ratio(image[5:5, 3:3], target_size=T)
    import cv2
    with pytest.raises(cv2.error):                                  # 0 x 6 crop: cv2.resize rejects the empty size
        letterbox_preserving_aspect_ratio(image[5:5, 3:9], target_size=T)
    # crop half of _estimate_rotation: 2 predictions x 3 views
    est = PoseEstimator(PoseEstimatorParams(target_size=T))
    capture = SimpleNamespace(images=[image, image, image], Ks=None, RTs=None)
    preds = [SimpleNamespace(boxes=np.asarray(g['boxes'][0:3], np.int64), centroids=np.zeros((3, 2)), capture=capture),
             SimpleNamespace(boxes=np.asarray(g['boxes'][3:6], np.int64), centroids=np.zeros((3, 2)), capture=capture)]
    tens = est.crop_inputs(preds)
    assert tuple(tens.shape) == (6, 3, T, T) and tens.is_cuda
    tens = tens.cpu().numpy()
    for r in range(6):
        want = ocrop.crop_tensor_ref(image, g['boxes'][r], target_size=T, swap_rb=True)
        assert np.array_equal(tens[r].view(np.uint32), want.view(np.uint32))


def test_estimate_rotation_with_stub_network(golden_crops, golden_scenes):
    """Network half with a stand-in model: rotation decode + pose assembly (process_pose.py:214-239)."""
    import torch
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams
    sc = golden_scenes.scene('clean10_0')
    image = golden_crops['image']
    capture = SimpleNamespace(images=[image] * 3, Ks=sc['Ks'], RTs=sc['RTs'])
    pred = SimpleNamespace(boxes=np.asarray(golden_crops['boxes'][0:3], np.int64), centroids=np.zeros((3, 2)), capture=capture,
                           t=np.array([1.0, 2.0, 3.0]))

    class Stub(torch.nn.Module):
        def forward(self, x):
            return x.mean(dim=(2, 3))[:, [0, 1, 2, 0]] + torch.tensor([0.1, 0.2, 0.3, 1.0], device=x.device)
    est = PoseEstimator(PoseEstimatorParams(target_size=64), pose_model=Stub().cuda(), rotation_mode='quat')
    est._estimate_rotation([pred])
    assert len(pred.rotation_preds) == 3 and pred.pose.shape == (4, 4)
    R = pred.final_rotation
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-5)
    np.testing.assert_allclose(pred.pose[:3, 3], pred.t)


def test_install_rebinds_reference_modules(golden_scenes):
    """install() on a stand-in `bpc` package tree (the real reference is not present on the GPU box)."""
    import sys
    import bpc_baseline_b200 as pkg
    names = ['fakebpc', 'fakebpc.inference', 'fakebpc.inference.utils', 'fakebpc.inference.utils.camera_utils',
             'fakebpc.inference.epipolar_matching', 'fakebpc.inference.process_pose', 'fakebpc.utils', 'fakebpc.utils.data_utils']
    mods = {n: types.ModuleType(n) for n in names}
    sentinel = lambda *a, **k: 'reference'
    mods['fakebpc.inference.utils.camera_utils'].compute_fundamental_matrix = sentinel
    mods['fakebpc.inference.epipolar_matching'].match_objects = sentinel
    mods['fakebpc.inference.process_pose'].match_objects = sentinel

    class RefEstimator:
        def _match(self, capture, detections):
            return 'reference'
    mods['fakebpc.inference.process_pose'].PoseEstimator = RefEstimator
    sys.modules.update(mods)
    try:
        done = pkg.install('fakebpc')
        assert 'fakebpc.inference.process_pose.PoseEstimator._match' in done
        assert mods['fakebpc.inference.epipolar_matching'].match_objects is not sentinel
        sc = golden_scenes.scene('clean10_2')
        est = RefEstimator()
        est.params = SimpleNamespace(matching_threshold=30)
        preds = est._match(SimpleNamespace(images=[None] * 3, Ks=sc['Ks'], RTs=sc['RTs']), _dets(sc))
        assert [p.match for p in preds] == [tuple(r) for r in sc['ref']['idx']]
        pkg.uninstall()
        assert mods['fakebpc.inference.epipolar_matching'].match_objects is sentinel
        assert RefEstimator()._match(None, None) == 'reference'
    finally:
        pkg.uninstall()
        for n in names:
            sys.modules.pop(n, None)


def test_load_scene_batch_equals_capture_from_dir(tmp_path):
    """f4: the batched BOP loader's device tensors are bit-equal to Capture.from_dir / load_camera_params per scene
    (reference data_utils.py:399-409, camera_utils.py:6-20), PNG and JPEG files, and feed the matcher directly."""
    import json
    import os
    import cv2
    import torch
    from bpc_baseline_b200 import batched, synth
    from bpc_baseline_b200.utils.data_utils import Capture, load_scene_batch
    cams = ['cam1', 'cam2', 'cam3']
    rng = np.random.default_rng(11)
    dirs = []
    for sd, ext in (('a', 'png'), ('b', 'jpg')):
        d = os.path.join(str(tmp_path), sd)
        os.makedirs(d)
        dirs.append(d)
        batch = synth.make_scenes(2, 6, seed=5 + len(dirs))
        for c, cid in enumerate(cams):
            per_image = {}
            for i in (0, 3):
                K, RT = batch.Ks[i // 3, c], batch.RTs[i // 3, c]
                per_image[str(i)] = {'cam_K': [float(v) for v in K.ravel()], 'cam_R_w2c': [float(v) for v in RT[:3, :3].ravel()],
                                     'cam_t_w2c': [float(v) for v in RT[:3, 3]]}
            with open(os.path.join(d, f'scene_camera_{cid}.json'), 'w') as fh:
                json.dump(per_image, fh)
            os.makedirs(os.path.join(d, f'rgb_{cid}'))
            for i in (0, 3):
                yy, xx = np.mgrid[0:40, 0:56]
                img = np.stack([xx + yy + 5 * c + 20 * k for k in range(3)], axis=2).astype(np.uint8) if ext == 'jpg' \
                    else rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)       # JPEG: smooth ramps (decoders agree closely on those)
                cv2.imwrite(os.path.join(d, f'rgb_{cid}', f'{i:06d}.{ext}'), img)
    scene_dirs, image_ids = [dirs[0], dirs[0], dirs[1], dirs[1]], [0, 3, 3, 0]
    sb = load_scene_batch(scene_dirs, cams, image_ids)
    assert sb.Ks.is_cuda and sb.Ks.dtype == torch.float32 and tuple(sb.Ks.shape) == (4, 3, 3, 3)
    assert sb.RTs.dtype == torch.float64 and tuple(sb.RTs.shape) == (4, 3, 4, 4)
    assert sb.images.dtype == torch.uint8 and tuple(sb.images.shape) == (12, 40, 56, 3) and set(sb.decoders) == {'cv2+pinned'}
    for s, (d, i) in enumerate(zip(scene_dirs, image_ids)):
        cap = Capture.from_dir(d, cams, i, 8)
        assert np.array_equal(sb.Ks[s].cpu().numpy(), np.stack(cap.Ks))
        assert np.array_equal(sb.RTs[s].cpu().numpy(), np.stack(cap.RTs))
        assert np.array_equal(sb.images[3 * s:3 * s + 3].cpu().numpy(), np.stack(cap.images))
        got = sb.capture(s, 8)
        assert all(np.array_equal(a, b) for a, b in zip(got.images, cap.images)) and got.RTs[1].dtype == np.float64
    # the tensors are the batched API's input layout
    F = batched.fundamental(sb.Ks, sb.RTs)
    assert tuple(F.shape) == (4, 3, 3, 3) and bool(torch.isfinite(F).all())
    # nvJPEG path: same shapes, pixels close to cv2's (not bit-identical by design: a different JPEG decoder)
    sj = load_scene_batch(scene_dirs[2:], cams, image_ids[2:], decode='nvjpeg')
    assert set(sj.decoders) == {'nvjpeg'} and tuple(sj.images.shape) == (6, 40, 56, 3)
    diff = (sj.images.int() - sb.images[6:].int()).abs()
    assert float(diff.float().mean()) < 3.0                              # chroma upsampling / IDCT differ between libjpeg-turbo and nvJPEG
