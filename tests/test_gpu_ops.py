"""torch.ops.bpc_b200.* (the thin C++ extension) and the CUDA-graphed SceneSession against the ctypes path / the oracle."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import crop as ocrop
from oracle import geometry as og
from tests.gpu_util import batch_to_dev, to_dev

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.view(torch.int64) if t.dtype == torch.float64 else (t.view(torch.int32) if t.dtype == torch.float32 else t)


def test_ops_equal_the_ctypes_path_bit_for_bit():
    from bpc_baseline_b200 import batched, ops, synth
    batch = synth.make_scenes(24, 9, seed=synth.SEED + 71, p_drop=0.15, sigma=1.5)
    Ks, RTs, cen, boxes, cnt = batch_to_dev(batch)
    a = batched.match_triangulate(Ks, RTs, cen, cnt, 30, want_reproj=True, want_F=True)
    idx, n, cost, X, reproj, F = ops.match_triangulate(Ks, RTs, cen, cnt, 30, want_F=True)
    for got, want in ((idx, a.idx), (n, a.n), (cost, a.cost), (X, a.X), (reproj, a.reproj), (F, a.F)):
        assert torch.equal(_bits(got), _bits(want))
    assert torch.equal(_bits(ops.box_centers(boxes)), _bits(batched.box_centers(boxes)))
    assert torch.equal(_bits(ops.fundamental(Ks, RTs)), _bits(batched.fundamental(Ks, RTs)))
    ios = to_dev(np.tile(np.arange(3, dtype=np.int32), (24, 1)))
    rois, offs = ops.build_rois(boxes, idx, n, ios)
    rois_b, offs_b = batched.build_rois(boxes, a.idx, a.n, ios)
    total = int(offs[-1])
    assert torch.equal(offs, offs_b) and torch.equal(rois[:total], rois_b[:total]) and total > 0
    images = to_dev(synth.make_images(3, seed=5))
    crops, status = ops.roi_crop(images, rois[:total].contiguous(), 96)
    assert torch.equal(_bits(crops), _bits(batched.roi_crop(images, rois_b[:total].contiguous(), T=96))) and int(status.sum()) == 0
    u8, _ = ops.roi_crop_u8(images, rois[:total].contiguous(), 96)
    assert torch.equal(u8, batched.roi_crop_u8(images, rois_b[:total].contiguous(), T=96))
    b16, st16 = ops.roi_crop_bf16(images, rois[:total].contiguous(), 96)
    assert b16.is_contiguous(memory_format=torch.channels_last) and int(st16.sum()) == 0
    assert torch.equal(b16.view(torch.int16), batched.roi_crop_bf16(images, rois_b[:total].contiguous(), T=96).view(torch.int16))
    buf = ops.pack_records(idx, n, cost, X, reproj, offs)
    used = 16 + ((24 * 4 + 15) & ~15) + int(n.clamp(min=0).sum()) * 64          # header | counts | valid records
    assert torch.equal(buf[:used], batched.pack_records(a, offs_b, 3)[:used])
    # the filter knob travels through the op as well
    f = ops.match_triangulate(Ks, RTs, cen, cnt, 30, reproj_thresh=1.0)
    fb = batched.match_triangulate(Ks, RTs, cen, cnt, 30, reproj_thresh=1.0)
    assert torch.equal(f[0], fb.idx) and torch.equal(f[1], fb.n)


def test_ops_validate_their_arguments():
    from bpc_baseline_b200 import ops, synth
    batch = synth.make_scenes(2, 4, seed=3)
    Ks, RTs, cen, boxes, cnt = batch_to_dev(batch)
    with pytest.raises(RuntimeError, match='must be Float'):
        ops.match_triangulate(Ks.double(), RTs, cen, cnt)
    with pytest.raises(RuntimeError, match='expected Ks'):
        ops.match_triangulate(Ks[:1].contiguous(), RTs, cen, cnt)
    with pytest.raises(RuntimeError, match='contiguous'):
        ops.box_centers(boxes.transpose(1, 2))
    with pytest.raises(NotImplementedError):
        ops.box_centers(boxes.cpu())                     # no CPU kernel is registered: there is no fallback
    with pytest.raises(RuntimeError, match='target size'):
        ops.roi_crop(to_dev(np.zeros((1, 64, 64, 3), np.uint8)), to_dev(np.zeros((1, 5), np.int32)), 4096)


def test_ops_capture_into_a_cuda_graph_and_trace_with_fake_tensors():
    from bpc_baseline_b200 import ops, synth
    from torch._subclasses.fake_tensor import FakeTensorMode
    batch = synth.make_scenes(8, 6, seed=synth.SEED + 5)
    Ks, RTs, cen, boxes, cnt = batch_to_dev(batch)
    ios = to_dev(np.tile(np.arange(3, dtype=np.int32), (8, 1)))
    images = to_dev(synth.make_images(3, seed=6, width=960, height=540))
    sc = np.array([960 / synth.IMG_W, 540 / synth.IMG_H] * 2)
    b = np.maximum((batch.boxes * sc).astype(np.int32), 0)
    b[..., 2] = np.maximum(b[..., 2], b[..., 0] + 8); b[..., 3] = np.maximum(b[..., 3], b[..., 1] + 8)
    boxes = to_dev(b)
    lut = ops.normalise_lut(images.device)
    out = torch.zeros((8 * 6 * 3, 3, 64, 64), dtype=torch.float32, device='cuda')

    def body():
        ops.box_centers(boxes)
        idx, n, cost, X, reproj, _ = ops.match_triangulate(Ks, RTs, cen, cnt, 30.0)      # centres of the full-resolution boxes
        rois, offs = ops.build_rois(boxes, idx, n, ios)
        ops.roi_crop(images, rois, 64, lut=lut, n_rois=offs[8:9], out=out)
        return idx, n, X, offs
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eager = [t.clone() for t in body()]
        eager_out = out.clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            res = body()
        out.zero_()
        g.replay()
    torch.cuda.synchronize()
    total = int(eager[3][-1])
    assert all(torch.equal(_bits(a), _bits(b)) for a, b in zip(eager, res))
    assert torch.equal(_bits(out[:total]), _bits(eager_out[:total])) and total > 0
    with FakeTensorMode() as mode:                       # Meta kernels: shapes and dtypes without running anything
        f = [mode.from_tensor(t) for t in (Ks, RTs, cen, cnt)]
        r = torch.ops.bpc_b200.match_triangulate(*f, 30.0, None, False)
        assert tuple(r[0].shape) == (8, 6, 3) and r[3].dtype == torch.float64


def test_scene_session_one_graph_replay_equals_the_python_surface_and_the_oracle():
    from bpc_baseline_b200 import scene, synth
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams
    images = synth.make_images(3, seed=12, width=1280, height=720)
    sess = scene.SceneSession(Dmax=12, T=128, image_shape=images.shape)
    sess.set_images(images)
    est = PoseEstimator(PoseEstimatorParams(target_size=128))
    sc = np.array([1280 / synth.IMG_W, 720 / synth.IMG_H] * 2)
    for s, D in enumerate((10, 12, 5, 10)):              # several scenes through the SAME captured graph
        b = synth.make_scenes(1, D, seed=synth.SEED + 300 + s, p_drop=0.2 if s else 0.0)
        Ks, RTs = b.capture_arrays(0)
        # matcher input = centres of the full-resolution boxes; the crop boxes are scaled into the small images
        boxes = [b.boxes[0, c, :b.counts[0, c]] for c in range(3)]
        r = sess.run(Ks, RTs, boxes)
        cen = [b.centers[0, c, :b.counts[0, c]] for c in range(3)]
        want = og.match_scene(Ks, RTs, cen, 30)
        assert r['n'] == len(want['idx']) and np.array_equal(r['idx'].numpy(), want['idx'])
        if r['n']:
            err = np.linalg.norm(r['X'].numpy() - want['X'], axis=1) / np.linalg.norm(want['X'], axis=1)
            assert err.max() < 1e-9
    # crops: a scene whose boxes lie inside the small images, against the Python call surface and the oracle
    b = synth.make_scenes(1, 8, seed=synth.SEED + 400, width=1280, height=720, side_lo=30, side_hi=200)
    Ks, RTs = b.capture_arrays(0)
    boxes = [b.boxes[0, c, :b.counts[0, c]] for c in range(3)]
    r = sess.run(Ks, RTs, boxes)
    cap = SimpleNamespace(images=[images[0], images[1], images[2]], Ks=Ks, RTs=RTs)
    preds = est._match(cap, b.detections(0))
    assert len(preds) == r['n'] and r['n'] > 0 and sess.rejected() == 0
    assert torch.equal(est.crop_inputs(preds), r['crops'])
    m, v = 0, 1
    ref = ocrop.crop_tensor_ref(images[v], boxes[v][int(r['idx'][m, v])], target_size=128)
    assert np.abs(r['crops'][3 * m + v].cpu().numpy() - ref).max() <= 1e-6
    with pytest.raises(RuntimeError, match='exceed Dmax'):
        sess.run(Ks, RTs, [np.zeros((13, 4), np.int32)] * 3)
