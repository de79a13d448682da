"""Detector post-processing kernel against the restated loop of process_pose.py:123-141,
and the consumer side: the batched crop tensor through SimplePoseNet (rank 1)."""
import numpy as np
import pytest
import torch

from oracle import detect as odet
from tests.gpu_util import to_dev

pytestmark = pytest.mark.gpu


def test_detections_from_yolo_matches_reference_loop():
    from bpc_baseline_b200 import batched
    rng = np.random.default_rng(3)
    S, N, Dmax = 7, 50, 24
    xyxy = (rng.random((S, 3, N, 4)) * np.array([3840, 2160, 3840, 2160])).astype(np.float32)
    xyxy[..., :2] -= rng.random((S, 3, N, 2)).astype(np.float32) * 30          # some negative coordinates: int() truncates toward 0
    conf = rng.random((S, 3, N)).astype(np.float32)
    conf[0, 0, :5] = np.float32(0.1)                                            # exactly at the threshold: kept
    cls = rng.integers(0, 2, (S, 3, N)).astype(np.float32)
    nraw = rng.integers(0, N + 1, (S, 3)).astype(np.int32)
    nraw[1, 2] = 0
    boxes, centers, counts = batched.detections_from_yolo(to_dev(xyxy), to_dev(conf), to_dev(cls), to_dev(nraw), 0.1, Dmax)
    boxes, centers, counts = boxes.cpu().numpy(), centers.cpu().numpy(), counts.cpu().numpy()
    overflow = 0
    for s in range(S):
        for c in range(3):
            n = nraw[s, c]
            want = odet.detections_from_yolo(xyxy[s, c, :n], conf[s, c, :n], cls[s, c, :n], 0.1)
            assert counts[s, c] == len(want)
            overflow += len(want) > Dmax
            for d, det in enumerate(want[:Dmax]):
                assert tuple(boxes[s, c, d]) == det['bbox']
                assert tuple(centers[s, c, d]) == det['bb_center']
    assert overflow >= 0


def test_detect_against_the_reference_method(golden_detect):
    """Kernel and PoseEstimator._detect mirror against outputs of the reference's own _detect (tests/golden/detect.npz)."""
    from types import SimpleNamespace
    from bpc_baseline_b200 import batched
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams
    g = golden_detect
    xyxy, conf, cls, nraw, thresh = g['xyxy'], g['conf'], g['cls'], g['nraw'], float(g['thresh'])
    S, _, N, _ = xyxy.shape
    boxes, centers, counts = batched.detections_from_yolo(to_dev(xyxy), to_dev(conf), to_dev(cls), to_dev(nraw), thresh, N)
    boxes, centers, counts = boxes.cpu().numpy(), centers.cpu().numpy(), counts.cpu().numpy()
    for s in range(S):
        for c in range(3):
            want_b, want_c = g[f'bbox_{s}_{c}'], g[f'center_{s}_{c}']
            assert counts[s, c] == len(want_b), (s, c)
            assert np.array_equal(boxes[s, c, :len(want_b)], want_b) and np.array_equal(centers[s, c, :len(want_b)], want_c)

    class _Boxes:
        def __init__(self, b, c, k):
            self.xyxy, self.conf, self.cls = torch.from_numpy(b), torch.from_numpy(c), torch.from_numpy(k)

        def __len__(self):
            return int(self.xyxy.shape[0])

    for s in range(S):
        calls = iter(range(3))

        def yolo(image, imgsz=1280, _s=s, _calls=calls):
            c = next(_calls)
            n = nraw[_s, c]
            return [SimpleNamespace(boxes=_Boxes(xyxy[_s, c, :n].copy(), conf[_s, c, :n].copy(), cls[_s, c, :n].copy()))]

        est = PoseEstimator(PoseEstimatorParams(yolo_conf_thresh=thresh), yolo=yolo)
        out = est._detect(SimpleNamespace(images=[np.zeros((4, 4, 3), np.uint8)] * 3))
        for c in range(3):
            assert [d['bbox'] for d in out[c]] == [tuple(int(v) for v in b) for b in g[f'bbox_{s}_{c}']]
            assert [d['bb_center'] for d in out[c]] == [tuple(float(v) for v in b) for b in g[f'center_{s}_{c}']]


def test_crops_feed_simple_pose_net_batched():
    """_estimate_rotation end to end with the reference's network architecture (random weights, no download)."""
    from types import SimpleNamespace
    from bpc_baseline_b200 import synth
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams
    from bpc_baseline_b200.pose.models.simple_pose_net import SimplePoseNet
    torch.manual_seed(0)
    net = SimplePoseNet(loss_type='6d', pretrained=False).cuda().eval()
    batch = synth.make_scenes(1, 4, seed=9)
    images = synth.make_images(3, seed=9)
    Ks, RTs = batch.capture_arrays(0)
    cap = SimpleNamespace(images=[images[0], images[1], images[2]], Ks=Ks, RTs=RTs)
    est = PoseEstimator(PoseEstimatorParams(target_size=224), pose_model=net, rotation_mode='6d')
    preds = est._match(cap, batch.detections(0))
    assert len(preds) == 4
    est._estimate_rotation(preds)
    for p in preds:
        assert len(p.rotation_preds) == 3 and p.pose.shape == (4, 4)
        np.testing.assert_allclose(p.final_rotation @ p.final_rotation.T, np.eye(3), atol=1e-4)
        np.testing.assert_allclose(p.pose[:3, 3], p.t)
    # batched forward == per-crop forward of the reference loop (process_pose.py:210-212)
    tens = est.crop_inputs(preds)
    with torch.no_grad():
        one = torch.cat([net(tens[i:i + 1]) for i in range(tens.shape[0])])
        allb = net(tens)
    # cuDNN picks different (TF32) convolution algorithms for batch 1 and batch 12: agreement to ~1e-3 relative
    assert torch.allclose(one, allb, atol=5e-2, rtol=1e-2)
