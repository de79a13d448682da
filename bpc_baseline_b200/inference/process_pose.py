"""CUDA counterpart of the hot-path half of bpc/inference/process_pose.py.

Same class and method names as the reference: ``PoseEstimatorParams``, ``PosePrediction``,
``PoseEstimator._match`` (process_pose.py:144-188) and ``PoseEstimator._estimate_rotation`` (:190-239), whose
crop / letterbox / normalise part (:199-209) is also exposed as ``PoseEstimator.crop_inputs``.  Detector and
pose network are third-party models outside the hot path: they are optional constructor arguments here
instead of being loaded from checkpoints inside ``__init__`` (:98-111).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .. import _host, batched
from ..utils.data_utils import calc_pose_matrix
from .epipolar_matching import compute_cost_matrix, triangulate_multi_view


@dataclass
class PoseEstimatorParams:
    """Reference process_pose.py:32-38, plus the two knobs of the crop transform."""
    yolo_model_path: str = "yolo11-detection-obj11.pt"
    pose_model_path: str = "best_model.pth"
    matching_threshold: int = 30
    yolo_conf_thresh: float = 0.1
    rotation_mode: str = None
    target_size: int = 256          # process_pose.py:204
    swap_rb: bool = True            # cv2.COLOR_BGR2RGB at :206 (False = the training-dataset order, data_utils.py:252)
    reproj_thresh: float = None     # optional: drop matches whose reprojection error (utils/triangulation.py:14-18)
                                    # exceeds this many pixels in any view; None = the reference (it filters nowhere)


class PosePrediction:
    """Detection triple + capture + triangulated translation -- reference process_pose.py:79-94."""

    def __init__(self, detections, capture, _t=None):
        self.boxes = np.array([x['bbox'] for x in detections])
        self.centroids = np.array([x['bb_center'] for x in detections])
        self.capture = capture
        self.t = self.triangulate() if _t is None else _t

    def triangulate(self):
        proj_mats = projection_matrices(self.capture, len(self.boxes))
        return triangulate_multi_view(proj_mats, self.centroids)


def projection_matrices(capture, n=3):
    """P = K (float32) @ RT[:3] (float64) per camera -- reference process_pose.py:88-92 (on the GPU)."""
    Ks = np.stack([np.asarray(capture.Ks[i]) for i in range(n)])
    RTs = np.stack([np.asarray(capture.RTs[i], np.float64) for i in range(n)])
    if Ks.dtype != np.float32:
        raise TypeError('capture.Ks must be float32 (camera_utils.py:16)')
    P = batched.projection(_host.to_dev(Ks, np.float32), _host.to_dev(RTs, np.float64))
    return [p for p in _host.to_host(P)]


def _scene_tensors(capture, detections):
    Ks = np.stack([np.asarray(k) for k in capture.Ks])
    if Ks.dtype != np.float32:
        raise TypeError('capture.Ks must be float32 (camera_utils.py:16)')
    RTs = np.stack([np.asarray(rt, np.float64) for rt in capture.RTs])
    dets = [detections[0], detections[1], detections[2]]
    D = max(1, max(len(d) for d in dets))
    centers = np.zeros((1, 3, D, 2), np.float64)
    counts = np.zeros((1, 3), np.int32)
    for c, dl in enumerate(dets):
        counts[0, c] = len(dl)
        for d, det in enumerate(dl):
            centers[0, c, d, 0], centers[0, c, d, 1] = det['bb_center'][0], det['bb_center'][1]
    return Ks[None], RTs[None], centers, counts, dets


class PoseEstimator:
    """Reference process_pose.py:98-239 with the hot path on the GPU.

    ``yolo`` / ``pose_model`` are optional already-constructed models; without them ``_detect`` and the
    network half of ``_estimate_rotation`` are unavailable, the hot path (``_match``, ``crop_inputs``) is not.
    """

    def __init__(self, params: PoseEstimatorParams, yolo=None, pose_model=None, rotation_mode=None, verbose=False):
        self.params = params
        self.yolo = yolo
        self.pose_model = pose_model
        self.rotation_mode = rotation_mode if rotation_mode is not None else params.rotation_mode
        self.verbose = verbose

    # ------------------------------------------------------------------------------------------------
    def _detect(self, capture):
        """YOLO on each view -> {cam: [{'bbox', 'bb_center'}]} -- reference process_pose.py:113-142.

        The detector is the caller's model; its post-processing (class / confidence filter, int() truncation, centres,
        :123-141) is one launch of ``bpc_detections_from_yolo`` over all views.  ``detect_tensors`` returns the same
        result as device tensors (the matcher's input layout) without the trip through Python dicts."""
        boxes, centers, counts = self.detect_tensors(capture)
        boxes, centers, counts = _host.to_host(boxes)[0], _host.to_host(centers)[0], _host.to_host(counts)[0]
        camera_predictions = {}
        for idx in range(boxes.shape[0]):
            camera_predictions[idx] = [
                {'bbox': tuple(int(v) for v in boxes[idx, d]), 'bb_center': (float(centers[idx, d, 0]), float(centers[idx, d, 1]))}
                for d in range(int(counts[idx]))]
        return camera_predictions

    def detect_tensors(self, capture):
        """(boxes i32 [1,C,Dmax,4], centers f64 [1,C,Dmax,2], counts i32 [1,C]) on the GPU for one capture."""
        if self.yolo is None:
            raise RuntimeError('no detector attached: pass yolo= to PoseEstimator (third-party model, outside the hot path)')
        dev = _host.device()
        raw = []
        for image in capture.images:
            if self.verbose:
                print(f"Processing image shape: {image.shape}")
            results = self.yolo(image, imgsz=1280)[0]
            b = results.boxes
            raw.append((torch.as_tensor(b.xyxy).to(dev, torch.float32).reshape(-1, 4),
                        torch.as_tensor(b.conf).to(dev, torch.float32).reshape(-1),
                        torch.as_tensor(b.cls).to(dev, torch.float32).reshape(-1)))
        C = len(raw)
        N = max(1, max(int(r[0].shape[0]) for r in raw))
        xyxy = torch.zeros((1, C, N, 4), dtype=torch.float32, device=dev)
        conf = torch.zeros((1, C, N), dtype=torch.float32, device=dev)
        cls = torch.full((1, C, N), -1.0, dtype=torch.float32, device=dev)
        nraw = torch.tensor([[int(r[0].shape[0]) for r in raw]], dtype=torch.int32, device=dev)
        for c, (bx, cf, cl) in enumerate(raw):
            n = int(bx.shape[0])
            xyxy[0, c, :n], conf[0, c, :n], cls[0, c, :n] = bx, cf, cl
        return batched.detections_from_yolo(xyxy, conf, cls, nraw, self.params.yolo_conf_thresh, N)

    # ------------------------------------------------------------------------------------------------
    def _match(self, capture, detections):
        """Epipolar matching of the three views -> list[PosePrediction] sorted by cost.

        Reference process_pose.py:144-188: fundamental matrices, N x M x P cost tensor, SciPy assignment,
        threshold, stable sort by cost, one PosePrediction (with DLT triangulation) per match.  Here: one
        kernel launch (bpc_match_triangulate).  The reference's cost-matrix statistics (:166-179, which also
        consume NumPy's global RNG) are printed only when ``verbose`` is set.
        """
        predictions = []
        Ks, RTs, centers, counts, dets = _scene_tensors(capture, detections)
        if min(len(d) for d in dets) == 0:
            if self.verbose:
                print("\nAt least one camera has zero detections => no matching.")
            return predictions
        res = batched.match_triangulate(_host.to_dev(Ks, np.float32), _host.to_dev(RTs, np.float64),
                                        _host.to_dev(centers, np.float64), _host.to_dev(counts, np.int32),
                                        self.params.matching_threshold, want_reproj=True, want_F=self.verbose,
                                        reproj_thresh=getattr(self.params, 'reproj_thresh', None))
        n = int(res.n.cpu()[0])
        if n < 0:
            batched.check_match_status(res.n)
        idx = _host.to_host(res.idx)[0, :n]
        X = _host.to_host(res.X)[0, :n]
        reproj = _host.to_host(res.reproj)[0, :n]
        cost = _host.to_host(res.cost)[0, :n]
        if self.verbose:
            F = _host.to_host(res.F)[0]
            cm = compute_cost_matrix(dets[0], dets[1], dets[2], F[0], F[1], F[2])
            print("\n--- Cost Matrix Stats ---")
            print(f"Shape: {cm.shape}")
            print(f"Min: {cm.min():.4f}, Max: {cm.max():.4f}, Mean: {cm.mean():.4f}")
        for m in range(n):
            i, j, k = (int(v) for v in idx[m])
            p = PosePrediction([dets[0][i], dets[1][j], dets[2][k]], capture, _t=X[m].copy())
            p.match = (i, j, k)
            p.cost = float(cost[m])
            p.reprojection_error = reproj[m].copy()
            predictions.append(p)
        return predictions

    # ------------------------------------------------------------------------------------------------
    def crop_inputs(self, predictions) -> torch.Tensor:
        """Network inputs of every prediction and view: float32 CUDA tensor [len(predictions) * 3, 3, T, T].

        The crop half of _estimate_rotation, reference process_pose.py:199-209: box crop, letterbox
        (INTER_AREA, white canvas), BGR->RGB, /255, ImageNet normalisation -- one batched launch instead of a
        Python loop with one H2D copy per crop (:210).  Row 3*p + v is prediction p, camera v.
        """
        T = int(self.params.target_size)
        if not predictions:
            return torch.empty((0, 3, T, T), dtype=torch.float32, device=_host.device())
        # the distinct images referenced by the predictions form the image pool
        pool, rois = {}, []
        for p in predictions:
            for v in range(len(p.boxes)):
                img = p.capture.images[v]
                key = id(img)
                if key not in pool:
                    pool[key] = (len(pool), img)
                x1, y1, x2, y2 = (int(t) for t in p.boxes[v])
                if x2 <= x1 or y2 <= y1:
                    raise ZeroDivisionError('float division by zero')       # empty crop: data_utils.py:36
                rois.append((pool[key][0], x1, y1, x2, y2))
        imgs = [np.asarray(img) for _, img in sorted(pool.values(), key=lambda t: t[0])]
        shapes = {im.shape for im in imgs}
        if len(shapes) != 1:
            raise ValueError('all views must have the same resolution')
        H, W = imgs[0].shape[:2]
        for (_, x1, y1, x2, y2) in rois:
            if x1 < 0 or y1 < 0 or x2 > W or y2 > H:
                raise ValueError('box outside the image (the reference would wrap or clip the slice)')
        images = _host.to_dev(np.stack(imgs), np.uint8)
        status = torch.zeros((len(rois),), dtype=torch.int32, device=images.device)
        out = batched.roi_crop(images, _host.to_dev(np.asarray(rois, np.int32), np.int32), T=T,
                               swap_rb=self.params.swap_rb, status=status)
        if int(status.sum()) != 0:
            raise ValueError('a box resizes to an empty image (cv2.resize would fail)')
        return out

    def _estimate_rotation(self, predictions):
        """Crops -> pose network -> rotation matrices -- reference process_pose.py:190-239.

        The crops are produced by ``crop_inputs``; the network runs once on the whole batch instead of once per
        crop.  Sets ``rotation_preds``, ``final_rotation`` (third camera, :238) and ``pose`` on every prediction.
        """
        tens = self.crop_inputs(predictions)
        if self.pose_model is None:
            raise RuntimeError('no pose network attached: pass pose_model= to PoseEstimator; '
                               'crop_inputs() returns the network inputs on their own')
        with torch.no_grad():
            raw = self.pose_model(tens).float().cpu()
        rot = decode_rotations(raw, self.rotation_mode)
        r = 0
        for p in predictions:
            p.rotation_preds = []
            for v in range(len(p.centroids)):
                p.rotation_preds.append(np.asarray(p.capture.RTs[v])[:3, :3].T @ rot[r])
                r += 1
            p.final_rotation = p.rotation_preds[2]
            p.pose = calc_pose_matrix(p.final_rotation, p.t)


def decode_rotations(raw: torch.Tensor, mode=None) -> np.ndarray:
    """Raw network outputs [n, 3 | 4 | 6] (CPU float32) -> rotation matrices [n, 3, 3] float32, batched form of the
    per-crop decode in reference process_pose.py:213-229: euler angles are wrapped to [-pi, pi) in NumPy first (:215),
    quaternions are normalised with eps 1e-8 (:220-222).  ``mode`` None picks the head from the output width."""
    if mode is None:
        mode = {3: 'euler', 4: 'quat', 6: '6d'}.get(raw.shape[1])
    if mode == 'euler':
        wrapped = ((raw.numpy() + np.pi) % (2 * np.pi)) - np.pi
        rot = rotmat_from_euler(torch.tensor(wrapped, dtype=torch.float32))
    elif mode == 'quat':
        rot = quat_to_rotmat(raw / (raw.norm(dim=1, keepdim=True) + 1e-8))
    elif mode == '6d':
        rot = rotmat_from_6d(raw)
    else:
        raise ValueError("Unsupported rotation mode.")
    return rot.numpy()


# ---- rotation decoders of the pose head (reference bpc/pose/models/losses.py:26-84; off the hot path) ----
def rotmat_from_euler(e):
    cx, sx = torch.cos(e[:, 0]), torch.sin(e[:, 0])
    cy, sy = torch.cos(e[:, 1]), torch.sin(e[:, 1])
    cz, sz = torch.cos(e[:, 2]), torch.sin(e[:, 2])
    rows = [torch.stack([cy * cz, -cy * sz, sy], dim=1),
            torch.stack([sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy], dim=1),
            torch.stack([-cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy], dim=1)]
    return torch.stack(rows, dim=1)


def quat_to_rotmat(quat):
    quat = quat / quat.norm(dim=1, keepdim=True)
    x, y, z, w = quat.unbind(dim=1)
    return torch.stack([1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
                        2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w,
                        2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y], dim=1).view(-1, 3, 3)


def rotmat_from_6d(rep6d):
    a1, a2 = rep6d[:, 0:3], rep6d[:, 3:6]
    b1 = torch.nn.functional.normalize(a1, dim=1, eps=1e-8)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(dim=1, keepdim=True) * b1, dim=1, eps=1e-8)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=2)
