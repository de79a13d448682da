"""CUDA counterpart of the hot-path helpers of bpc/utils/data_utils.py (same names and arguments)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from .. import _host, batched
from ..inference.utils.camera_utils import load_camera_params


def _resize_error(msg):
    try:
        import cv2
        return cv2.error(msg)
    except Exception:                                    # cv2 is not a dependency of this package
        return ValueError(msg)


def letterbox_preserving_aspect_ratio(img, target_size=256, fill_color=(255, 255, 255)):
    """Aspect-preserving INTER_AREA resize onto a fill_color canvas -- reference data_utils.py:34-44.

    Returns (canvas uint8 [T, T, 3], scale, dx, dy) exactly as the reference; the resize runs on the GPU
    and is bit-identical to cv2.resize(..., interpolation=cv2.INTER_AREA).
    """
    img = np.asarray(img)
    h, w = img.shape[:2]
    scale = float(target_size) / max(h, w)                # ZeroDivisionError on an empty crop, as the reference
    new_w = int(round(w * scale))
    new_h = int(round(h * scale))
    dx = (target_size - new_w) // 2
    dy = (target_size - new_h) // 2
    if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
        raise TypeError('img must be an H x W x 3 uint8 array')
    if new_w < 1 or new_h < 1:
        raise _resize_error("OpenCV(-215:Assertion failed) !dsize.empty() in function 'resize'")
    if not 1 <= int(target_size) <= 256:
        raise NotImplementedError('target_size above 256 is not supported by the CUDA crop kernels')
    images = _host.to_dev(img[None], np.uint8)
    rois = _host.to_dev([[0, 0, 0, w, h]], np.int32)
    canvas = batched.roi_crop_u8(images, rois, T=int(target_size), fill=fill_color)
    return _host.to_host(canvas)[0].copy(), scale, dx, dy


def calc_pose_matrix(R_mat, t):
    """4x4 float64 pose from R, t -- reference data_utils.py:383-387 (a container, no arithmetic)."""
    pose = np.eye(4)
    pose[:3, :3] = R_mat
    pose[:3, 3] = t
    return pose


def _read_json(path):
    with open(path) as fh:
        return json.load(fh)


def load_gt_poses(scene_dir, scene_id, cam_ids, image_id, obj_id):
    """4x4 ground-truth poses of object ``obj_id`` in image ``image_id`` of the FIRST camera only, or [] when the
    scene_gt / scene_gt_info files or the image entry are missing -- behaviour of reference data_utils.py:355-380."""
    cam = list(cam_ids)[0] if len(cam_ids) else None
    if cam is None:
        return []
    base = os.path.join(scene_dir, scene_id)
    gt_file, info_file = (os.path.join(base, f"{stem}_{cam}.json") for stem in ("scene_gt", "scene_gt_info"))
    if not (os.path.exists(gt_file) and os.path.exists(info_file)):
        return []
    gt, info = _read_json(gt_file), _read_json(info_file)
    key = str(image_id)
    if key not in gt or key not in info:
        return []
    entries = gt[key][:len(info[key])]                      # the reference zips the two lists
    return [calc_pose_matrix(np.asarray(e["cam_R_m2c"], dtype=np.float32).reshape(3, 3),
                             np.asarray(e["cam_t_m2c"], dtype=np.float32))
            for e in entries if e["obj_id"] == obj_id]


class Capture:
    """Images + intrinsics + extrinsics of one multi-camera capture -- reference data_utils.py:390-409."""

    def __init__(self, images, Ks, RTs, obj_id, gt_poses=None):
        self.images = images
        self.Ks = Ks
        self.RTs = RTs
        self.obj_id = obj_id
        if gt_poses:
            self.gt_poses = np.linalg.inv(RTs[0]) @ gt_poses

    @classmethod
    def from_dir(cls, scene_dir, cam_ids, image_id, obj_id):
        """Read one capture from a BOP scene directory (reference data_utils.py:399-409): intrinsics / extrinsics from
        scene_camera_<cam>.json, the image ``rgb_<cam>/<image_id:06d>.{png,jpg}`` per camera, RT as float64 4x4."""
        import cv2                                        # image decoding only (host I/O)
        cams = load_camera_params(scene_dir, cam_ids)
        Ks = [cams[c]['K'][image_id] for c in cam_ids]
        RTs = [calc_pose_matrix(cams[c]['R'][image_id], cams[c]['t'][image_id]) for c in cam_ids]
        images = []
        for c in cam_ids:
            hits = glob.glob(os.path.join(scene_dir, f"rgb_{c}", f"{image_id:06d}.*g"))
            images.append(cv2.imread(hits[0]))
        return cls(images, Ks, RTs, obj_id, load_gt_poses(scene_dir, '', cam_ids, image_id, obj_id))
