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


def load_gt_poses(scene_dir, scene_id, cam_ids, image_id, obj_id):
    """Ground-truth poses of one object in the first camera -- reference data_utils.py:355-380 (host I/O)."""
    gt_poses = []
    scene_path = os.path.join(scene_dir, scene_id)
    for cam_id in cam_ids[:1]:
        gt_path = os.path.join(scene_path, f"scene_gt_{cam_id}.json")
        info_path = os.path.join(scene_path, f"scene_gt_info_{cam_id}.json")
        if not os.path.exists(gt_path) or not os.path.exists(info_path):
            continue
        with open(gt_path, "r") as f:
            gt_data = json.load(f)
        with open(info_path, "r") as f:
            info_data = json.load(f)
        img_key = str(image_id)
        if img_key not in gt_data or img_key not in info_data:
            continue
        for obj, _bbox in zip(gt_data[img_key], info_data[img_key]):
            if obj["obj_id"] != obj_id:
                continue
            rotation_matrix = np.array(obj["cam_R_m2c"], dtype=np.float32).reshape(3, 3)
            translation = np.array(obj["cam_t_m2c"], dtype=np.float32)
            gt_poses.append(calc_pose_matrix(rotation_matrix, translation))
    return gt_poses


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
        import cv2                                        # image decoding only (host I/O)
        cam_params = load_camera_params(scene_dir, cam_ids)
        Ks = [cam_params[x]['K'][image_id] for x in cam_ids]
        Rs = [cam_params[x]['R'][image_id] for x in cam_ids]
        Ts = [cam_params[x]['t'][image_id] for x in cam_ids]
        RTs = [calc_pose_matrix(r, t) for r, t in zip(Rs, Ts)]
        image_paths = [glob.glob(os.path.join(scene_dir, f"rgb_{cam_id}", f"{image_id:06d}.*g"))[0] for cam_id in cam_ids]
        images = [cv2.imread(x) for x in image_paths]
        gt_poses = load_gt_poses(scene_dir, '', cam_ids, image_id, obj_id)
        return cls(images, Ks, RTs, obj_id, gt_poses)
