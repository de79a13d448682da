"""CUDA counterpart of bpc/inference/utils/camera_utils.py (same names and arguments)."""
from __future__ import annotations

import json
import os

import numpy as np

from ... import _host, batched

_FIELDS = (('K', 'cam_K', (3, 3)), ('R', 'cam_R_w2c', (3, 3)), ('t', 'cam_t_w2c', (-1,)))


def load_camera_params(scene_dir, cam_ids):
    """BOP ``scene_camera_<cam>.json`` -> {cam: {'K' | 'R' | 't': {image_id: float32 array}}}.

    Same structure and dtypes as the reference loader (camera_utils.py:6-20): K and R are 3x3, t is flat, all
    float32 -- which is what makes R and t reach the matcher as float32-rounded values.  Host I/O only.
    """
    params = {}
    for cam in cam_ids:
        with open(os.path.join(scene_dir, f"scene_camera_{cam}.json")) as fh:
            per_image = json.load(fh)
        params[cam] = {name: {int(image_id): np.asarray(entry[key], dtype=np.float32).reshape(shape)
                              for image_id, entry in per_image.items()}
                       for name, key, shape in _FIELDS}
    return params


def compute_fundamental_matrix(K1, R1, t1, K2, R2, t2):
    """Fundamental matrix between two cameras -- reference camera_utils.py:23-46, on the GPU.

    The arithmetic follows the dtype flow of PoseEstimator._match (process_pose.py:154-159): K float32
    (its inverse is taken in float32), R / t float64 holding float32-rounded values.  K must be float32;
    R and t are widened to float64 if they are not already.  Returns a float64 (3, 3) array.
    """
    K1 = np.asarray(K1); K2 = np.asarray(K2)
    if K1.dtype != np.float32 or K2.dtype != np.float32:
        raise TypeError('K1 and K2 must be float32 (camera_utils.py:16 loads them as float32)')
    Ks = np.stack([K1.reshape(3, 3), K2.reshape(3, 3), K2.reshape(3, 3)])[None]
    RTs = np.zeros((1, 3, 4, 4), np.float64)
    for c, (R, t) in enumerate(((R1, t1), (R2, t2), (R2, t2))):
        RTs[0, c, :3, :3] = np.asarray(R, np.float64).reshape(3, 3)
        RTs[0, c, :3, 3] = np.asarray(t, np.float64).flatten()
        RTs[0, c, 3, 3] = 1.0
    F = batched.fundamental(_host.to_dev(Ks, np.float32), _host.to_dev(RTs, np.float64))
    return _host.to_host(F)[0, 0].copy()
