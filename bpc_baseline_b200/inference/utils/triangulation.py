"""CUDA counterpart of bpc/inference/utils/triangulation.py (same names and arguments)."""
from __future__ import annotations

import numpy as np

from ... import _host, batched


def triangulate_multi_view(proj_mats, points):
    """DLT over all views -- reference utils/triangulation.py:3-12 (= epipolar_matching.py:118-127)."""
    P = np.stack([np.asarray(p, np.float64).reshape(3, 4) for p in proj_mats])[None]
    pts = np.asarray([[float(x), float(y)] for (x, y) in points], np.float64)[None]
    if P.shape[1] != pts.shape[1]:
        V = min(P.shape[1], pts.shape[1])        # zip() semantics of the reference
        P, pts = P[:, :V], pts[:, :V]
    X = batched.triangulate_views(_host.to_dev(P, np.float64), _host.to_dev(pts, np.float64))
    return _host.to_host(X)[0].copy()


def compute_reprojection_error(P, X, point_2d):
    """Pixel reprojection error of one view -- reference utils/triangulation.py:14-18."""
    Pm = np.asarray(P, np.float64).reshape(1, 1, 3, 4).repeat(3, axis=1)
    pts = np.asarray(point_2d, np.float64).reshape(1, 1, 2).repeat(3, axis=1)
    err = batched.reprojection_error(_host.to_dev(Pm, np.float64), _host.to_dev(np.asarray(X, np.float64).reshape(1, 3), np.float64),
                                     _host.to_dev(pts, np.float64))
    return float(_host.to_host(err)[0, 0])


def compute_final_pose(final_pose_array, triangulated_points):
    """(N, 6) float32 rows [mean Rx, Ry, Rz over the cameras | X, Y, Z] -- reference utils/triangulation.py:20-45
    (host glue with no caller in the reference; off the hot path)."""
    angles = np.asarray(final_pose_array)[:, :, :3]
    pose = np.empty((angles.shape[0], 6), dtype=np.float32)
    pose[:, :3] = angles.mean(axis=1)
    pose[:, 3:] = np.asarray(triangulated_points)[:angles.shape[0]]
    return pose
