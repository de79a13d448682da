"""CUDA counterpart of bpc/inference/epipolar_matching.py (same names, arguments and return types)."""
from __future__ import annotations

import numpy as np

from .. import _host, batched
from .utils.triangulation import triangulate_multi_view as _triangulate_multi_view


def epipolar_error(pt1, pt2, F, img1=None, img2=None):
    """Symmetric epipolar distance (float) -- reference epipolar_matching.py:5-71.

    The reference's visualisation branch uses ``plt`` without importing matplotlib (:53) and therefore
    raises NameError whenever both images are given; that behaviour is kept rather than silently fixed.
    """
    e = batched.epipolar_error(_host.to_dev(np.asarray(F, np.float64).reshape(1, 3, 3), np.float64),
                               _host.to_dev([[pt1[0], pt1[1]]], np.float64), _host.to_dev([[pt2[0], pt2[1]]], np.float64))
    error = float(_host.to_host(e)[0])
    if img1 is not None and img2 is not None:
        raise NameError("name 'plt' is not defined")
    return error


def epipolar_error_full(pt1, pt2, pt3, F12, F13, F23):
    """(e12 + e13 + e23) / 3 in float64 -- reference epipolar_matching.py:73-81."""
    F = np.stack([np.asarray(F12, np.float64), np.asarray(F13, np.float64), np.asarray(F23, np.float64)]).reshape(1, 3, 3, 3)
    pts = np.asarray([[pt1[0], pt1[1]], [pt2[0], pt2[1]], [pt3[0], pt3[1]]], np.float64).reshape(1, 3, 2)
    e = batched.epipolar_error_full(_host.to_dev(F, np.float64), _host.to_dev(pts, np.float64))
    return float(_host.to_host(e)[0])


def compute_cost_matrix(dets1, dets2, dets3, F12, F13, F23, img1=None, img2=None, img3=None):
    """N x M x P float32 cost tensor from the 'bb_center' of each detection -- reference :83-98.

    ``img1..3`` are accepted and ignored, as in the reference.
    """
    N, M, P = len(dets1), len(dets2), len(dets3)
    if N == 0 or M == 0 or P == 0:
        return np.zeros((N, M, P), dtype=np.float32)
    D = max(N, M, P)
    centers = np.zeros((1, 3, D, 2), np.float64)
    for c, dets in enumerate((dets1, dets2, dets3)):
        for d, det in enumerate(dets):
            centers[0, c, d, 0], centers[0, c, d, 1] = det['bb_center'][0], det['bb_center'][1]
    F = np.stack([np.asarray(F12, np.float64), np.asarray(F13, np.float64), np.asarray(F23, np.float64)]).reshape(1, 3, 3, 3)
    cost = batched.cost_tensor(_host.to_dev(F, np.float64), _host.to_dev(centers, np.float64),
                               _host.to_dev([[N, M, P]], np.int32))
    return _host.to_host(cost[0, :N, :M, :P]).copy()


def match_objects(cost_matrix, threshold):
    """Flatten -> rectangular assignment (SciPy-exact) -> keep cost < threshold -> [(i, j, k)], ascending
    r = i*M + j -- reference epipolar_matching.py:100-116."""
    cost_matrix = np.asarray(cost_matrix)
    N, M, P = cost_matrix.shape
    if N * M == 0 or P == 0:
        return []
    cost = _host.to_dev(cost_matrix.reshape(1, N, M, P), np.float32)
    idx, n = batched.match_objects(cost, threshold)
    n = int(n.cpu()[0])
    if n < 0:
        raise ValueError('matrix contains invalid numeric entries')     # what scipy.optimize.linear_sum_assignment raises
    rows = _host.to_host(idx)[0, :n].astype(np.int64)
    return [(r[0], r[1], r[2]) for r in rows]


def triangulate_multi_view(proj_mats, points_2D):
    """Triangulate using the Direct Linear Transform -- reference epipolar_matching.py:118-127."""
    return _triangulate_multi_view(proj_mats, points_2D)
