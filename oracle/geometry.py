"""Oracle (test infrastructure): NumPy restatement of the reference geometry path.

Every function follows the cited reference lines operation for operation, using the same
NumPy calls (``@``, ``np.dot``, ``np.linalg.norm/inv/svd``) so that on one machine it is
bit-identical to the reference; ``tests/test_oracle_golden.py`` pins that against outputs of
the reference itself.  Not importable from the product package.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment


def calc_pose_matrix(R_mat, t):
    """4x4 float64 pose from R, t -- reference bpc/utils/data_utils.py:383-387."""
    pose = np.eye(4)
    pose[:3, :3] = R_mat
    pose[:3, 3] = t
    return pose


def fundamental_matrix(K1, R1, t1, K2, R2, t2):
    """F (f64 3x3) mapping a cam-1 point to its cam-2 epipolar line.

    Reference bpc/inference/utils/camera_utils.py:23-46: R_rel :27, t_rel :28, [t]x cast to
    float32 :31-35, E :37, float32 inverses :38-39, F :40, normalise by F[2,2] :43-44.
    """
    t1 = t1.flatten()
    t2 = t2.flatten()
    R_rel = R2 @ R1.T
    t_rel = t2 - R_rel @ t1
    tx = np.array([[0, -t_rel[2], t_rel[1]],
                   [t_rel[2], 0, -t_rel[0]],
                   [-t_rel[1], t_rel[0], 0]], dtype=np.float32)
    E = tx @ R_rel
    K1_inv = np.linalg.inv(K1)
    K2_inv = np.linalg.inv(K2)
    F = K2_inv.T @ E @ K1_inv
    if abs(F[2, 2]) > 1e-8:
        F /= F[2, 2]
    return F


def scene_fundamentals(Ks, RTs):
    """(F12, F13, F23) exactly as PoseEstimator._match builds them (process_pose.py:154-159)."""
    K1, K2, K3 = Ks
    R1, R2, R3 = [x[:3, :3] for x in RTs]
    t1, t2, t3 = [x[:3, 3] for x in RTs]
    return (fundamental_matrix(K1, R1, t1, K2, R2, t2),
            fundamental_matrix(K1, R1, t1, K3, R3, t3),
            fundamental_matrix(K2, R2, t2, K3, R3, t3))


def epipolar_error(pt1, pt2, F):
    """Symmetric epipolar distance -- reference epipolar_matching.py:5-28 (viz branch dropped)."""
    pt1_h = np.array([pt1[0], pt1[1], 1.0])
    pt2_h = np.array([pt2[0], pt2[1], 1.0])
    l2 = F @ pt1_h
    l1 = F.T @ pt2_h
    norm_l1 = np.linalg.norm(l1[:2])
    norm_l2 = np.linalg.norm(l2[:2])
    if norm_l1 > 1e-8:
        l1 /= norm_l1
    if norm_l2 > 1e-8:
        l2 /= norm_l2
    d1 = abs(np.dot(l1, pt1_h)) if norm_l1 > 1e-8 else 9999
    d2 = abs(np.dot(l2, pt2_h)) if norm_l2 > 1e-8 else 9999
    return 0.5 * (d1 + d2)


def epipolar_error_full(pt1, pt2, pt3, F12, F13, F23):
    """(e12 + e13 + e23) / 3 in f64 -- reference epipolar_matching.py:73-81."""
    e12 = epipolar_error(pt1, pt2, F12)
    e13 = epipolar_error(pt1, pt3, F13)
    e23 = epipolar_error(pt2, pt3, F23)
    return (e12 + e13 + e23) / 3


def cost_tensor_loop(c1, c2, c3, F12, F13, F23):
    """The reference's triple loop, as written -- epipolar_matching.py:83-98.  O(N*M*P) Python."""
    N, M, P = len(c1), len(c2), len(c3)
    cost = np.zeros((N, M, P), dtype=np.float32)
    for i in range(N):
        for j in range(M):
            for k in range(P):
                cost[i, j, k] = epipolar_error_full(c1[i], c2[j], c3[k], F12, F13, F23)
    return cost


def _pair_matrix(pa, pb, F):
    """e[i, j] = epipolar_error(pa[i], pb[j], F), the same scalar ops hoisted out of the loops.

    The normalised line l2 depends only on pa[i] and l1 only on pb[j]
    (epipolar_matching.py:13-23), so they are computed once per point with the reference's own
    calls; the two 3-term dot products per pair (:25-26) stay ``np.dot`` calls so the rounding
    (BLAS ddot) is the reference's.
    """
    na, nb = len(pa), len(pb)
    pah = [np.array([p[0], p[1], 1.0]) for p in pa]
    pbh = [np.array([p[0], p[1], 1.0]) for p in pb]
    l2s, ok2 = [], []
    for i in range(na):
        l2 = F @ pah[i]
        n2 = np.linalg.norm(l2[:2])
        if n2 > 1e-8:
            l2 /= n2
        l2s.append(l2)
        ok2.append(n2 > 1e-8)
    l1s, ok1 = [], []
    for j in range(nb):
        l1 = F.T @ pbh[j]
        n1 = np.linalg.norm(l1[:2])
        if n1 > 1e-8:
            l1 /= n1
        l1s.append(l1)
        ok1.append(n1 > 1e-8)
    e = np.empty((na, nb), np.float64)
    for i in range(na):
        for j in range(nb):
            d1 = abs(np.dot(l1s[j], pah[i])) if ok1[j] else 9999
            d2 = abs(np.dot(l2s[i], pbh[j])) if ok2[i] else 9999
            e[i, j] = 0.5 * (d1 + d2)
    return e


def pair_matrices(c1, c2, c3, F12, F13, F23):
    return _pair_matrix(c1, c2, F12), _pair_matrix(c1, c3, F13), _pair_matrix(c2, c3, F23)


def cost_tensor(c1, c2, c3, F12, F13, F23):
    """Separable restatement of compute_cost_matrix: bit-identical to the loop.

    cost[i,j,k] = f32(((e12[i,j] + e13[i,k]) + e23[j,k]) / 3), f64 until the final store
    (epipolar_matching.py:81,88,96).
    """
    e12, e13, e23 = pair_matrices(c1, c2, c3, F12, F13, F23)
    s = (e12[:, :, None] + e13[:, None, :]) + e23[None, :, :]
    return (s / 3).astype(np.float32)


def _pair_matrix_fast(pa, pb, F):
    """_pair_matrix with the D*D dot products batched into two matrix products.

    Lines still come from the reference's per-point calls; only ``np.dot(l, p)`` per pair becomes a
    row of a (D,3)@(3,D) product.  On the hosts tried the BLAS gemm accumulates the three products in
    the same fused order as ddot, so the result is bit-identical; ``cost_tensor_fast`` verifies that
    on its first calls and falls back to ``_pair_matrix`` if it ever is not.
    """
    pah = np.concatenate([np.asarray(pa, np.float64), np.ones((len(pa), 1))], axis=1)
    pbh = np.concatenate([np.asarray(pb, np.float64), np.ones((len(pb), 1))], axis=1)
    L2 = np.stack([F @ p for p in pah])
    L1 = np.stack([F.T @ p for p in pbh])
    n2 = np.array([np.linalg.norm(l[:2]) for l in L2])
    n1 = np.array([np.linalg.norm(l[:2]) for l in L1])
    ok2, ok1 = n2 > 1e-8, n1 > 1e-8
    L2 = np.where(ok2[:, None], L2 / np.where(ok2, n2, 1.0)[:, None], L2)
    L1 = np.where(ok1[:, None], L1 / np.where(ok1, n1, 1.0)[:, None], L1)
    d1 = np.where(ok1[None, :], np.abs(pah @ L1.T), 9999.0)
    d2 = np.where(ok2[:, None], np.abs(L2 @ pbh.T), 9999.0)
    return 0.5 * (d1 + d2)


_FAST_STATE = {'checked': 0, 'ok': True}


def cost_tensor_fast(c1, c2, c3, F12, F13, F23):
    """cost_tensor at NumPy speed (the CPU-baseline workhorse); self-checking, see _pair_matrix_fast."""
    fn = _pair_matrix_fast if _FAST_STATE['ok'] else _pair_matrix
    e12, e13, e23 = fn(c1, c2, F12), fn(c1, c3, F13), fn(c2, c3, F23)
    if _FAST_STATE['ok'] and _FAST_STATE['checked'] < 3:
        _FAST_STATE['checked'] += 1
        if not np.array_equal(e12, _pair_matrix(c1, c2, F12)):
            _FAST_STATE['ok'] = False
            return cost_tensor(c1, c2, c3, F12, F13, F23)
    s = (e12[:, :, None] + e13[:, None, :]) + e23[None, :, :]
    return (s / 3).astype(np.float32)


def match_objects(cost_matrix, threshold):
    """Flatten -> SciPy LSAP -> keep cost < threshold -- reference epipolar_matching.py:100-116."""
    N, M, P = cost_matrix.shape
    matched = []
    flattened = cost_matrix.reshape(N * M, P)
    row_idx, col_idx = linear_sum_assignment(flattened)
    for r, c in zip(row_idx, col_idx):
        val = flattened[r, c]
        if val < threshold:
            matched.append((int(r // M), int(r % M), int(c)))
    return matched


def triangulate_multi_view(proj_mats, points_2D):
    """DLT via the last right singular vector -- reference epipolar_matching.py:118-127."""
    A = []
    for P, (x, y) in zip(proj_mats, points_2D):
        A.append(x * P[2] - P[0])
        A.append(y * P[2] - P[1])
    A = np.array(A)
    _, _, Vt = np.linalg.svd(A)
    X = Vt[-1]
    return X[:3] / X[3]


def reprojection_error(P, X, point_2d):
    """Pixel reprojection error -- reference bpc/inference/utils/triangulation.py:14-18."""
    proj = P @ np.append(X, 1.0)
    proj /= proj[2]
    return np.linalg.norm(proj[:2] - point_2d)


def projection_matrices(Ks, RTs):
    """P_c = K_c (f32) @ RT_c[:3] (f64) -> f64 3x4 -- reference process_pose.py:88-92."""
    return [K @ RT[:3] for K, RT in zip(Ks, RTs)]


def match_scene(Ks, RTs, centers, threshold=30, cost_fn=cost_tensor):
    """Restatement of PoseEstimator._match (process_pose.py:144-188) on bare arrays.

    ``centers`` = three arrays [n_c, 2] of bb_center values.  Returns a dict with
      idx    int64 [n, 3]  matched (i, j, k), sorted by cost (stable; ties keep ascending r) :183
      cost   f32   [n]     cost_matrix[i, j, k] of each match
      X      f64   [n, 3]  triangulated points (PosePrediction.t, process_pose.py:84-94)
      reproj f64   [n, 3]  per-view reprojection error (triangulation.py:14-18; unused by _match)
      F      f64   [3,3,3] F12, F13, F23
    Zero detections in any view -> empty result (process_pose.py:161-163).
    """
    F12, F13, F23 = scene_fundamentals(Ks, RTs)
    out = {'F': np.stack([F12, F13, F23]),
           'idx': np.zeros((0, 3), np.int64), 'cost': np.zeros(0, np.float32),
           'X': np.zeros((0, 3)), 'reproj': np.zeros((0, 3))}
    c1, c2, c3 = centers
    if len(c1) == 0 or len(c2) == 0 or len(c3) == 0:
        return out
    cost = cost_fn(c1, c2, c3, F12, F13, F23)
    matches = match_objects(cost, threshold)
    matches = sorted(matches, key=lambda t: cost[t[0], t[1], t[2]])
    if not matches:
        return out
    Ps = projection_matrices(Ks, RTs)
    X, rep = [], []
    for (i, j, k) in matches:
        pts = np.array([c1[i], c2[j], c3[k]], dtype=np.float64)
        x = triangulate_multi_view(Ps, pts)
        X.append(x)
        rep.append([reprojection_error(Ps[v], x, pts[v]) for v in range(3)])
    out['idx'] = np.asarray(matches, np.int64)
    out['cost'] = np.asarray([cost[i, j, k] for (i, j, k) in matches], np.float32)
    out['X'] = np.asarray(X)
    out['reproj'] = np.asarray(rep)
    return out
