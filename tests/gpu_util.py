"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch


def to_dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def pack_scenes(scenes):
    """List of golden scene dicts (tests/conftest.py) -> batched arrays padded to a common Dmax."""
    D = max(int(sc['centers_arr'].shape[1]) for sc in scenes)
    S = len(scenes)
    Ks = np.zeros((S, 3, 3, 3), np.float32)
    RTs = np.zeros((S, 3, 4, 4), np.float64)
    centers = np.zeros((S, 3, D, 2), np.float64)
    boxes = np.zeros((S, 3, D, 4), np.int32)
    counts = np.zeros((S, 3), np.int32)
    for s, sc in enumerate(scenes):
        d = sc['centers_arr'].shape[1]
        Ks[s], RTs[s], counts[s] = sc['Ks_arr'], sc['RTs_arr'], sc['counts']
        centers[s, :, :d] = sc['centers_arr']
        boxes[s, :, :d] = sc['boxes']
    return Ks, RTs, centers, boxes, counts


def batch_to_dev(batch):
    return (to_dev(batch.Ks), to_dev(batch.RTs), to_dev(batch.centers), to_dev(batch.boxes), to_dev(batch.counts))


def rel_err(a, b):
    """Norm-wise relative error per row (||dX|| / ||X_ref||)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-300)
