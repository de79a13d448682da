"""Synthetic IPD-like scenes for tests and bench.py.

NumPy only; no oracle, no CUDA.  Scenes are generated in fixed chunks of
``CHUNK`` scenes, chunk ``c`` seeded with ``default_rng([seed, c])``, so a rank
that owns scenes ``[lo, hi)`` reproduces exactly the scenes a single process
would have produced for that range (the multi-GPU equality tests rely on it).

Layout of a ``SceneBatch`` (the batched layout the CUDA path consumes):

    Ks      f32 [S, 3, 3, 3]      intrinsics per camera (camera_utils.py:16)
    RTs     f64 [S, 3, 4, 4]      world->camera, f32-rounded values widened to f64
                                  (data_utils.py:383-387 builds them with np.eye(4))
    boxes   i32 [S, 3, Dmax, 4]   (x1, y1, x2, y2), int-truncated (process_pose.py:134)
    centers f64 [S, 3, Dmax, 2]   0.5*(x1+x2), 0.5*(y1+y2) (process_pose.py:135-136)
    counts  i32 [S, 3]            detections per camera (ragged)
    truth   i32 [S, 3, Dmax]      object id behind each detection, -1 = false detection
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED = 20250131
CHUNK = 256
IMG_W, IMG_H = 3840, 2160


@dataclass
class SceneBatch:
    Ks: np.ndarray
    RTs: np.ndarray
    boxes: np.ndarray
    centers: np.ndarray
    counts: np.ndarray
    truth: np.ndarray
    width: int = IMG_W
    height: int = IMG_H

    def __len__(self) -> int:
        return self.Ks.shape[0]

    def slice(self, lo: int, hi: int) -> "SceneBatch":
        return SceneBatch(self.Ks[lo:hi], self.RTs[lo:hi], self.boxes[lo:hi], self.centers[lo:hi],
                          self.counts[lo:hi], self.truth[lo:hi], self.width, self.height)

    def detections(self, s: int):
        """Scene ``s`` as the reference's ``detections`` dict (process_pose.py:137-141)."""
        out = {}
        for c in range(3):
            n = int(self.counts[s, c])
            out[c] = [{'bbox': tuple(int(v) for v in self.boxes[s, c, d]),
                       'bb_center': (float(self.centers[s, c, d, 0]), float(self.centers[s, c, d, 1]))}
                      for d in range(n)]
        return out

    def capture_arrays(self, s: int):
        """(Ks list of f32 3x3, RTs list of f64 4x4) as ``Capture`` holds them."""
        return [self.Ks[s, c].copy() for c in range(3)], [self.RTs[s, c].copy() for c in range(3)]


def _normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _chunk(rng, S, D, sigma, p_drop, n_dup, n_false, width, height, side_lo, side_hi):
    # --- rig: 3 cameras on a ring looking at the bin centre --------------------------------
    az = np.array([0.0, 2.1, 4.2])[None, :] + rng.uniform(-0.2, 0.2, (S, 3))
    el = rng.uniform(0.9, 1.2, (S, 3))
    rad = rng.uniform(1600.0, 2000.0, (S, 3))
    c = np.stack([rad * np.cos(el) * np.cos(az), rad * np.cos(el) * np.sin(az), rad * np.sin(el)], -1)
    zc = _normalize(-c)                                   # optical axis
    up = np.broadcast_to(np.array([0.0, 0.0, 1.0]), zc.shape)
    xc = _normalize(np.cross(zc, up))
    yc = np.cross(zc, xc)                                 # image y points "down"
    R = np.stack([xc, yc, zc], axis=-2)                   # rows = camera axes (w2c)
    t = -np.einsum('scij,scj->sci', R, c)
    R32, t32 = R.astype(np.float32), t.astype(np.float32)
    RTs = np.zeros((S, 3, 4, 4), np.float64)
    RTs[..., :3, :3] = R32
    RTs[..., :3, 3] = t32
    RTs[..., 3, 3] = 1.0
    f = rng.uniform(3981.0, 4209.0, (S, 3))
    Ks = np.zeros((S, 3, 3, 3), np.float32)
    Ks[..., 0, 0] = f
    Ks[..., 1, 1] = f
    Ks[..., 0, 2] = width / 2
    Ks[..., 1, 2] = height / 2
    Ks[..., 2, 2] = 1.0
    # --- objects in a 600 x 400 x 150 mm bin ----------------------------------------------
    X = rng.uniform(-0.5, 0.5, (S, D, 3)) * np.array([600.0, 400.0, 150.0])
    Dmax = D + n_dup + n_false
    boxes = np.zeros((S, 3, Dmax, 4), np.int32)
    truth = np.full((S, 3, Dmax), -1, np.int32)
    counts = np.zeros((S, 3), np.int32)
    Xc = np.einsum('scij,sdj->scdi', RTs[..., :3, :3], X) + RTs[..., None, :3, 3]
    uvw = np.einsum('scij,scdj->scdi', Ks.astype(np.float64), Xc)
    uv = uvw[..., :2] / uvw[..., 2:3] + rng.normal(0.0, sigma, (S, 3, D, 2))
    perm = np.argsort(rng.random((S, 3, D)), axis=-1)
    keep = rng.random((S, 3, D)) >= p_drop
    wh = rng.integers(side_lo, side_hi, (S, 3, Dmax, 2))
    false_uv = rng.uniform(0.15, 0.85, (S, 3, max(n_false, 1), 2)) * np.array([width, height])
    dup_src = rng.integers(0, D, (S, 3, max(n_dup, 1)))
    if p_drop == 0.0 and n_dup == 0 and n_false == 0:
        # every camera keeps all D objects: the per-(scene, camera) loop below, vectorised (same values)
        pts = np.take_along_axis(uv, perm[..., None], axis=2)
        w = wh[..., 0].astype(np.float64)
        h = wh[..., 1].astype(np.float64)
        x1 = np.clip(np.trunc(pts[..., 0] - w / 2).astype(np.int64), 0, width - 8)
        y1 = np.clip(np.trunc(pts[..., 1] - h / 2).astype(np.int64), 0, height - 8)
        x2 = np.clip(x1 + w.astype(np.int64), x1 + 8, width)
        y2 = np.clip(y1 + h.astype(np.int64), y1 + 8, height)
        boxes[...] = np.stack([x1, y1, x2, y2], -1)
        truth[...] = perm
        counts[...] = D
        vectorised = True
    else:
        vectorised = False
    for s in range(0 if vectorised else S):
        for cam in range(3):
            ids = perm[s, cam][keep[s, cam][perm[s, cam]]]
            pts = uv[s, cam, ids]
            ident = ids.astype(np.int32)
            if n_dup and len(ids):
                src = dup_src[s, cam, :n_dup] % len(ids)
                pts = np.concatenate([pts, pts[src]])
                ident = np.concatenate([ident, ident[src]])
            if n_false:
                pts = np.concatenate([pts, false_uv[s, cam, :n_false]])
                ident = np.concatenate([ident, np.full(n_false, -1, np.int32)])
            n = len(pts)
            if n == 0:
                continue
            w = wh[s, cam, :n, 0].astype(np.float64)
            h = wh[s, cam, :n, 1].astype(np.float64)
            if n_dup and len(ids):        # duplicates are exact copies of the box, too
                nd = len(ids)
                w[nd:nd + n_dup] = w[src]
                h[nd:nd + n_dup] = h[src]
            x1 = np.trunc(pts[:, 0] - w / 2).astype(np.int64)
            y1 = np.trunc(pts[:, 1] - h / 2).astype(np.int64)
            x1 = np.clip(x1, 0, width - 8)
            y1 = np.clip(y1, 0, height - 8)
            x2 = np.clip(x1 + w.astype(np.int64), x1 + 8, width)
            y2 = np.clip(y1 + h.astype(np.int64), y1 + 8, height)
            boxes[s, cam, :n] = np.stack([x1, y1, x2, y2], -1)
            truth[s, cam, :n] = ident
            counts[s, cam] = n
    centers = np.zeros((S, 3, Dmax, 2), np.float64)
    centers[..., 0] = 0.5 * (boxes[..., 0].astype(np.int64) + boxes[..., 2])
    centers[..., 1] = 0.5 * (boxes[..., 1].astype(np.int64) + boxes[..., 3])
    return Ks, RTs, boxes, centers, counts, truth


def make_scenes(S: int, D: int, *, seed: int = SEED, first: int = 0, sigma: float = 1.0,
                p_drop: float = 0.0, n_dup: int = 0, n_false: int = 0,
                width: int = IMG_W, height: int = IMG_H,
                side_lo: int = 60, side_hi: int = 400) -> SceneBatch:
    """Scenes ``first .. first+S-1`` of the stream defined by (seed, D, sigma, ...).

    ``first`` must be a multiple of CHUNK.  D objects per scene; every camera sees a random
    permutation of them, each dropped with ``p_drop``; ``n_dup`` exact duplicate detections and
    ``n_false`` false detections per camera exercise LSAP ties / conflicts.
    """
    if first % CHUNK:
        raise ValueError(f"first={first} must be a multiple of {CHUNK}")
    parts = []
    done = 0
    while done < S:
        n = min(CHUNK, S - done)
        rng = np.random.default_rng([seed, (first + done) // CHUNK])
        part = _chunk(rng, CHUNK, D, sigma, p_drop, n_dup, n_false, width, height, side_lo, side_hi)
        parts.append(tuple(a[:n] for a in part))
        done += n
    cat = [np.concatenate([p[i] for p in parts]) for i in range(6)]
    return SceneBatch(*cat, width=width, height=height)


def make_images(B: int, *, seed: int = SEED, width: int = IMG_W, height: int = IMG_H) -> np.ndarray:
    """Pool of ``B`` BGR uint8 images [B, H, W, 3]: smooth gradients + noise.

    Gradients land on .5 rounding boundaries of the resize far more often than white noise.
    """
    rng = np.random.default_rng([seed, 7777])
    yy = np.arange(height, dtype=np.float32)[:, None]
    xx = np.arange(width, dtype=np.float32)[:, None]
    out = np.empty((B, height, width, 3), np.uint8)
    for b in range(B):
        fx = rng.uniform(0.01, 0.2, 3).astype(np.float32)
        fy = rng.uniform(0.01, 0.2, 3).astype(np.float32)
        ph = rng.uniform(0, 6.28, 3).astype(np.float32)
        gx = (127.5 + 70.0 * np.sin(xx * fx + ph)).astype(np.float32)       # [W, 3]
        gy = (50.0 * np.cos(yy * fy - ph)).astype(np.float32)               # [H, 3]
        img = np.rint(gy[:, None, :] + gx[None, :, :]).astype(np.int16)
        img += rng.integers(-12, 13, img.shape, dtype=np.int8)
        out[b] = np.clip(img, 0, 255).astype(np.uint8)
    return out


def rois_for_matches(boxes: np.ndarray, idx: np.ndarray, n: np.ndarray, image_of_scene: np.ndarray) -> np.ndarray:
    """ROI records i32 [R, 5] = (image index, x1, y1, x2, y2), 3 per match, in (scene, match, view) order.

    ``idx`` i32 [S, Kmax, 3] / ``n`` i32 [S] are the matcher's outputs; ``image_of_scene`` [S, 3]
    maps (scene, view) to an index into the image pool.  Host mirror of the device ROI builder.
    """
    rois = []
    for s in range(len(n)):
        for m in range(int(n[s])):
            for v in range(3):
                d = int(idx[s, m, v])
                rois.append((int(image_of_scene[s, v]), *[int(t) for t in boxes[s, v, d]]))
    return np.asarray(rois, np.int32).reshape(-1, 5)
