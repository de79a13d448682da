"""One scene at a time at launch-latency cost: the whole hot path as ONE CUDA graph.

The reference processes a scene with ``PoseEstimator._match`` + the crop half of ``_estimate_rotation``
(process_pose.py:144-209): ~0.5 s of Python at 20 detections.  On the GPU the kernels for one scene take tens of
microseconds, so the cost of a single-scene call is launch overhead and host<->device round trips.  ``SceneSession``
removes both: pinned staging buffers and device buffers are allocated once, and

    H2D(K, RT, boxes, counts) -> bpc_box_centers -> bpc_match_triangulate -> bpc_build_rois -> bpc_roi_crop
    -> D2H(idx, n, cost, X, reproj, number of ROIs)

is captured once into a ``torch.cuda.CUDAGraph`` (the C-ABI launchers never allocate or synchronise, so they capture
as they are).  ``run`` copies the scene into the pinned staging buffers, replays the graph and waits for the stream:
one launch, one synchronisation.  The crops stay in HBM (the pose network consumes them there).

The images of the capture are uploaded by ``set_images`` (3 x 24.9 MB at IPD resolution); in a live pipeline the
detector has already put them on the GPU, and ``set_images`` accepts device tensors without a copy.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import batched
from ._host import nvtx_range as _Range


class SceneSession:
    def __init__(self, Dmax: int = 32, T: int = 256, image_shape: Sequence[int] = (3, 2160, 3840, 3), threshold=30,
                 swap_rb: bool = True, fill=(255, 255, 255), reproj_thresh: Optional[float] = None, device='cuda'):
        self.D, self.T = int(Dmax), int(T)
        self.device = dev = torch.device(device if str(device) != 'cuda' else f'cuda:{torch.cuda.current_device()}')
        self.threshold, self.swap_rb, self.fill, self.reproj_thresh = threshold, swap_rb, tuple(fill), reproj_thresh
        D = self.D
        pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
        # pinned staging: inputs, then results
        self.h = {'Ks': pin((1, 3, 3, 3), torch.float32), 'RTs': pin((1, 3, 4, 4), torch.float64),
                  'boxes': pin((1, 3, D, 4), torch.int32), 'counts': pin((1, 3), torch.int32),
                  'idx': pin((1, D, 3), torch.int32), 'n': pin((1,), torch.int32), 'cost': pin((1, D), torch.float32),
                  'X': pin((1, D, 3), torch.float64), 'reproj': pin((1, D, 3), torch.float64), 'n_rois': pin((1,), torch.int32)}
        self.d = {k: torch.zeros(self.h[k].shape, dtype=self.h[k].dtype, device=dev) for k in ('Ks', 'RTs', 'boxes', 'counts')}
        self.images = torch.zeros(tuple(image_shape), dtype=torch.uint8, device=dev)
        self._images_pinned = None
        self.image_of_scene = torch.arange(3, dtype=torch.int32, device=dev).reshape(1, 3)
        self.rois = torch.zeros((D * 3, 5), dtype=torch.int32, device=dev)
        self.crops = torch.zeros((D * 3, 3, self.T, self.T), dtype=torch.float32, device=dev)
        self.status = torch.zeros((D * 3,), dtype=torch.int32, device=dev)
        self.lut = batched.normalise_lut(dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = None
        self._res = None

    # ------------------------------------------------------------------------------------------------
    def set_images(self, images) -> None:
        """The three views of the capture: uint8 [3,H,W,3] (NumPy / list of arrays on the host, or a CUDA tensor)."""
        if isinstance(images, torch.Tensor) and images.is_cuda:
            if tuple(images.shape) != tuple(self.images.shape) or images.dtype != torch.uint8:
                raise RuntimeError(f'images must be uint8 {tuple(self.images.shape)}')
            with torch.cuda.stream(self.stream):
                self.images.copy_(images, non_blocking=True)
            return
        if self._images_pinned is None:
            self._images_pinned = torch.empty(tuple(self.images.shape), dtype=torch.uint8).pin_memory()
        dst = self._images_pinned.numpy()
        for v in range(self.images.shape[0]):
            a = np.asarray(images[v])
            if a.shape != dst[v].shape or a.dtype != np.uint8:
                raise RuntimeError(f'view {v}: expected uint8 {dst[v].shape}, got {a.dtype} {a.shape}')
            np.copyto(dst[v], a)
        with torch.cuda.stream(self.stream):
            self.images.copy_(self._images_pinned, non_blocking=True)

    def _body(self):
        d, h = self.d, self.h
        for k in d:
            d[k].copy_(h[k], non_blocking=True)
        with _Range('bpc.match'):
            centers = batched.box_centers(d['boxes'])
            res = batched.match_triangulate(d['Ks'], d['RTs'], centers, d['counts'], self.threshold, want_reproj=True,
                                            reproj_thresh=self.reproj_thresh)
        with _Range('bpc.build_rois'):
            rois, offs = batched.build_rois(d['boxes'], res.idx, res.n, self.image_of_scene, rois=self.rois)
        with _Range('bpc.crop'):
            batched.roi_crop(self.images, rois, T=self.T, fill=self.fill, swap_rb=self.swap_rb, lut=self.lut,
                             n_rois=offs[1:2], roi_first=0, out=self.crops, status=self.status)
        h['idx'].copy_(res.idx, non_blocking=True); h['n'].copy_(res.n, non_blocking=True)
        h['cost'].copy_(res.cost, non_blocking=True); h['X'].copy_(res.X, non_blocking=True)
        h['reproj'].copy_(res.reproj, non_blocking=True); h['n_rois'].copy_(offs[1:2], non_blocking=True)
        return res

    def capture(self) -> None:
        """Warm up (tensor maps, workspaces, function attributes) and capture the graph; called lazily by ``run``."""
        with torch.cuda.stream(self.stream):
            for _ in range(2):
                self._body()
        self.stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            self._res = self._body()
        self.graph = g

    # ------------------------------------------------------------------------------------------------
    def run(self, Ks, RTs, boxes: Sequence, sync: bool = True) -> dict:
        """One scene.  ``Ks`` 3 x (3,3) float32, ``RTs`` 3 x (4,4) float64, ``boxes[c]`` int (n_c, 4) = (x1, y1, x2, y2).

        Returns the pinned result tensors (valid after the stream synchronisation this call performs when ``sync``):
        idx [n,3], cost [n], X [n,3], reproj [n,3], ``n`` and ``crops`` = float32 CUDA tensor [3n,3,T,T]
        (row 3m + v = match m, view v; a view of the session's buffer, overwritten by the next ``run``)."""
        h = self.h
        hk, hrt, hb, hc = h['Ks'].numpy(), h['RTs'].numpy(), h['boxes'].numpy(), h['counts'].numpy()
        for c in range(3):
            k = np.asarray(Ks[c])
            if k.dtype != np.float32:
                raise TypeError('capture.Ks must be float32 (camera_utils.py:16)')
            hk[0, c] = k
            hrt[0, c] = np.asarray(RTs[c], np.float64)
            b = np.asarray(boxes[c], np.int32).reshape(-1, 4)
            if len(b) > self.D:
                raise RuntimeError(f'{len(b)} detections exceed Dmax={self.D} of this session')
            hb[0, c, :len(b)] = b
            hc[0, c] = len(b)
        if self.graph is None:
            self.capture()
        with _Range('bpc.scene'), torch.cuda.stream(self.stream):      # CUDAGraph.replay launches on the CURRENT stream
            self.graph.replay()
        if not sync:
            return None
        self.stream.synchronize()
        n = int(h['n'][0])
        if n < 0:
            batched.check_match_status(h['n'])
        return {'n': n, 'idx': h['idx'][0, :n], 'cost': h['cost'][0, :n], 'X': h['X'][0, :n], 'reproj': h['reproj'][0, :n],
                'crops': self.crops[:3 * n], 'n_rois': int(h['n_rois'][0])}

    def rejected(self) -> int:
        """Number of ROIs of the last scene the crop kernel rejected (empty box / resized side < 1): device read."""
        return int(self.status[:int(self.h['n_rois'][0])].sum())
