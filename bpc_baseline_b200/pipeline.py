"""Batched match -> triangulate -> ROI-crop pipeline on one GPU (the user-facing batched call).

One pass = PoseEstimator._match (process_pose.py:144-188) for S scenes in one kernel, the ROI list of
every matched detection (process_pose.py:195-201), and the network-input crops
(process_pose.py:199-209) produced chunk by chunk into a reusable buffer -- at T=224 a single crop is
602 KB, so the crops of a large batch never exist all at once; each finished chunk is handed to
``consumer`` (e.g. the pose network), in stream order.

``run_device`` works on device-resident tensors; ``run_host`` is the same pass from pinned host
buffers, with the host->device copies of its inputs and the device->host read of the pose records.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import batched
from ._host import nvtx_range


class MatchCropPipeline:
    def __init__(self, S: int, Dmax: int, *, T: int = 224, chunk_rois: int = 16384, threshold=30,
                 swap_rb: bool = True, fill=(255, 255, 255), device='cuda', want_reproj: bool = True,
                 reproj_thresh: Optional[float] = None, crops: Optional[torch.Tensor] = None,
                 crop_dtype: torch.dtype = torch.float32):
        self.S, self.D, self.T = int(S), int(Dmax), int(T)
        self.device = torch.device(device)
        self.threshold = threshold
        self.swap_rb, self.fill, self.want_reproj, self.reproj_thresh = swap_rb, tuple(fill), want_reproj, reproj_thresh
        self.cap = self.S * self.D * 3                     # ROI capacity: 3 views x at most Dmax matches
        self.chunk = max(1, min(int(chunk_rois), self.cap))
        dev = self.device
        self.rois = torch.zeros((self.cap, 5), dtype=torch.int32, device=dev)
        if crop_dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError('crop_dtype must be torch.float32 or torch.bfloat16')
        self.crop_dtype = crop_dtype
        if crop_dtype == torch.bfloat16:                   # bf16 channels-last chunk buffer: the pose head's input as it stands
            if crops is not None:
                raise RuntimeError('a caller-owned chunk buffer is float32 only')
            self.crops = torch.empty((self.chunk, 3, self.T, self.T), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
        elif crops is not None:                            # a caller-owned chunk buffer shared between pipelines
            if crops.dtype != torch.float32 or not crops.is_contiguous() or crops.numel() < self.chunk * 3 * self.T * self.T:
                raise RuntimeError('crops must be a contiguous float32 buffer of at least chunk_rois * 3 * T * T elements')
            self.crops = crops.view(-1)[:self.chunk * 3 * self.T * self.T].view(self.chunk, 3, self.T, self.T)
        else:
            self.crops = torch.empty((self.chunk, 3, self.T, self.T), dtype=torch.float32, device=dev)
        self.status = torch.zeros((self.cap,), dtype=torch.int32, device=dev)
        self.lut = batched.normalise_lut(dev)
        self._host = None
        self.last_crop_launches = 0

    # ---------------------------------------------------------------------------------------------
    def run_device(self, Ks, RTs, centers, counts, boxes, images, image_of_scene,
                   consumer: Optional[Callable[[torch.Tensor, int], None]] = None,
                   n_rois_host: Optional[int] = None, events=None):
        """One pass on device-resident inputs.  Returns (MatchResult, scene_offset i32 [S+1]).

        ``n_rois_host``: if the caller knows the ROI count (e.g. from a previous identical pass) only
        the chunks that contain ROIs are launched; otherwise every chunk of the capacity is launched and
        the kernel skips records beyond the device-side count -- no host synchronisation either way.
        ``events``: optional (e_match, e_crop0, e_crop1) CUDA events recorded after the matcher, before
        the first and after the last crop launch.
        """
        with nvtx_range('bpc.match'):
            res = batched.match_triangulate(Ks, RTs, centers, counts, self.threshold, want_reproj=self.want_reproj,
                                            reproj_thresh=self.reproj_thresh)
        with nvtx_range('bpc.build_rois'):
            rois, offs = batched.build_rois(boxes, res.idx, res.n, image_of_scene, rois=self.rois)
        total = offs[self.S:self.S + 1]
        if events is not None:
            events[0].record()
            events[1].record()
        upto = self.cap if n_rois_host is None else min(self.cap, int(n_rois_host))
        launches = 0
        for first in range(0, upto, self.chunk):
            r = min(self.chunk, self.cap - first)
            with nvtx_range('bpc.crop_chunk'):
                crop = batched.roi_crop_bf16 if self.crop_dtype == torch.bfloat16 else batched.roi_crop
                out = crop(images, rois[first:first + r], T=self.T, fill=self.fill, swap_rb=self.swap_rb,
                           lut=self.lut, n_rois=total, roi_first=first, out=self.crops, status=self.status[first:first + r])
            launches += 1
            if consumer is not None:
                consumer(out[:r], first)
        if events is not None:
            events[2].record()
        self.last_crop_launches = launches
        return res, offs

    # ---------------------------------------------------------------------------------------------
    def _alloc_host(self, images_shape):
        S, D, dev = self.S, self.D, self.device
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        h = {
            'Ks': pin((S, 3, 3, 3), torch.float32), 'RTs': pin((S, 3, 4, 4), torch.float64),
            'boxes': pin((S, 3, D, 4), torch.int32), 'counts': pin((S, 3), torch.int32),
            'image_of_scene': pin((S, 3), torch.int32), 'images': pin(tuple(images_shape), torch.uint8),
            'idx': pin((S, D, 3), torch.int32), 'n': pin((S,), torch.int32), 'cost': pin((S, D), torch.float32),
            'X': pin((S, D, 3), torch.float64), 'n_rois': pin((1,), torch.int32),
        }
        d = {k: torch.empty(h[k].shape, dtype=h[k].dtype, device=dev)
             for k in ('Ks', 'RTs', 'boxes', 'counts', 'image_of_scene', 'images')}
        self._host = (h, d)

    def host_buffers(self, images_shape):
        """Pinned staging buffers (dict of CPU tensors) the caller fills before ``run_host``."""
        if self._host is None or tuple(self._host[0]['images'].shape) != tuple(images_shape):
            self._alloc_host(images_shape)
        return self._host[0]

    def run_host(self, consumer=None, n_rois_host=None):
        """One pass from the pinned host buffers of ``host_buffers``: H2D of all inputs, the device
        pass, D2H of the pose records.  Returns the dict of pinned host tensors (idx, n, cost, X, n_rois).

        Centres are derived on the device from the integer boxes (process_pose.py:134-136).
        """
        h, d = self._host
        for k in d:
            d[k].copy_(h[k], non_blocking=True)
        centers = batched.box_centers(d['boxes'])
        res, offs = self.run_device(d['Ks'], d['RTs'], centers, d['counts'], d['boxes'], d['images'],
                                    d['image_of_scene'], consumer=consumer, n_rois_host=n_rois_host)
        h['idx'].copy_(res.idx, non_blocking=True)
        h['n'].copy_(res.n, non_blocking=True)
        h['cost'].copy_(res.cost, non_blocking=True)
        h['X'].copy_(res.X, non_blocking=True)
        h['n_rois'].copy_(offs[self.S:self.S + 1], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h

    def run_host_stream(self, steps: int, refill: Optional[Callable[[dict, int], None]] = None,
                        on_result: Optional[Callable[[dict, int], None]] = None, consumer=None):
        """``steps`` passes from pinned host buffers with the uploads double buffered.

        The host->device copy of step k+1 runs on a copy stream while step k computes, and the device->host
        read of step k is collected one step later, so in steady state a step costs max(compute, copies)
        instead of their sum.  ``refill(host_buffers, k)`` is called before step k is uploaded (the buffers are
        reusable then); ``on_result(results, k)`` receives the pinned result tensors of step k.  Every step's
        inputs are copied H2D and every step's results D2H.
        """
        h, d = self._host
        dev = self.device
        if not hasattr(self, '_stream_state'):
            d2 = {k: torch.empty_like(v) for k, v in d.items()}
            res_keys = ('idx', 'n', 'cost', 'X', 'n_rois')
            h2 = {k: torch.empty_like(h[k]).pin_memory() for k in res_keys}
            self._stream_state = {'d': (d, d2), 'hres': ({k: h[k] for k in res_keys}, h2),
                                  'copy': torch.cuda.Stream(device=dev)}
        st = self._stream_state
        main = torch.cuda.current_stream(dev)
        up_done = [torch.cuda.Event(), torch.cuda.Event()]        # upload into set j finished
        free = [torch.cuda.Event(), torch.cuda.Event()]           # compute finished reading set j
        res_done = [torch.cuda.Event(), torch.cuda.Event()]       # results of set j are in pinned memory
        host_read = [torch.cuda.Event(), torch.cuda.Event()]      # the copy stream finished reading the pinned inputs

        def upload(k):
            j = k & 1
            if refill is not None:
                if k >= 1:
                    host_read[(k - 1) & 1].synchronize()           # previous upload no longer reads the pinned inputs
                refill(h, k)
            with torch.cuda.stream(st['copy']):
                if k >= 2:
                    st['copy'].wait_event(free[j])
                for key, t in st['d'][j].items():
                    t.copy_(h[key], non_blocking=True)
                up_done[j].record(st['copy'])
                host_read[j].record(st['copy'])

        upload(0)
        for k in range(steps):
            j = k & 1
            if k + 1 < steps:
                upload(k + 1)
            main.wait_event(up_done[j])
            dj = st['d'][j]
            centers = batched.box_centers(dj['boxes'])
            res, offs = self.run_device(dj['Ks'], dj['RTs'], centers, dj['counts'], dj['boxes'], dj['images'],
                                        dj['image_of_scene'], consumer=consumer)
            free[j].record(main)
            if k >= 2:
                res_done[j].synchronize()                           # pinned result set j was handed out two steps ago
            hr = st['hres'][j]
            hr['idx'].copy_(res.idx, non_blocking=True)
            hr['n'].copy_(res.n, non_blocking=True)
            hr['cost'].copy_(res.cost, non_blocking=True)
            hr['X'].copy_(res.X, non_blocking=True)
            hr['n_rois'].copy_(offs[self.S:self.S + 1], non_blocking=True)
            res_done[j].record(main)
            if k >= 1:                                              # collect step k-1 while step k runs
                res_done[(k - 1) & 1].synchronize()
                if on_result is not None:
                    on_result(st['hres'][(k - 1) & 1], k - 1)
        res_done[(steps - 1) & 1].synchronize()
        if on_result is not None:
            on_result(st['hres'][(steps - 1) & 1], steps - 1)
        return st['hres'][(steps - 1) & 1]

    def h2d_bytes(self) -> int:
        h, d = self._host
        return int(sum(h[k].numel() * h[k].element_size() for k in d))

    def d2h_bytes(self) -> int:
        h, _ = self._host
        return int(sum(h[k].numel() * h[k].element_size() for k in ('idx', 'n', 'cost', 'X', 'n_rois')))


def algorithmic_crop_bytes(rois: np.ndarray, T: int, out_elem_bytes: int = 4) -> int:
    """per crop 3*h*w source bytes read once + 3*T*T*4 output bytes written once + 20 B record
    (``out_elem_bytes`` = 2 for the bfloat16 variant, 1 for the uint8 letterboxed image)."""
    rois = np.asarray(rois, np.int64)
    w = rois[:, 3] - rois[:, 1]
    h = rois[:, 4] - rois[:, 2]
    return int((3 * w * h).sum() + len(rois) * (3 * T * T * out_elem_bytes + 20))
