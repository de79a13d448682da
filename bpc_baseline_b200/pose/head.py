"""Batched pose head on the crop stream: SimplePoseNet (a torchvision ResNet50 -- LIBRARY code, cuDNN / cuBLAS kernels) run in
bfloat16 channels-last directly on the chunk buffer that the crop kernel fills, plus the batched rotation decode.

Reference: bpc/inference/process_pose.py:210-239 runs the network once per crop (batch 1, float32, one H2D copy per crop)
and decodes on the host.  Here the chunk buffer of ``MatchCropPipeline(crop_dtype=torch.bfloat16)`` -- bfloat16, channels-last,
written by ``bpc_roi_crop_bf16`` -- is the network input as it stands: no cast, no layout pass, no copy.  Nothing in this file
is a hand-written kernel; it exists to close the loop and to measure whether the consumer can drain the crop kernel.
"""
from __future__ import annotations

from typing import Optional

import torch


class PoseHeadConsumer:
    """``consumer`` for ``MatchCropPipeline.run_device``: forwards every finished chunk through the pose network.

    ``raw`` float32 [capacity, out_dim] receives the network outputs at the ROI's index (row 3 * match + view);
    ``rotations(n)`` decodes the first n rows to rotation matrices on the device (process_pose.py:213-229)."""

    def __init__(self, model: torch.nn.Module, capacity: int, *, mode: Optional[str] = None, sub_batch: int = 2048,
                 dtype: torch.dtype = torch.bfloat16, device='cuda'):
        self.device = torch.device(device)
        self.dtype = dtype
        self.model = model.to(self.device).eval()
        if dtype != torch.float32:
            self.model = self.model.to(dtype)
        self.model = self.model.to(memory_format=torch.channels_last)
        with torch.no_grad():
            probe = torch.zeros((1, 3, 64, 64), dtype=dtype, device=self.device).contiguous(memory_format=torch.channels_last)
            self.out_dim = int(self.model(probe).shape[1])
        self.mode = mode or {3: 'euler', 4: 'quat', 6: '6d'}[self.out_dim]
        self.sub_batch = int(sub_batch)
        self.raw = torch.zeros((int(capacity), self.out_dim), dtype=torch.float32, device=self.device)
        self.crops_seen = 0

    @torch.no_grad()
    def __call__(self, crops: torch.Tensor, first: int) -> None:
        if crops.dtype != self.dtype:
            crops = crops.to(self.dtype)
        if not crops.is_contiguous(memory_format=torch.channels_last):
            crops = crops.contiguous(memory_format=torch.channels_last)
        n = int(crops.shape[0])
        for lo in range(0, n, self.sub_batch):
            hi = min(n, lo + self.sub_batch)
            self.raw[first + lo:first + hi] = self.model(crops[lo:hi]).float()
        self.crops_seen += n

    @torch.no_grad()
    def rotations(self, n: Optional[int] = None) -> torch.Tensor:
        """Rotation matrices float32 [n, 3, 3] on the device, the batched decode of process_pose.py:213-229."""
        from ..inference.process_pose import quat_to_rotmat, rotmat_from_6d, rotmat_from_euler
        raw = self.raw if n is None else self.raw[:n]
        if self.mode == 'euler':
            wrapped = torch.remainder(raw + torch.pi, 2 * torch.pi) - torch.pi
            return rotmat_from_euler(wrapped)
        if self.mode == 'quat':
            return quat_to_rotmat(raw / (raw.norm(dim=1, keepdim=True) + 1e-8))
        if self.mode == '6d':
            return rotmat_from_6d(raw)
        raise ValueError('Unsupported rotation mode.')


def geodesic_degrees(Ra: torch.Tensor, Rb: torch.Tensor) -> torch.Tensor:
    """Angle of Ra^T Rb in degrees, per matrix pair."""
    tr = (Ra.transpose(1, 2) @ Rb).diagonal(dim1=1, dim2=2).sum(dim=1)
    return torch.rad2deg(torch.acos(((tr - 1) / 2).clamp(-1, 1)))
