"""Host <-> device glue for the single-scene drop-in functions (NumPy in, NumPy out, CUDA in between)."""
from __future__ import annotations

import numpy as np
import torch


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('bpc_baseline_b200 needs a CUDA device: the hot path has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def to_dev(a, dtype) -> torch.Tensor:
    """Contiguous copy of a host array on the current CUDA device with the given NumPy dtype."""
    arr = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    return torch.from_numpy(arr).to(device())


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


class nvtx_range:
    """NVTX range around a stage of the hot path (`bpc.match`, `bpc.build_rois`, `bpc.crop_chunk`,
    `bpc.scene`): visible in Nsight Systems / ncu --nvtx, a no-op costing two C calls otherwise."""
    __slots__ = ('name',)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False
