"""Scene-sharded data parallelism over the GPUs of one node.

Scenes are independent (no term of PoseEstimator._match couples two scenes), so the batch is cut into
contiguous blocks, one per rank, and every rank runs the whole hot path on its block with no exchange.  The
only collective is the final gather of fixed-width pose records; crops stay on the GPU that produced them
(their consumer, the pose network, is data-parallel as well -- gathering float32 crops into one GPU would be
bounded by a single NVLink ingest, ~7x below one GPU's HBM rate).

The pack / gather / unpack helpers work on CPU tensors with the gloo backend too; that is how the host-side
logic is tested without GPUs.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

RECORD_WIDTH = 8            # i, j, k, cost, X, Y, Z, n_of_scene  (float64: 64 bytes per match slot)


def shard_range(total: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` scenes owned by ``rank``; block starts are multiples of ``align``."""
    if not 0 <= rank < world:
        raise ValueError('rank out of range')
    units = (total + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(total, lo_u * align), min(total, hi_u * align)


def pack_records(idx: torch.Tensor, n: torch.Tensor, cost: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """[S, Kmax, 8] float64 records from the matcher outputs (integers and float32 costs are exact in float64)."""
    S, K, _ = idx.shape
    rec = torch.empty((S, K, RECORD_WIDTH), dtype=torch.float64, device=idx.device)
    rec[..., 0:3] = idx.to(torch.float64)
    rec[..., 3] = cost.to(torch.float64)
    rec[..., 4:7] = X
    rec[..., 7] = n.to(torch.float64)[:, None]
    return rec


def unpack_records(rec: torch.Tensor) -> dict:
    """Inverse of pack_records on a [..., S, Kmax, 8] tensor (leading rank dimension is folded into S)."""
    rec = rec.reshape(-1, rec.shape[-2], RECORD_WIDTH)
    return {'idx': rec[..., 0:3].to(torch.int32), 'cost': rec[..., 3].to(torch.float32), 'X': rec[..., 4:7].clone(),
            'n': rec[:, 0, 7].to(torch.int32)}


def gather_records(rec: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks receive [world, S, Kmax, 8]; every rank must contribute the same S and Kmax.

    NCCL: one all_gather_into_tensor over NVLink / NVSwitch.  Other backends (gloo on CPU): all_gather.
    """
    if not dist.is_available() or not dist.is_initialized():
        return rec.unsqueeze(0)
    world = dist.get_world_size(group)
    out = torch.empty((world, *rec.shape), dtype=rec.dtype, device=rec.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    else:
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec.contiguous(), group=group)
        for r, p in enumerate(parts):
            out[r] = p
    return out
