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

RECORD_BYTES = 64           # idx i32 x3 | cost f32 | X f64 x3 | reproj f64 x3   (struct PoseRecord, csrc/match.cu)


def shard_range(total: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` scenes owned by ``rank``; block starts are multiples of ``align``."""
    if not 0 <= rank < world:
        raise ValueError('rank out of range')
    units = (total + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(total, lo_u * align), min(total, hi_u * align)


def records_bytes(S: int, K: int) -> int:
    """Size of the record buffer of S scenes x K match slots (= bpc_pack_records_bytes)."""
    return 16 + ((S * 4 + 15) & ~15) + S * K * RECORD_BYTES


def pack_records(res, scene_offset: Optional[torch.Tensor] = None, offset_div: int = 3,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Matcher outputs (``batched.MatchResult`` on the GPU) -> one uint8 record buffer: ONE kernel (bpc_pack_records),
    native dtypes, valid slots only, compacted in scene order behind a header and the per-scene counts."""
    from . import batched
    return batched.pack_records(res, scene_offset, offset_div, out)


def pack_records_torch(idx: torch.Tensor, n: torch.Tensor, cost: torch.Tensor, X: torch.Tensor,
                       reproj: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The same buffer written with torch ops on any device: the executable statement of the layout.  The gloo tests
    use it on CPU tensors, the GPU tests to check the kernel byte for byte; the product path packs with the kernel."""
    S, K, _ = idx.shape
    dev = idx.device
    buf = torch.zeros((records_bytes(S, K),), dtype=torch.uint8, device=dev)
    nn = n.clamp(min=0).to(torch.int64)
    valid = torch.arange(K, device=dev)[None, :] < nn[:, None]
    total = int(nn.sum())
    head = torch.tensor([total, S, K, 0], dtype=torch.int32, device=dev)
    buf[:16] = head.view(torch.uint8)
    buf[16:16 + 4 * S] = n.to(torch.int32).contiguous().view(torch.uint8)
    rec = torch.zeros((total, 16), dtype=torch.int32, device=dev)
    rec[:, 0:3] = idx[valid].to(torch.int32)
    rec[:, 3] = cost[valid].to(torch.float32).view(torch.int32)
    r64 = rec.view(torch.float64)                                     # [total, 8]: X in 2..4, reproj in 5..7
    r64[:, 2:5] = X[valid].to(torch.float64)
    r64[:, 5:8] = reproj[valid].to(torch.float64) if reproj is not None else float('nan')
    off = 16 + ((S * 4 + 15) & ~15)
    buf[off:off + total * RECORD_BYTES] = rec.view(torch.uint8).reshape(-1)
    return buf


def unpack_records(buf: torch.Tensor) -> dict:
    """Inverse of pack_records on one buffer, or on the [world, nbytes] result of gather_records (ranks concatenated in
    rank order): padded idx i32 [S,K,3] (-1), n i32 [S], cost f32 [S,K], X / reproj f64 [S,K,3] (NaN)."""
    if buf.dim() == 2:
        parts = [unpack_records(b) for b in buf.unbind(0)]
        return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
    buf = buf.contiguous()
    dev = buf.device
    total, S, K, _ = (int(v) for v in buf[:16].view(torch.int32).tolist())
    n = buf[16:16 + 4 * S].view(torch.int32).clone()
    off = 16 + ((S * 4 + 15) & ~15)
    rec = buf[off:off + total * RECORD_BYTES].view(torch.int32).reshape(total, 16)
    nn = n.clamp(min=0).to(torch.int64)
    valid = torch.arange(K, device=dev)[None, :] < nn[:, None]
    idx = torch.full((S, K, 3), -1, dtype=torch.int32, device=dev)
    cost = torch.full((S, K), float('nan'), dtype=torch.float32, device=dev)
    X = torch.full((S, K, 3), float('nan'), dtype=torch.float64, device=dev)
    reproj = torch.full((S, K, 3), float('nan'), dtype=torch.float64, device=dev)
    idx[valid] = rec[:, 0:3]
    cost[valid] = rec[:, 3].contiguous().view(torch.float32)
    r64 = rec.view(torch.float64)
    X[valid] = r64[:, 2:5]
    reproj[valid] = r64[:, 5:8]
    return {'idx': idx, 'n': n, 'cost': cost, 'X': X, 'reproj': reproj}


def gather_records(buf: torch.Tensor, group: Optional[dist.ProcessGroup] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All ranks receive [world, nbytes] uint8; every rank must contribute a buffer of the same size (same S and Kmax).

    NCCL: one all_gather_into_tensor over NVLink / NVSwitch.  Other backends (gloo on CPU): all_gather.
    """
    if not dist.is_available() or not dist.is_initialized():
        return buf.unsqueeze(0)
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, buf.numel()), dtype=buf.dtype, device=buf.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf.contiguous(), group=group)
        for r, p in enumerate(parts):
            out[r] = p
    return out


# ------------------------------------------------------------------------------------------------------------
# Crop gather (options 2 and 3).  float32 crops are 602 KB each and one GPU ingests at most
# ~0.9 TB/s over NVLink 5, so the crops cross the wire as the uint8 letterboxed images (147 KB, bpc_roi_crop_u8)
# and the BGR2RGB + to_tensor + normalize tail (process_pose.py:206-209) runs on the receiver.  Two transports:
#   'p2p'  -- the receiver's normalise kernel reads every rank's uint8 buffer directly through peer-mapped
#             pointers: transfer and conversion are ONE kernel, nothing is staged in the receiver's HBM;
#   'nccl' -- dist.gather of the uint8 buffers into a staging area on the receiver, then the same kernel locally.
# ------------------------------------------------------------------------------------------------------------
class PeerBuffers:
    """One uint8 buffer of ``nbytes`` per rank, each mapped into every process of the group (one node, NVLink).

    ``local`` is this rank's buffer; ``views[r]`` aliases rank r's buffer (``views[rank] is local``).  The
    mapping is torch's symmetric memory (CUDA VMM allocations whose handles are exchanged at the rendezvous and
    mapped with access for the local device), so a kernel on this GPU can dereference ``views[r].data_ptr()``.
    Collective: every rank of the group must construct it.
    """

    method = 'symm_mem'

    def __init__(self, nbytes: int, device=None, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError('PeerBuffers needs an initialised process group')
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.nbytes = int(nbytes)
        self.local = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self._handle = symm.rendezvous(self.local, group=group if group is not None else dist.group.WORLD)
        self.views = [self.local if r == self.rank else self._handle.get_buffer(r, (self.nbytes,), torch.uint8)
                      for r in range(self.world)]

    def close(self):
        """Drop the peer mappings (collective: every rank must call it before the buffers are freed)."""
        self.views = []
        self._handle = None
        if dist.is_initialized():
            dist.barrier(group=self.group)


class CropGather:
    """Gathers the crops of every rank onto ``root`` as the float32 network input, chunk by chunk.

    Every rank writes the uint8 crops of chunk i into ``slot(i)`` ([chunk_rois,T,T,3], double buffered) with
    ``batched.roi_crop_u8`` and then calls ``collect(i, counts)``; on ``root`` this returns float32
    [sum(counts),3,T,T] (a view of a reusable buffer), elsewhere None.  ``counts[r]`` = valid crops of rank r in
    this chunk (default: full chunks).  Ordering between ranks comes from one tiny stream-ordered all-reduce per
    chunk: it completes on the receiver only after every producer has enqueued -- and therefore finished -- its crop
    kernel, and a producer reuses slot(i) at chunk i+2, after the all-reduce of chunk i+1, which the receiver
    joins only after it has read chunk i.
    """

    def __init__(self, chunk_rois: int, T: int = 224, root: int = 0, transport: str = 'p2p', swap_rb: bool = True,
                 device=None, group: Optional[dist.ProcessGroup] = None):
        from . import batched
        self._batched = batched
        self.group, self.root, self.T, self.chunk = group, int(root), int(T), int(chunk_rois)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.swap_rb = swap_rb
        self.crop_bytes = self.T * self.T * 3
        if transport not in ('p2p', 'nccl'):
            raise ValueError("transport must be 'p2p' or 'nccl'")
        self.transport = transport
        shape = (self.chunk, self.T, self.T, 3)
        if transport == 'p2p':
            self.peers = PeerBuffers(2 * self.chunk * self.crop_bytes, device=self.device, group=group)
            self._slots = [[v[b * self.chunk * self.crop_bytes:(b + 1) * self.chunk * self.crop_bytes].view(shape)
                            for v in self.peers.views] for b in range(2)]
        else:
            self.peers = None
            self._slots = [[torch.empty(shape, dtype=torch.uint8, device=self.device)] for _ in range(2)]
            self._staging = (torch.empty((self.world,) + shape, dtype=torch.uint8, device=self.device)
                             if self.rank == self.root else None)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.out = (torch.empty((self.world * self.chunk, 3, self.T, self.T), dtype=torch.float32, device=self.device)
                    if self.rank == self.root else None)
        self.lut = batched.normalise_lut(self.device)

    def produce(self, i: int, fn) -> None:
        """Run ``fn(slot(i))`` -- the uint8 crop launch of chunk i -- where it overlaps best.  On the producers that is the
        current stream (they have nothing else to do).  On the root it is a side stream: the root's own production of chunk
        i + 1 (issue-bound) then runs under the conversion of chunk i (bound by NVLink ingest and HBM writes) instead of in
        front of it; ``collect(i)`` waits for it, and the slot is not overwritten before the conversion that last read it
        (chunk i - 2) has finished."""
        if self.rank != self.root:
            fn(self.slot(i))
            return
        if not hasattr(self, '_side'):
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_prod = [torch.cuda.Event(), torch.cuda.Event()]
            self._ev_conv = [None, None]
        main = torch.cuda.current_stream(self.device)
        b = i & 1
        with torch.cuda.stream(self._side):
            if self._ev_conv[b] is not None:
                self._side.wait_event(self._ev_conv[b])            # conversion of chunk i - 2 has read slot(i)
            else:
                self._side.wait_stream(main)                       # first use: whatever prepared the inputs
            fn(self.slot(i))
            self._ev_prod[b].record(self._side)
        self._pending = b

    def slot(self, i: int) -> torch.Tensor:
        """This rank's uint8 [chunk_rois,T,T,3] buffer for chunk i."""
        views = self._slots[i & 1]
        return views[self.rank] if self.transport == 'p2p' else views[0]

    def collect(self, i: int, counts=None):
        counts = [self.chunk] * self.world if counts is None else [int(c) for c in counts]
        if len(counts) != self.world or any(not 0 <= c <= self.chunk for c in counts):
            raise ValueError('counts must hold one value in [0, chunk_rois] per rank')
        def wait_own():                                                # the root's own chunk was produced on the side stream
            if getattr(self, '_pending', None) is not None:
                torch.cuda.current_stream(self.device).wait_event(self._ev_prod[self._pending])
        if self.transport == 'p2p':
            dist.all_reduce(self._flag, group=self.group)              # stream-ordered rendezvous, 4 bytes
            if self.rank != self.root:
                return None
            wait_own()
            srcs = [self._slots[i & 1][r][:counts[r]] for r in range(self.world)]
        else:
            wait_own()
            mine = self.slot(i)
            if self.rank == self.root:
                dist.gather(mine, list(self._staging.unbind(0)), dst=self.root, group=self.group)
                srcs = [self._staging[r, :counts[r]] for r in range(self.world)]
            else:
                dist.gather(mine, None, dst=self.root, group=self.group)
                return None
        res = self._batched.crops_normalise(srcs, self.T, swap_rb=self.swap_rb, lut=self.lut, out=self.out,
                                            device=self.device)[:sum(counts)]
        if getattr(self, '_pending', None) is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._ev_conv[self._pending] = ev
            self._pending = None
        return res

    def wire_bytes(self, counts=None) -> int:
        """Bytes that cross NVLink into the root for one chunk (the root's own crops do not travel)."""
        counts = [self.chunk] * self.world if counts is None else counts
        return int(sum(c for r, c in enumerate(counts) if r != self.root) * self.crop_bytes)

    def close(self):
        self._slots = []
        if self.peers is not None:
            self.peers.close()
