"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes shard a scene stream, pack their
pose records, gather them, and every rank must see exactly what a single process would have produced."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bpc_baseline_b200 import distributed, synth


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_results(lo, hi, K):
    """Deterministic stand-in for the matcher outputs of scenes [lo, hi) (the CUDA kernels need a GPU)."""
    S = hi - lo
    sid = torch.arange(lo, hi)
    n = (sid % (K + 1)).to(torch.int32)
    idx = torch.full((S, K, 3), -1, dtype=torch.int32)
    cost = torch.full((S, K), float('nan'), dtype=torch.float32)
    X = torch.full((S, K, 3), float('nan'), dtype=torch.float64)
    for s in range(S):
        for m in range(int(n[s])):
            idx[s, m] = torch.tensor([m, (m + int(sid[s])) % K, (2 * m) % K])
            cost[s, m] = float(np.float32(0.37 * (m + 1) + 1e-3 * int(sid[s])))
            X[s, m] = torch.tensor([1.0 / 3.0 + m, -2.5 * int(sid[s]), 7.0e-5 * m])
    reproj = (X.abs() * 0.125 + 0.5)
    return idx, n, cost, X, reproj


def _worker(rank, world, port, total, K, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = distributed.shard_range(total, rank, world)
        rec = distributed.pack_records_torch(*_fake_results(lo, hi, K))
        allrec = distributed.gather_records(rec)
        got = distributed.unpack_records(allrec)
        torch.save(got, os.path.join(out_dir, f'rank{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_and_align():
    for total, world, align in [(4096, 8, 256), (1000, 3, 1), (131072, 8, 256), (5, 8, 1), (700, 2, 256)]:
        spans = [distributed.shard_range(total, r, world, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        assert all(a % align == 0 or a == total for a, _ in spans)
    # a rank's shard of the synthetic stream equals the same range of a single-process stream
    lo, hi = distributed.shard_range(1024, 1, 2, synth.CHUNK)
    whole = synth.make_scenes(1024, 5)
    part = synth.make_scenes(hi - lo, 5, first=lo)
    assert np.array_equal(whole.boxes[lo:hi], part.boxes) and np.array_equal(whole.RTs[lo:hi], part.RTs)


def test_pack_unpack_roundtrip_is_exact():
    idx, n, cost, X, reproj = _fake_results(10, 42, 6)
    n[3] = -1                                               # a status value travels as it is; the scene has no records
    buf = distributed.pack_records_torch(idx, n, cost, X, reproj)
    assert buf.numel() == distributed.records_bytes(32, 6)
    got = distributed.unpack_records(buf)
    idx[3], cost[3], X[3], reproj[3] = -1, float('nan'), float('nan'), float('nan')
    assert torch.equal(got['idx'], idx) and torch.equal(got['n'], n)
    assert torch.equal(got['cost'].view(torch.int32), cost.view(torch.int32))
    assert torch.equal(got['X'].view(torch.int64), X.view(torch.int64))
    assert torch.equal(got['reproj'].view(torch.int64), reproj.view(torch.int64))


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_single_process(tmp_path):
    total, K, world = 64, 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, K, str(tmp_path)), nprocs=world, join=True)
    want = distributed.unpack_records(distributed.pack_records_torch(*_fake_results(0, total, K)))
    for rank in range(world):
        got = torch.load(os.path.join(str(tmp_path), f'rank{rank}.pt'))
        for key in ('idx', 'n'):
            assert torch.equal(got[key], want[key]), (rank, key)
        assert torch.equal(got['cost'].view(torch.int32), want['cost'].view(torch.int32))
        assert torch.equal(got['X'].view(torch.int64), want['X'].view(torch.int64))
        assert torch.equal(got['reproj'].view(torch.int64), want['reproj'].view(torch.int64))
