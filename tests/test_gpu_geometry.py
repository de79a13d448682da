"""CUDA geometry path (through the C ABI) against the reference's own outputs (tests/golden) and the oracle.

Bars (BASELINE.json north_star): match indices bit-exact including order; triangulated points within
1e-4 relative (norm-wise; we assert 1e-9); float32 costs bit-equal; F within 1e-12 relative.
"""
import numpy as np
import pytest
import torch

from oracle import geometry as og
from tests.gpu_util import batch_to_dev, pack_scenes, rel_err, to_dev

pytestmark = pytest.mark.gpu


def _run(Ks, RTs, centers, counts, threshold=30):
    from bpc_baseline_b200 import batched
    res = batched.match_triangulate(to_dev(Ks), to_dev(RTs), to_dev(centers), to_dev(counts), threshold,
                                    want_reproj=True, want_F=True)
    torch.cuda.synchronize()
    return {k: getattr(res, k).cpu().numpy() for k in ('idx', 'n', 'cost', 'X', 'reproj', 'F')}


def test_golden_scenes_full_path(golden_scenes):
    scenes = [golden_scenes.scene(n) for n in golden_scenes.names]
    Ks, RTs, centers, boxes, counts = pack_scenes(scenes)
    out = _run(Ks, RTs, centers, counts)
    for s, sc in enumerate(scenes):
        ref = sc['ref']
        n = int(out['n'][s])
        assert n == len(ref['idx']), sc['name']
        assert np.array_equal(out['idx'][s, :n], ref['idx']), sc['name']
        assert np.all(out['idx'][s, n:] == -1)
        np.testing.assert_allclose(out['F'][s], ref['F'], rtol=1e-12, atol=0, err_msg=sc['name'])
        if n:
            want_cost = ref['cost'][tuple(ref['idx'].T)]
            assert np.array_equal(out['cost'][s, :n].view(np.uint32), want_cost.view(np.uint32)), sc['name']
            assert rel_err(out['X'][s, :n], ref['X']).max() < 1e-9, sc['name']
            np.testing.assert_allclose(out['reproj'][s, :n], ref['reproj'], rtol=1e-6, atol=1e-7)


def test_bop_scene(golden_bop):
    g = golden_bop
    out = _run(g['Ks'][None], g['RTs'][None], g['centers'][None], g['counts'][None])
    n = int(out['n'][0])
    assert n == 10 and np.array_equal(out['idx'][0, :n], g['ref_idx'])
    assert rel_err(out['X'][0, :n], g['ref_X']).max() < 1e-9


def test_fundamental_and_cost_tensor_kernels(golden_scenes):
    from bpc_baseline_b200 import batched
    scenes = [golden_scenes.scene(n) for n in golden_scenes.names if not n.startswith('dense')]
    Ks, RTs, centers, boxes, counts = pack_scenes(scenes)
    F = batched.fundamental(to_dev(Ks), to_dev(RTs))
    cost = batched.cost_tensor(F, to_dev(centers), to_dev(counts)).cpu().numpy()
    F = F.cpu().numpy()
    nbits = 0
    for s, sc in enumerate(scenes):
        ref = sc['ref']
        np.testing.assert_allclose(F[s], ref['F'], rtol=1e-12, atol=0)
        N, M, P = sc['counts']
        if min(N, M, P) == 0:
            continue
        got = cost[s, :N, :M, :P]
        nbits += int((got.view(np.uint32) != ref['cost'].view(np.uint32)).sum())
        np.testing.assert_allclose(got, ref['cost'], rtol=2e-7, atol=0)
    assert nbits == 0, f'{nbits} cost elements differ from the reference in the last float32 bit'


def test_match_objects_explicit_cost(golden_scenes):
    """match_objects drop-in on the reference's own cost tensors (ascending-r order, no sort)."""
    from bpc_baseline_b200 import batched
    for name in golden_scenes.names:
        sc = golden_scenes.scene(name)
        cost = sc['ref']['cost']
        if cost.size == 0:
            continue
        idx, n = batched.match_objects(to_dev(cost[None]), 30)
        n = int(n.cpu()[0])
        want = og.match_objects(cost, 30)
        assert n == len(want), name
        assert [tuple(r) for r in idx.cpu().numpy()[0, :n]] == want, name


def test_match_objects_ties_vs_scipy():
    """Heavy exact ties / sentinel costs: the optimum returned must be SciPy's."""
    from bpc_baseline_b200 import batched
    rng = np.random.default_rng(5)
    for t in range(120):
        N, M, P = rng.integers(1, 7, 3)
        mode = t % 3
        if mode == 0:
            cost = rng.integers(0, 4, (N, M, P)).astype(np.float32)
        elif mode == 1:
            cost = rng.integers(0, 3, (N, M, P)).astype(np.float32)
            cost[rng.random((N, M, P)) < 0.3] = 9999
        else:
            cost = (rng.random((N, M, P)) * 40).astype(np.float32)
            if N > 1:
                cost[1] = cost[0]
        idx, n = batched.match_objects(to_dev(cost[None]), 30)
        n = int(n.cpu()[0])
        want = og.match_objects(cost, 30)
        assert [tuple(r) for r in idx.cpu().numpy()[0, :n]] == want, (t, cost.shape)


@pytest.mark.parametrize('D,kw,S', [(20, dict(), 96), (14, dict(p_drop=0.25, sigma=2.0), 96),
                                    (9, dict(n_dup=2, p_drop=0.1), 64), (10, dict(n_false=3, p_drop=0.2), 64),
                                    (4, dict(p_drop=0.35), 96)])
def test_against_oracle_seeded(D, kw, S):
    """Same seeded scenes through the CUDA path and the oracle (oracle = reference restatement + SciPy)."""
    from bpc_baseline_b200 import batched, synth
    batch = synth.make_scenes(S, D, seed=synth.SEED + 11, **kw)
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    res = batched.match_triangulate(Ks, RTs, centers, counts, 30)
    idx = res.idx.cpu().numpy(); n = res.n.cpu().numpy(); X = res.X.cpu().numpy(); cost = res.cost.cpu().numpy()
    nconf = 0
    for s in range(S):
        Kl, RTl = batch.capture_arrays(s)
        cen = [batch.centers[s, c, :batch.counts[s, c]] for c in range(3)]
        want = og.match_scene(Kl, RTl, cen, 30)
        assert int(n[s]) == len(want['idx']), s
        assert np.array_equal(idx[s, :n[s]], want['idx']), s
        if n[s]:
            assert np.array_equal(cost[s, :n[s]].view(np.uint32), want['cost'].view(np.uint32)), s
            assert rel_err(X[s, :n[s]], want['X']).max() < 1e-9, s
            nconf += len({(i, j) for i, j, _ in want['idx']}) < len(want['idx'])
    print(f'D={D} {kw}: {int(n.sum())} matches over {S} scenes')


def test_dense_scene_vs_oracle():
    """Dense bin (large virtual cost tensor): D=60 with dropped detections, LSAP conflicts exercised."""
    from bpc_baseline_b200 import batched, synth
    batch = synth.make_scenes(6, 60, seed=synth.SEED + 12, p_drop=0.1, sigma=2.0)
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    res = batched.match_triangulate(Ks, RTs, centers, counts, 30)
    idx = res.idx.cpu().numpy(); n = res.n.cpu().numpy()
    for s in range(6):
        Kl, RTl = batch.capture_arrays(s)
        cen = [batch.centers[s, c, :batch.counts[s, c]] for c in range(3)]
        want = og.match_scene(Kl, RTl, cen, 30)
        assert int(n[s]) == len(want['idx'])
        assert np.array_equal(idx[s, :n[s]], want['idx'])


def test_zero_detections_and_empty_batch():
    from bpc_baseline_b200 import batched, synth
    batch = synth.make_scenes(4, 5, seed=synth.SEED + 13)
    batch.counts[1, 2] = 0
    batch.counts[3, 0] = 0
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    res = batched.match_triangulate(Ks, RTs, centers, counts, 30)
    n = res.n.cpu().numpy()
    assert n[1] == 0 and n[3] == 0 and n[0] > 0 and n[2] > 0          # process_pose.py:161-163
    assert torch.all(res.idx[1] == -1)
    empty = batched.match_triangulate(Ks[:0], RTs[:0], centers[:0], counts[:0], 30)
    assert empty.n.numel() == 0


def test_triangulate_and_reprojection_kernels(golden_scenes):
    from bpc_baseline_b200 import batched
    sc = golden_scenes.scene('clean20_0')
    Ps = np.stack(og.projection_matrices(sc['Ks'], sc['RTs']))
    ref = sc['ref']
    n = len(ref['idx'])
    P = np.broadcast_to(Ps, (n, 3, 3, 4)).copy()
    X = batched.triangulate(to_dev(P), to_dev(ref['centroids']))
    err = batched.reprojection_error(to_dev(P), X, to_dev(ref['centroids']))
    assert rel_err(X.cpu().numpy(), ref['X']).max() < 1e-9
    np.testing.assert_allclose(err.cpu().numpy(), ref['reproj'], rtol=1e-6, atol=1e-7)


def test_rejects_cpu_tensors_and_missing_library():
    from bpc_baseline_b200 import batched
    with pytest.raises(RuntimeError):
        batched.fundamental(torch.zeros(1, 3, 3, 3), torch.zeros(1, 3, 4, 4, dtype=torch.float64))


def test_skewed_intrinsics_golden(golden_skew):
    """K with skew or a general 3x3 (tests/golden/skew.npz, recorded from the reference): np.linalg.inv computes in
    float64 and rounds to float32, and so does the kernel -- F to 1e-12, float32 costs bit-equal, same matches."""
    from bpc_baseline_b200 import batched
    scenes = [golden_skew.scene(n) for n in golden_skew.names]
    Ks, RTs, centers, boxes, counts = pack_scenes(scenes)
    out = _run(Ks, RTs, centers, counts)
    cost = batched.cost_tensor(to_dev(out['F']), to_dev(centers), to_dev(counts)).cpu().numpy()
    for s, sc in enumerate(scenes):
        ref = sc['ref']
        n = int(out['n'][s])
        np.testing.assert_allclose(out['F'][s], ref['F'], rtol=1e-12, atol=0, err_msg=sc['name'])
        assert n == len(ref['idx']) and np.array_equal(out['idx'][s, :n], ref['idx']), sc['name']
        N, M, P = sc['counts']
        assert np.array_equal(cost[s, :N, :M, :P].view(np.uint32), ref['cost'].view(np.uint32)), sc['name']
        assert rel_err(out['X'][s, :n], ref['X']).max() < 1e-9, sc['name']


def test_skewed_intrinsics_general_inverse():
    """Many more skewed scenes against the oracle (which calls np.linalg.inv like the reference)."""
    from bpc_baseline_b200 import batched, synth
    batch = synth.make_scenes(16, 8, seed=synth.SEED + 61)
    batch.Ks[:, :, 0, 1] = np.float32(3.5)                     # skew
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    res = batched.match_triangulate(Ks, RTs, centers, counts, 30, want_F=True)
    F = res.F.cpu().numpy(); idx = res.idx.cpu().numpy(); n = res.n.cpu().numpy()
    for s in range(16):
        Kl, RTl = batch.capture_arrays(s)
        want = og.match_scene(Kl, RTl, [batch.centers[s, c, :batch.counts[s, c]] for c in range(3)], 30)
        np.testing.assert_allclose(F[s], want['F'], rtol=1e-12, atol=0)
        assert int(n[s]) == len(want['idx']) and np.array_equal(idx[s, :n[s]], want['idx'])


def test_reprojection_filter(golden_scenes):
    """Optional reproj_thresh (a9): off = the reference's _match; on = drop a match when any view's
    compute_reprojection_error (utils/triangulation.py:14-18) exceeds it, order of the survivors kept."""
    from bpc_baseline_b200 import batched
    scenes = [golden_scenes.scene(n) for n in golden_scenes.names]
    Ks, RTs, centers, boxes, counts = pack_scenes(scenes)
    args = (to_dev(Ks), to_dev(RTs), to_dev(centers), to_dev(counts), 30)
    base = batched.match_triangulate(*args)
    allrep = np.concatenate([sc['ref']['reproj'].reshape(-1, 3) for sc in scenes])
    thr = float(np.median(allrep.max(axis=1)))                     # drops about half of the matches
    for t, expect_all in ((1e9, True), (thr, False), (0.0, False)):
        res = batched.match_triangulate(*args, reproj_thresh=t)
        dropped = 0
        for s, sc in enumerate(scenes):
            ref = sc['ref']
            keep = ~(ref['reproj'] > t).any(axis=1) if len(ref['idx']) else np.zeros(0, bool)
            n = int(res.n[s])
            assert n == int(keep.sum()), (sc['name'], t)
            assert np.array_equal(res.idx[s, :n].cpu().numpy(), ref['idx'][keep])
            assert np.all(res.idx[s, n:].cpu().numpy() == -1) and np.isnan(res.X[s, n:].cpu().numpy()).all()
            if n:
                want_cost = ref['cost'][tuple(ref['idx'][keep].T)]
                assert np.array_equal(res.cost[s, :n].cpu().numpy().view(np.uint32), want_cost.view(np.uint32))
                assert rel_err(res.X[s, :n].cpu().numpy(), ref['X'][keep]).max() < 1e-9
                np.testing.assert_allclose(res.reproj[s, :n].cpu().numpy(), ref['reproj'][keep], rtol=1e-6, atol=1e-7)
            dropped += len(ref['idx']) - n
        if expect_all:
            assert dropped == 0 and torch.equal(res.idx, base.idx) and torch.equal(res.n, base.n)
        else:
            assert dropped > 0


def test_counts_above_dmax_are_reported_not_dropped():
    """A scene whose count exceeds Dmax (the overflow count of bpc_detections_from_yolo) gets n = -2, not a silent 0."""
    from bpc_baseline_b200 import batched, synth
    batch = synth.make_scenes(4, 6, seed=synth.SEED + 63)
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    counts = counts.clone()
    counts[2, 1] = 7
    res = batched.match_triangulate(Ks, RTs, centers, counts, 30)
    n = res.n.cpu().numpy()
    assert n[2] == batched.N_OVERFLOW and (n[[0, 1, 3]] == 6).all()
    with pytest.raises(RuntimeError, match='exceed Dmax'):
        batched.check_match_status(res.n)


def test_large_dmax_uses_the_workspace():
    """Dmax beyond what fits shared memory (~550 with one warp per scene): the scene state moves to the caller's workspace;
    same results."""
    from bpc_baseline_b200 import _lib, batched, synth
    D = 24
    batch = synth.make_scenes(3, D, seed=synth.SEED + 64, p_drop=0.1, sigma=2.0)
    Dbig = 640
    assert _lib.load().bpc_match_workspace_bytes(3, 520) == 0 and _lib.load().bpc_match_workspace_bytes(3, Dbig) > 0
    centers = np.zeros((3, 3, Dbig, 2)); centers[:, :, :batch.centers.shape[2]] = batch.centers
    res = batched.match_triangulate(to_dev(batch.Ks), to_dev(batch.RTs), to_dev(centers), to_dev(batch.counts), 30)
    small = batched.match_triangulate(*batch_to_dev(batch)[:3], to_dev(batch.counts), 30)
    n = small.n.cpu().numpy()
    assert np.array_equal(res.n.cpu().numpy(), n)
    for s in range(3):
        assert torch.equal(res.idx[s, :n[s]], small.idx[s, :n[s]])
        assert torch.equal(res.X[s, :n[s]], small.X[s, :n[s]]) and torch.equal(res.cost[s, :n[s]], small.cost[s, :n[s]])
        Kl, RTl = batch.capture_arrays(s)
        want = og.match_scene(Kl, RTl, [batch.centers[s, c, :batch.counts[s, c]] for c in range(3)], 30)
        assert np.array_equal(res.idx[s, :n[s]].cpu().numpy(), want['idx'])


def test_pack_records_kernel_equals_layout_spec(golden_scenes):
    """bpc_pack_records against the torch statement of the layout, byte for byte, and the unpack round trip."""
    from bpc_baseline_b200 import batched, distributed
    scenes = [golden_scenes.scene(n) for n in golden_scenes.names]
    Ks, RTs, centers, boxes, counts = pack_scenes(scenes)
    res = batched.match_triangulate(to_dev(Ks), to_dev(RTs), to_dev(centers), to_dev(counts), 30)
    ios = to_dev(np.zeros((len(scenes), 3), np.int32))
    _, offs = batched.build_rois(to_dev(boxes), res.idx, res.n, ios)
    buf = distributed.pack_records(res, offs, 3)
    want = distributed.pack_records_torch(res.idx, res.n, res.cost, res.X, res.reproj)
    total = int(res.n.clamp(min=0).sum())
    used = distributed.records_bytes(len(scenes), res.idx.shape[1]) - (len(scenes) * res.idx.shape[1] - total) * 64
    assert torch.equal(buf[:used], want[:used])
    got = distributed.unpack_records(buf)
    assert torch.equal(got['idx'], res.idx) and torch.equal(got['n'], res.n)
    assert torch.equal(got['X'].view(torch.int64), res.X.view(torch.int64))
    assert torch.equal(got['reproj'].view(torch.int64), res.reproj.view(torch.int64))
    assert torch.equal(got['cost'].view(torch.int32), res.cost.view(torch.int32))
    assert torch.equal(distributed.unpack_records(distributed.pack_records(res))['idx'], res.idx)     # offsets computed inside


@pytest.mark.parametrize('threshold', [0.5, 2.25, 12.5, 1e9])
def test_threshold_is_compared_in_float32(threshold):
    """`val < threshold` on float32 costs (epipolar_matching.py:110-111; NumPy >= 2 compares in float32)."""
    from bpc_baseline_b200 import batched, synth
    batch = synth.make_scenes(24, 9, seed=synth.SEED + 62, p_drop=0.2, sigma=3.0)
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    res = batched.match_triangulate(Ks, RTs, centers, counts, threshold)
    idx = res.idx.cpu().numpy(); n = res.n.cpu().numpy()
    for s in range(24):
        Kl, RTl = batch.capture_arrays(s)
        want = og.match_scene(Kl, RTl, [batch.centers[s, c, :batch.counts[s, c]] for c in range(3)], threshold)
        assert int(n[s]) == len(want['idx']) and np.array_equal(idx[s, :n[s]], want['idx'])
