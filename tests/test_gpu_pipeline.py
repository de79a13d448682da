"""MatchCropPipeline: chunked crops, device-side ROI counts, host path and the double-buffered stream path."""
import numpy as np
import pytest
import torch

from oracle import crop as ocrop
from tests.gpu_util import to_dev

pytestmark = pytest.mark.gpu


def _setup(S=48, D=6, W=960, H=540, B=2, T=64, chunk=64):
    from bpc_baseline_b200 import pipeline, synth
    batch = synth.make_scenes(S, D, seed=synth.SEED + 51, p_drop=0.15, width=W, height=H, side_lo=20, side_hi=200)
    images = synth.make_images(B * 3, seed=52, width=W, height=H)
    ios = ((np.arange(S)[:, None] % B) * 3 + np.arange(3)[None, :]).astype(np.int32)
    pipe = pipeline.MatchCropPipeline(S, batch.boxes.shape[2], T=T, chunk_rois=chunk)
    return batch, images, ios, pipe


def test_run_device_chunks_cover_every_roi_once():
    from bpc_baseline_b200 import synth
    batch, images, ios, pipe = _setup()
    got = {}

    def consumer(chunk, first):
        got[first] = chunk.clone()
    res, offs = pipe.run_device(to_dev(batch.Ks), to_dev(batch.RTs), to_dev(batch.centers), to_dev(batch.counts),
                                to_dev(batch.boxes), to_dev(images), to_dev(ios), consumer=consumer)
    total = int(offs[-1].item())
    rois = synth.rois_for_matches(batch.boxes, res.idx.cpu().numpy(), res.n.cpu().numpy(), ios)
    assert total == len(rois) and total > pipe.chunk          # several chunks
    assert sorted(got) == list(range(0, pipe.cap, pipe.chunk))
    rng = np.random.default_rng(0)
    for g in rng.choice(total, 24, replace=False):
        first = (g // pipe.chunk) * pipe.chunk
        b, x1, y1, x2, y2 = rois[g]
        want = ocrop.crop_tensor_ref(images[b], (x1, y1, x2, y2), target_size=pipe.T)
        assert np.array_equal(got[first][g - first].cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert int(pipe.status[:total].sum()) == 0


def test_host_stream_equals_device_path_step_by_step():
    from bpc_baseline_b200 import batched
    batch, images, ios, pipe = _setup()
    hb = pipe.host_buffers(images.shape)
    steps = 5
    perms = [np.random.default_rng(k).permutation(len(batch)) for k in range(steps)]

    def refill(h, k):
        p = perms[k]
        h['Ks'].copy_(torch.from_numpy(batch.Ks[p])); h['RTs'].copy_(torch.from_numpy(batch.RTs[p]))
        h['boxes'].copy_(torch.from_numpy(batch.boxes[p])); h['counts'].copy_(torch.from_numpy(batch.counts[p]))
        h['image_of_scene'].copy_(torch.from_numpy(ios[p])); h['images'].copy_(torch.from_numpy(images))
    results = {}

    def on_result(r, k):
        results[k] = {key: r[key].clone() for key in ('idx', 'n', 'cost', 'X', 'n_rois')}
    pipe.run_host_stream(steps, refill=refill, on_result=on_result)
    assert sorted(results) == list(range(steps))
    for k in range(steps):
        p = perms[k]
        want = batched.match_triangulate(to_dev(batch.Ks[p]), to_dev(batch.RTs[p]), to_dev(batch.centers[p]),
                                         to_dev(batch.counts[p]), 30)
        assert torch.equal(results[k]['idx'], want.idx.cpu()) and torch.equal(results[k]['n'], want.n.cpu())
        assert torch.equal(results[k]['cost'].view(torch.int32), want.cost.cpu().view(torch.int32))
        assert int(results[k]['n_rois'][0]) == 3 * int(want.n.clamp(min=0).sum())
    # the serial host path gives the same answer
    refill(hb, 0)
    out = pipe.run_host()
    assert torch.equal(out['idx'], results[0]['idx'])


def test_pose_head_on_the_bf16_chunk_buffer_matches_the_per_crop_float32_loop():
    """f1: MatchCropPipeline(crop_dtype=bfloat16) -> PoseHeadConsumer (SimplePoseNet in bf16 channels-last, fed the chunk
    buffer as it stands) against the reference's flow -- float32 crops, the network called once per crop, decode per crop
    (process_pose.py:210-229).  Tolerance: 5 degrees of geodesic distance between the decoded rotations (measured 0.4 - 1.3 degrees
    over seeds for the 6d and quaternion heads: bf16 has 8 mantissa bits and a random-init ResNet50 amplifies input rounding; an
    euler head with random weights emits angles of tens of radians, where the same relative error is tens of degrees -- not used here)."""
    import numpy as np
    from bpc_baseline_b200 import batched, pipeline, synth
    from bpc_baseline_b200.inference.process_pose import decode_rotations
    from bpc_baseline_b200.pose.head import PoseHeadConsumer, geodesic_degrees
    from bpc_baseline_b200.pose.models.simple_pose_net import SimplePoseNet
    from tests.gpu_util import batch_to_dev, to_dev
    torch.manual_seed(0)
    S, D, T = 4, 6, 224
    batch = synth.make_scenes(S, D, seed=123)
    images = to_dev(synth.make_images(3, seed=3))
    Ks, RTs, centers, boxes, counts = batch_to_dev(batch)
    image_of_scene = to_dev(np.tile(np.arange(3, dtype=np.int32), (S, 1)))
    model = SimplePoseNet(loss_type='6d', pretrained=False).cuda().eval()
    ref_model = SimplePoseNet(loss_type='6d', pretrained=False).cuda().eval()
    ref_model.load_state_dict(model.state_dict())
    pipe16 = pipeline.MatchCropPipeline(S, D, T=T, chunk_rois=32, crop_dtype=torch.bfloat16)
    head = PoseHeadConsumer(model, pipe16.cap, sub_batch=16)
    res, offs = pipe16.run_device(Ks, RTs, centers, counts, boxes, images, image_of_scene, consumer=head)
    n = int(offs[S].item())
    assert n > 0 and n % 3 == 0
    got = head.rotations(n)
    # the reference flow on the float32 crops: batch 1, float32, decode per crop
    rois = pipe16.rois[:n]
    f32 = batched.roi_crop(images, rois, T=T)
    want = []
    with torch.no_grad():
        for r in range(n):
            raw = ref_model(f32[r:r + 1]).float().cpu()
            want.append(decode_rotations(raw, '6d')[0])
    want = torch.tensor(np.stack(want), device='cuda')
    deg = geodesic_degrees(got, want)
    assert float(deg.max()) < 5.0, float(deg.max())
    # and the rotations are rotations
    eye = torch.eye(3, device='cuda').expand(n, 3, 3)
    assert float((got @ got.transpose(1, 2) - eye).abs().max()) < 1e-3
