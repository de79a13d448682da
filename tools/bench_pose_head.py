#!/usr/bin/env python
"""f1 measurement (one GPU): can the pose head drain the crop kernel?

  1. crop kernel, float32 NCHW vs bfloat16 channels-last output: crops/s and the roofline fraction of EACH against its own
     algorithmic bytes (3 h w + 3 T^2 x {4, 2} + 20 per crop);
  2. SimplePoseNet (torchvision ResNet50, LIBRARY kernels) in bf16 channels-last on a resident chunk: crops/s;
  3. the full pass (match -> ROI list -> bf16 crops -> pose head -> rotation decode) through MatchCropPipeline(consumer=...).

    python tools/bench_pose_head.py [--scenes 1024] [--chunk 8192] [--sub-batch 2048]      -> one JSON object
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scenes', type=int, default=1024)
    ap.add_argument('--chunk', type=int, default=8192)
    ap.add_argument('--sub-batch', type=int, default=2048)
    ap.add_argument('--target', type=int, default=224)
    args = ap.parse_args()
    import numpy as np
    import torch
    from bpc_baseline_b200 import batched, pipeline, synth
    from bpc_baseline_b200.pose.head import PoseHeadConsumer
    from bpc_baseline_b200.pose.models.simple_pose_net import SimplePoseNet
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        peak = 6562.9
    dev = torch.device('cuda')
    S, D, T = args.scenes, 20, args.target
    batch = synth.make_scenes(S, D)
    images = torch.from_numpy(synth.make_images(8 * 3, seed=44)).to(dev)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Ks, RTs, centers, boxes, counts = to(batch.Ks), to(batch.RTs), to(batch.centers), to(batch.boxes), to(batch.counts)
    ios = to((np.arange(S)[:, None] % 8 * 3 + np.arange(3)[None, :]).astype(np.int32))

    def timed(fn, warm=2, reps=4):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {'workload': f'{S} scenes x {D} detections, T={T}, chunk {args.chunk} ROIs, pool of 8 triplets', 'peak_GBps': peak}
    # ---- 1. crop kernel per output type on one chunk of real ROIs
    pipe = pipeline.MatchCropPipeline(S, D, T=T, chunk_rois=args.chunk)
    res, offs = pipe.run_device(Ks, RTs, centers, counts, boxes, images, ios)
    n = int(offs[S].item())
    R = min(args.chunk, n)
    rois = pipe.rois[:R].contiguous()
    rois_h = rois.cpu().numpy()
    buf32 = pipe.crops[:R]
    buf16 = torch.empty((R, 3, T, T), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    for name, fn, elem in (('float32_nchw', lambda: batched.roi_crop(images, rois, T=T, out=buf32), 4),
                           ('bfloat16_nhwc', lambda: batched.roi_crop_bf16(images, rois, T=T, out=buf16), 2)):
        ms = timed(fn)
        nbytes = pipeline.algorithmic_crop_bytes(rois_h, T, elem)
        out['crop_' + name] = {'rois': R, 'ms': ms, 'crops_per_s': R / ms * 1e3, 'algorithmic_bytes': nbytes,
                               'roofline': {'bound': 'hbm', 'achieved': nbytes / ms / 1e6, 'peak': peak, 'unit': 'GB/s', 'frac': nbytes / ms / 1e6 / peak}}
    # ---- 2. the pose head alone on the resident bf16 chunk
    torch.manual_seed(0)
    head = PoseHeadConsumer(SimplePoseNet('6d', pretrained=False), S * D * 3, sub_batch=args.sub_batch)
    ms = timed(lambda: head(buf16, 0), warm=2, reps=3)
    out['pose_head_bf16_channels_last'] = {'crops': R, 'ms': ms, 'crops_per_s': R / ms * 1e3, 'sub_batch': args.sub_batch,
                                           'network': 'torchvision ResNet50 + Linear(2048, 6), random init, eval -- library kernels (cuDNN / cuBLAS)'}
    # ---- 3. the full pass with the head as the pipeline's consumer
    pipe16 = pipeline.MatchCropPipeline(S, D, T=T, chunk_rois=args.chunk, crop_dtype=torch.bfloat16)

    def full():
        pipe16.run_device(Ks, RTs, centers, counts, boxes, images, ios, consumer=head, n_rois_host=n)
        head.rotations(n)
    ms = timed(full, warm=1, reps=2)
    ms_nohead = timed(lambda: pipe16.run_device(Ks, RTs, centers, counts, boxes, images, ios, n_rois_host=n), warm=1, reps=2)
    out['full_pass'] = {'scenes': S, 'rois': n, 'ms_with_pose_head': ms, 'scenes_per_s_with_pose_head': S / ms * 1e3,
                        'crops_per_s_with_pose_head': n / ms * 1e3, 'ms_crops_only_bf16': ms_nohead,
                        'crops_per_s_crops_only_bf16': n / ms_nohead * 1e3,
                        'pose_head_share_of_pass': 1 - ms_nohead / ms}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
