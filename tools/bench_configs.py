#!/usr/bin/env python
"""All five BASELINE.json configurations on one GPU (bench.py itself times config 2 only).

    python tools/bench_configs.py [--quick] > profiles/rNN_configs.json

Prints one JSON object per configuration.  Timings use CUDA events on the launching stream after warm-up.
"""
from __future__ import annotations

import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bpc_baseline_b200 import batched, pipeline, synth  # noqa: E402

PEAK = 6562.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()


def timed(fn, warm=3, reps=5):
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


def config1():
    """Single IPD-style scene, 3 cameras x 10 detections, through the reference's Python call surface."""
    from bpc_baseline_b200.inference.process_pose import PoseEstimator, PoseEstimatorParams
    batch = synth.make_scenes(1, 10, seed=synth.SEED + 2)
    images = synth.make_images(3, seed=3)
    Ks, RTs = batch.capture_arrays(0)
    cap = SimpleNamespace(images=[images[0], images[1], images[2]], Ks=Ks, RTs=RTs)
    est = PoseEstimator(PoseEstimatorParams(target_size=224))
    dets = batch.detections(0)
    for _ in range(3):
        preds = est._match(cap, dets); est.crop_inputs(preds); torch.cuda.synchronize()
    t0 = time.perf_counter(); preds = est._match(cap, dets); torch.cuda.synchronize(); t1 = time.perf_counter()
    tens = est.crop_inputs(preds); torch.cuda.synchronize(); t2 = time.perf_counter()
    # the images travel host -> device inside crop_inputs (3 x 24.9 MB), as the reference holds them on the host
    return {'config': 1, 'what': 'single scene, obj 8, 3 x 10 detections, host API incl. H2D of 3 full-res images',
            'matches': len(preds), 'crops': int(tens.shape[0]), 'match_ms': (t1 - t0) * 1e3, 'crop_ms': (t2 - t1) * 1e3}


def geometry(S, D, **kw):
    batch = synth.make_scenes(S, D, **kw)
    Ks, RTs, cen, cnt = dev(batch.Ks), dev(batch.RTs), dev(batch.centers), dev(batch.counts)
    ms = timed(lambda: batched.match_triangulate(Ks, RTs, cen, cnt, 30))
    res = batched.match_triangulate(Ks, RTs, cen, cnt, 30)
    n = int(res.n.clamp(min=0).sum())
    elems = float((batch.counts.astype(np.float64).prod(axis=1)).sum())
    return {'scenes': S, 'dets': D, **kw, 'ms': ms, 'scenes_per_s': S / ms * 1e3, 'matches': n,
            'virtual_cost_elements_per_s': elems / ms * 1e3}


def config4(lo, hi, T, R, B=8):
    images = dev(synth.make_images(B * 3, seed=44))
    rng = np.random.default_rng([44, lo, hi, T])
    w = rng.integers(lo, hi, R); h = rng.integers(lo, hi, R)
    x1 = (rng.random(R) * (synth.IMG_W - w)).astype(np.int64); y1 = (rng.random(R) * (synth.IMG_H - h)).astype(np.int64)
    rois = np.stack([rng.integers(0, B * 3, R), x1, y1, x1 + w, y1 + h], axis=1).astype(np.int32)
    chunk = 16384
    out = torch.empty((chunk, 3, T, T), dtype=torch.float32, device='cuda')
    drois = dev(rois)

    def run():
        for first in range(0, R, chunk):
            r = min(chunk, R - first)
            batched.roi_crop(images, drois[first:first + r], T=T, out=out)
    ms = timed(run, warm=2, reps=3)
    nbytes = pipeline.algorithmic_crop_bytes(rois, T)
    return {'config': 4, 'sides': [lo, hi], 'T': T, 'rois': R, 'ms': ms, 'crops_per_s': R / ms * 1e3,
            'algorithmic_GBps': nbytes / ms / 1e6, 'roofline_frac_of_measured': nbytes / ms / 1e6 / PEAK}


def config5(S=131072, D=20, T=224, pool=8):
    batch = synth.make_scenes(S, D)
    images = dev(synth.make_images(pool * 3))
    ios = dev(((np.arange(S)[:, None] % pool) * 3 + np.arange(3)[None, :]).astype(np.int32))
    Ks, RTs, cen, cnt, boxes = dev(batch.Ks), dev(batch.RTs), dev(batch.centers), dev(batch.counts), dev(batch.boxes)
    pipe = pipeline.MatchCropPipeline(S, D, T=T, chunk_rois=16384)
    res, offs = pipe.run_device(Ks, RTs, cen, cnt, boxes, images, ios)
    n_rois = int(offs[-1])
    ms = timed(lambda: pipe.run_device(Ks, RTs, cen, cnt, boxes, images, ios, n_rois_host=n_rois), warm=1, reps=2)
    return {'config': 5, 'scenes': S, 'dets': D, 'rois': n_rois, 'ms': ms, 'scenes_per_s': S / ms * 1e3, 'crops_per_s': n_rois / ms * 1e3}


def main():
    quick = '--quick' in sys.argv
    out = [config1()]
    out.append({'config': 2, 'geometry_only': geometry(4096, 20)})
    out.append({'config': 3, 'geometry_only': geometry(2048 if quick else 16384, 200)})
    out.append({'config': '3 (stress: p_drop 0.1, sigma 2 px)', 'geometry_only': geometry(512 if quick else 2048, 200, p_drop=0.1, sigma=2.0)})
    for (lo, hi) in ((32, 96), (60, 400), (300, 900)):
        for T in (224, 256):
            out.append(config4(lo, hi, T, 1 << (14 if quick else 16)))
    out.append(config5(S=16384 if quick else 131072))
    for o in out:
        print(json.dumps(o))


if __name__ == '__main__':
    main()
