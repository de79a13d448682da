#!/usr/bin/env python
"""bench.py -- throughput of the match + triangulate + ROI-crop hot path on B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1       # CPU arm (oracle port, all host cores)

A step = one pass of the hot path over one batch of synthetic scenes per GPU (BASELINE.json config 2:
4,096 3-camera scenes x 20 detections): fundamental matrices, virtual epipolar cost tensor, SciPy-exact
assignment, DLT triangulation + reprojection error, and the 224x224 float32 network input of EVERY matched
detection in all three views (~245,760 crops, 148 GB of output per step, produced chunk-wise into a
reusable buffer).  Scenes are independent: with N ranks every rank processes its own 4,096 scenes (weak
scaling) and the pose records are all-gathered over NCCL at the end of each step; crops stay on the GPU that
made them (their consumer, the pose network, is data-parallel too).

One JSON line is printed by rank 0; see the keys in main().
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'matched+triangulated scenes/s with the ROI crops of every match (crops/s alongside)'
UNIT = 'scenes/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--scenes', type=int, default=4096, help='scenes per GPU per step')
    ap.add_argument('--dets', type=int, default=20, help='detections per camera')
    ap.add_argument('--target', type=int, default=224, help='crop side T')
    ap.add_argument('--pool', type=int, default=8, help='image pool: number of 3-view full-resolution triplets')
    ap.add_argument('--chunk-rois', type=int, default=16384, help='ROIs per crop launch (16384 x 602 KB = 9.9 GB chunk buffer)')
    ap.add_argument('--p-drop', type=float, default=0.0)
    ap.add_argument('--sigma', type=float, default=1.0)
    ap.add_argument('--no-crops', action='store_true', help='geometry only (dense-bin experiments)')
    ap.add_argument('--cpu-scenes', type=int, default=256, help='scenes in the bounded CPU-baseline sample (~12 s on one core)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-crop-gather', action='store_true', help='skip the separate crop-gather stage at N > 1')
    ap.add_argument('--gather-chunk', type=int, default=4096, help='uint8 crops per rank and chunk in the crop-gather stage')
    ap.add_argument('--gather-steps', type=int, default=8)
    return ap.parse_args()


def workload_config(args):
    name = ('config 2' if (args.scenes, args.dets) == (4096, 20) and not args.no_crops
            else 'config 3 (dense bin)' if args.dets == 200 else 'custom')
    path = ('geometry only (no crops)' if args.no_crops else
            f'full path incl. {args.target}x{args.target} float32 crops of every matched detection in 3 views')
    return {
        'workload': f'{name}: {args.scenes} synthetic 3-camera scenes x {args.dets} detections per GPU per step '
                    f'(sigma={args.sigma}px, p_drop={args.p_drop}); {path}',
        'scenes_per_gpu': args.scenes, 'detections_per_camera': args.dets, 'target_size': args.target,
        'image_pool': f'{args.pool} triplets of 3840x2160 BGR uint8 ({args.pool * 3 * 3840 * 2160 * 3 / 1e6:.0f} MB), '
                      'scene s uses triplet s % pool',
        'crop_chunk_rois': args.chunk_rois,
        'l2': 'inputs (image pool) and outputs (crop chunk buffer) are both larger than the 126 MB L2; no explicit flush',
        'threshold': 30,
    }


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived inside the timed region [t0, t1] (host clock); if the region was
        shorter than the sampling period, of the samples taken under load since warm-up began."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (ts, r) in self.rows if t0 is None or t0 <= ts <= t1 + 0.12]
        scope = 'timed region'
        if not rows:
            rows = [r for (ts, r) in self.rows[1:]]
            scope = 'warm-up + timed region (timed region shorter than the 100 ms sampling period)'
        for row in rows:
            parts = [p.strip() for p in row.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smmax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smmax)), 'power_w_max': float(max(power)),
                'samples': len(sm), 'scope': scope, 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy + the reference's own SciPy / OpenCV / torchvision calls)
# ---------------------------------------------------------------------------------------------------------
def _cpu_scene(batch, images, s, T, pool):
    """One scene through the oracle: _match restatement + the crop of every match in 3 views."""
    from oracle import crop as ocrop
    from oracle import geometry as og
    Ks, RTs = batch.capture_arrays(s)
    cen = [batch.centers[s, c, :batch.counts[s, c]] for c in range(3)]
    res = og.match_scene(Ks, RTs, cen, 30, cost_fn=og.cost_tensor_fast)
    crops = 0
    for (i, j, k) in res['idx']:
        for v, d in enumerate((i, j, k)):
            ocrop.crop_tensor_ref(images[(s % pool) * 3 + v], batch.boxes[s, v, d], target_size=T, swap_rb=True)
            crops += 1
    return len(res['idx']), crops


_W = {}


def _cpu_worker_init(scenes, dets, sigma, p_drop, pool, T):
    import cv2
    import torch
    from bpc_baseline_b200 import synth
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    _W['batch'] = synth.make_scenes(scenes, dets, sigma=sigma, p_drop=p_drop)
    _W['images'] = synth.make_images(pool * 3).reshape(pool * 3, synth.IMG_H, synth.IMG_W, 3)
    _W['T'], _W['pool'] = T, pool


def _cpu_worker_run(s):
    return _cpu_scene(_W['batch'], _W['images'], s, _W['T'], _W['pool'])


def cpu_baseline_single(args, batch, images):
    """Bounded single-thread sample of the same workload on this host (the `cpu_baseline` object)."""
    import cv2
    import torch
    cv2.setNumThreads(1)
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    n = min(args.cpu_scenes, len(batch))
    _cpu_scene(batch, images, 0, args.target, args.pool)          # warm caches / imports
    t0 = time.perf_counter()
    crops = 0
    for s in range(n):
        crops += _cpu_scene(batch, images, s, args.target, args.pool)[1]
    dt = time.perf_counter() - t0
    torch.set_num_threads(nthreads)
    return {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'crops_per_s': crops / dt,
            'sample': f'{n} of the {len(batch)} scenes ({crops} crops) through oracle.geometry.match_scene '
                      f'(vectorised bit-identical restatement of _match + scipy LSAP + numpy SVD) and oracle.crop.crop_tensor_ref '
                      f'(cv2.resize INTER_AREA + cvtColor + torchvision to_tensor/normalize), 1 thread, {dt:.1f} s; '
                      f'host has {len(os.sched_getaffinity(0))} cores'}


def run_reference(args):
    """--impl reference: the oracle port on all host cores (the reference is pure Python: there is no
    oracle/_ref binary; /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    per_step = max(cores * 4, 16)
    pool_imgs = min(args.pool, 2)                                   # image content does not change the CPU cost
    total = per_step * (args.steps + args.warmup)
    ctx = mp.get_context('fork')
    t_setup = time.perf_counter()
    with ctx.Pool(cores, initializer=_cpu_worker_init,
                  initargs=(min(total, 1024), args.dets, args.sigma, args.p_drop, pool_imgs, args.target)) as pool:
        nscn = min(total, 1024)
        pool.map(_cpu_worker_run, range(cores))                     # every worker initialised and warm
        for w in range(args.warmup):
            pool.map(_cpu_worker_run, [(w * per_step + i) % nscn for i in range(per_step)], chunksize=1)
        t0 = time.perf_counter()
        crops = 0
        for k in range(args.steps):
            res = pool.map(_cpu_worker_run, [((args.warmup + k) * per_step + i) % nscn for i in range(per_step)], chunksize=1)
            crops += sum(c for _, c in res)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = (f'{per_step} scenes per step ({crops // max(args.steps, 1)} crops) of the config-2 workload, fanned out over '
              f'{cores} processes (oracle port: NumPy restatement + scipy LSAP + cv2 INTER_AREA + torchvision normalize)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / max(args.steps, 1) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64 geometry / u8+f32 crops', 'data': 'synthetic', 'config': workload_config(args),
        'crops_per_s': crops / dt,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'setup_s': t0 - t_setup,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    from bpc_baseline_b200 import _lib, batched, distributed, pipeline, synth

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL writes its version banner / log lines to stdout (the level may come from /etc/nccl.conf, not the
        # environment); stdout carries the one JSON line, so send them to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()

    S, D, T = args.scenes, args.dets, args.target
    first = ((rank * S + synth.CHUNK - 1) // synth.CHUNK) * synth.CHUNK
    batch = synth.make_scenes(S, D, first=first, sigma=args.sigma, p_drop=args.p_drop)
    Dmax = batch.boxes.shape[2]
    H, W = synth.IMG_H, synth.IMG_W
    nimg = args.pool * 3
    # image pool: generated on the host once (also the e2e source), resident in HBM for the device-timed run
    images_h = synth.make_images(nimg)
    ios = ((np.arange(S)[:, None] % args.pool) * 3 + np.arange(3)[None, :]).astype(np.int32)
    todev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    Ks, RTs, centers, counts, boxes = (todev(batch.Ks), todev(batch.RTs), todev(batch.centers), todev(batch.counts), todev(batch.boxes))
    images = todev(images_h)
    ios_d = todev(ios)

    pipe = pipeline.MatchCropPipeline(S, Dmax, T=T, chunk_rois=args.chunk_rois, device=dev)

    # one untimed pass to learn the ROI count and the algorithmic bytes (identical inputs every step)
    res, offs = pipe.run_device(Ks, RTs, centers, counts, boxes, images, ios_d)
    torch.cuda.synchronize()
    n_rois = int(offs[-1].item())
    n_matches = int(res.n.clamp(min=0).sum().item())
    rois_h = pipe.rois[:n_rois].cpu().numpy()
    rejected = int(pipe.status[:n_rois].sum().item())
    crop_bytes = pipeline.algorithmic_crop_bytes(rois_h, T)
    n_rois_arg = 0 if args.no_crops else n_rois

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step(k=None):
        e = ev[k] if k is not None else None
        if e is not None:
            e[3].record()
        r, _ = pipe.run_device(Ks, RTs, centers, counts, boxes, images, ios_d, n_rois_host=n_rois_arg,
                               events=(e[0], e[1], e[2]) if e is not None else None)
        if world > 1:                     # final gather of the pose records over NVLink
            distributed.gather_records(distributed.pack_records(r.idx, r.n, r.cost, r.X))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.time()
    t_start.record()
    for k in range(args.steps):
        step(k)
    t_stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    launches = _lib.launch_count() - launches0
    ms = t_start.elapsed_time(t_stop)
    crop_ms = sum(e[1].elapsed_time(e[2]) for e in ev)
    match_ms = sum(e[3].elapsed_time(e[0]) for e in ev)
    crop_launches = pipe.last_crop_launches * args.steps
    if world > 1:
        t = torch.tensor([ms, crop_ms, match_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, crop_ms, match_ms = [float(v) for v in t.cpu()]

    # ---- end-to-end from pinned host buffers (H2D of all inputs incl. the image pool, D2H of pose records) ----
    e2e = None
    if not args.no_e2e and not args.no_crops:
        hb = pipe.host_buffers(images_h.shape)
        hb['Ks'].copy_(torch.from_numpy(batch.Ks)); hb['RTs'].copy_(torch.from_numpy(batch.RTs))
        hb['boxes'].copy_(torch.from_numpy(batch.boxes)); hb['counts'].copy_(torch.from_numpy(batch.counts))
        hb['image_of_scene'].copy_(torch.from_numpy(ios)); hb['images'].copy_(torch.from_numpy(images_h))
        out = pipe.run_host(n_rois_host=None)                       # serial variant: warm-up + check
        assert int(out['n_rois'][0]) == n_rois and bool((out['idx'] == res.idx.cpu()).all())
        pipe.run_host_stream(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # K steps, uploads double buffered behind the previous step's kernels (MatchCropPipeline.run_host_stream);
        # the wall clock brackets everything incl. the first upload and the last read-back
        t0 = time.perf_counter()
        out = pipe.run_host_stream(args.steps)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        assert int(out['n_rois'][0]) == n_rois and bool((out['idx'] == res.idx.cpu()).all())
        e2e = {'value': world * S * args.steps / (e2e_ms * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': pipe.h2d_bytes(), 'd2h_bytes_per_step': pipe.d2h_bytes(),
               'ms_per_step': e2e_ms / args.steps,
               'note': 'host wall clock over K steps; inputs (K, RT, boxes, counts, image pool) copied from pinned host memory '
                       'every step (double buffered behind the previous step), centres derived on device; pose records '
                       '(idx, n, cost, X) read back; crops stay in HBM for the on-device pose network '
                       '(the reference moves them H2D at process_pose.py:210)'}

    # ---- crop gather to rank 0 (SURVEY.md 8e options 2/3), its own stage with its own bound: uint8 crops pulled
    #      over NVLink by the receiver's normalise kernel; NOT part of `value` (crops stay sharded there) ----
    crop_gather = None
    if world > 1 and not args.no_crops and not args.no_crop_gather:
        try:
            gchunk = min(args.gather_chunk, n_rois)
            cg = distributed.CropGather(gchunk, T=T, root=0, transport='p2p', device=dev)
            g_rois = pipe.rois[:gchunk]
            for i in range(3):
                batched.roi_crop_u8(images, g_rois, T=T, out=cg.slot(i))
                cg.collect(i)
            torch.cuda.synchronize()
            dist.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for i in range(args.gather_steps):
                batched.roi_crop_u8(images, g_rois, T=T, out=cg.slot(i))
                cg.collect(i)
            g1.record()
            torch.cuda.synchronize()
            t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item()) * 1e-3 / args.gather_steps
            crop_gather = {'transport': 'uint8 crops pulled over NVLink peer memory by the receiver kernel (bpc_crops_normalise)',
                           'root': 0, 'chunk_rois_per_rank': gchunk, 'chunks': args.gather_steps,
                           'crops_per_s': world * gchunk / sec, 'ms_per_chunk': sec * 1e3,
                           'nvlink_ingest_GBps': cg.wire_bytes() / sec / 1e9, 'nvlink_peak_GBps': 900.0,
                           'root_hbm_write_GBps': world * gchunk * 3 * T * T * 4 / sec / 1e9,
                           'note': 'every rank produces a chunk of uint8 crops, rank 0 converts all of them to the float32 '
                                   'network input; separate from `value`, where crops stay on the GPU that produced them'}
            cg.close()
        except Exception as exc:              # noqa: BLE001 -- the headline numbers do not depend on this stage
            crop_gather = {'unavailable': f'{type(exc).__name__}: {exc}'[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    roofline = None
    if crop_launches and crop_ms > 0:
        per_launch_bytes = crop_bytes / pipe.last_crop_launches
        avg_launch_s = crop_ms * 1e-3 / crop_launches
        achieved = per_launch_bytes / avg_launch_s / 1e9
        # DRAM bytes of one full 16384-ROI launch of the default workload from the committed ncu --set full capture
        # (profiles/r01_crop_warp_kernel_ncu.txt: dram__bytes_read.sum 3.069 GB + dram__bytes_write.sum 9.835 GB)
        default_wl = (S, D, T, args.chunk_rois, args.pool, args.p_drop, args.sigma) == (4096, 20, 224, 16384, 8, 0.0, 1.0)
        roofline = {'bound': 'hbm', 'kernel': 'bpc_crop_warp_kernel<false,224,true,true> (+ prep, generic)', 'achieved': achieved,
                    'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': 12.905e9 if default_wl else None,
                    'traffic_source': 'ncu capture of a full 16384-ROI launch, bytes' if default_wl else None,
                    'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': per_launch_bytes, 'avg_launch_ms': avg_launch_s * 1e3,
                    'launches_per_step': pipe.last_crop_launches, 'share_of_step': crop_ms / ms}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_single(args, batch, images_h)

    value = world * S * args.steps / (ms * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 geometry / u8+f32 crops', 'data': 'synthetic', 'config': workload_config(args),
        'crops_per_s': world * n_rois * args.steps / (ms * 1e-3) if not args.no_crops else 0.0,
        'matches_per_step_per_gpu': n_matches, 'rois_per_step_per_gpu': n_rois, 'rois_rejected': rejected,
        'geometry_ms_per_step': match_ms / args.steps, 'crop_ms_per_step': crop_ms / args.steps,
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
    }
    if crop_gather is not None:
        line['crop_gather'] = crop_gather
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
