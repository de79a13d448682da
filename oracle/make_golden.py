"""Oracle (test infrastructure): generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference, read-only):

    python -m oracle.make_golden

Imports the unmodified reference modules (pyrender / matplotlib / ultralytics are absent and
unused by the hot path, so empty stubs are put in sys.modules first -- SURVEY.md F8) and
records their outputs on seeded synthetic scenes:

  geometry.npz   per scene: inputs, compute_fundamental_matrix x3, compute_cost_matrix (the
                 as-written triple loop), match_objects + the sort of process_pose.py:183,
                 PoseEstimator._match (called unbound) -> boxes / centroids / t, and
                 compute_reprojection_error per view.
  bop_scene.npz  config 1: a temp BOP directory read back through Capture.from_dir so the
                 f32 -> f64 dtype flow is the real one, then _match.
  crops.npz      one source image, ROIs in all three INTER_AREA regimes, the reference's
                 letterbox_preserving_aspect_ratio outputs (uint8), the normalisation table from
                 torchvision's to_tensor/normalize, and full f32 tensors for a few crops.

  detect.npz     PoseEstimator._detect called unbound with a fake yolo callable (seeded boxes / conf / cls).
  skew.npz       scenes with skewed / general intrinsics (np.linalg.inv is then a real LU, not the pinhole form).

/root/reference does not exist on the GPU box: nothing else may import it.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE = '/root/reference'


def import_reference():
    for name in ('pyrender', 'matplotlib', 'matplotlib.pyplot', 'ultralytics'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['ultralytics'].YOLO = object
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import bpc.inference.epipolar_matching as em
    import bpc.inference.process_pose as pp
    import bpc.inference.utils.camera_utils as cu
    import bpc.inference.utils.triangulation as tri
    import bpc.utils.data_utils as du
    return SimpleNamespace(em=em, pp=pp, cu=cu, tri=tri, du=du)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_match(ref, Ks, RTs, dets, threshold=30):
    """Everything the reference computes for one scene, from its own functions."""
    K1, K2, K3 = Ks
    R = [x[:3, :3] for x in RTs]
    t = [x[:3, 3] for x in RTs]
    F12 = ref.cu.compute_fundamental_matrix(K1, R[0], t[0], K2, R[1], t[1])
    F13 = ref.cu.compute_fundamental_matrix(K1, R[0], t[0], K3, R[2], t[2])
    F23 = ref.cu.compute_fundamental_matrix(K2, R[1], t[1], K3, R[2], t[2])
    out = {'F': np.stack([F12, F13, F23])}
    params = ref.pp.PoseEstimatorParams(matching_threshold=threshold)
    capture = SimpleNamespace(images=[None] * 3, Ks=Ks, RTs=RTs)
    preds = quiet(ref.pp.PoseEstimator._match, SimpleNamespace(params=params), capture, dets)
    n = len(preds)
    out['boxes'] = np.array([p.boxes for p in preds], np.int64).reshape(n, 3, 4)
    out['centroids'] = np.array([p.centroids for p in preds], np.float64).reshape(n, 3, 2)
    out['X'] = np.array([p.t for p in preds], np.float64).reshape(n, 3)
    if min(len(dets[0]), len(dets[1]), len(dets[2])) == 0:
        out['cost'] = np.zeros((len(dets[0]), len(dets[1]), len(dets[2])), np.float32)
        out['idx'] = np.zeros((0, 3), np.int64)
        out['reproj'] = np.zeros((0, 3))
        return out
    cost = ref.em.compute_cost_matrix(dets[0], dets[1], dets[2], F12, F13, F23)
    matches = ref.em.match_objects(cost, threshold=threshold)
    matches = sorted(matches, key=lambda m: cost[m[0], m[1], m[2]])
    out['cost'] = cost
    out['idx'] = np.array(matches, np.int64).reshape(len(matches), 3)
    # _match and the standalone calls must agree (same code, same inputs)
    assert len(matches) == n
    for m, p in zip(matches, preds):
        for v in range(3):
            assert tuple(p.boxes[v]) == tuple(dets[v][m[v]]['bbox'])
    Ps = [K @ RT[:3] for K, RT in zip(Ks, RTs)]
    out['reproj'] = np.array([[ref.tri.compute_reprojection_error(Ps[v], p.t, p.centroids[v])
                               for v in range(3)] for p in preds]).reshape(n, 3)
    return out


def golden_geometry(ref):
    from bpc_baseline_b200 import synth
    cases = [
        # name, D, kwargs, scenes
        ('clean10', 10, dict(), 4),
        ('clean20', 20, dict(), 2),
        ('drop12', 12, dict(p_drop=0.25, sigma=2.0), 6),
        ('dup8', 8, dict(n_dup=2), 4),
        ('false9', 9, dict(n_false=3, p_drop=0.15), 4),
        ('tiny3', 3, dict(p_drop=0.3), 6),
        ('dense40', 40, dict(p_drop=0.1, sigma=2.0), 1),
    ]
    blob = {}
    names = []
    for name, D, kw, S in cases:
        batch = synth.make_scenes(S, D, seed=synth.SEED + 1, **kw)
        for s in range(S):
            tag = f'{name}_{s}'
            Ks, RTs = batch.capture_arrays(s)
            dets = batch.detections(s)
            if name == 'tiny3' and s == 5:          # a view with zero detections
                dets[1] = []
                batch.counts[s, 1] = 0
            res = run_match(ref, Ks, RTs, dets)
            names.append(tag)
            blob[f'{tag}/Ks'] = batch.Ks[s]
            blob[f'{tag}/RTs'] = batch.RTs[s]
            blob[f'{tag}/boxes'] = batch.boxes[s]
            blob[f'{tag}/centers'] = batch.centers[s]
            blob[f'{tag}/counts'] = batch.counts[s]
            for k, v in res.items():
                blob[f'{tag}/ref_{k}'] = v
            print(f'  {tag}: counts={batch.counts[s].tolist()} matches={len(res["idx"])}')
    blob['names'] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, 'geometry.npz'), **blob)


def golden_bop_scene(ref):
    """Config 1: single IPD-style scene, obj_id 8, 3 cameras x 10 detections, through Capture.from_dir."""
    import cv2
    from bpc_baseline_b200 import synth
    batch = synth.make_scenes(1, 10, seed=synth.SEED + 2)
    cam_ids = ['cam1', 'cam2', 'cam3']
    with tempfile.TemporaryDirectory() as d:
        for c, cid in enumerate(cam_ids):
            cam = {'0': {'cam_K': [float(v) for v in batch.Ks[0, c].reshape(-1)],
                         'cam_R_w2c': [float(v) for v in batch.RTs[0, c, :3, :3].reshape(-1)],
                         'cam_t_w2c': [float(v) for v in batch.RTs[0, c, :3, 3]],
                         'depth_scale': 1.0}}
            with open(os.path.join(d, f'scene_camera_{cid}.json'), 'w') as f:
                json.dump(cam, f)
            os.makedirs(os.path.join(d, f'rgb_{cid}'))
            cv2.imwrite(os.path.join(d, f'rgb_{cid}', '000000.png'), np.zeros((16, 16, 3), np.uint8))
        cap = quiet(ref.du.Capture.from_dir, d, cam_ids, 0, 8)
    assert all(k.dtype == np.float32 for k in cap.Ks) and all(rt.dtype == np.float64 for rt in cap.RTs)
    res = run_match(ref, cap.Ks, cap.RTs, batch.detections(0))
    blob = {'Ks': np.stack(cap.Ks), 'RTs': np.stack(cap.RTs), 'boxes': batch.boxes[0],
            'centers': batch.centers[0], 'counts': batch.counts[0]}
    blob.update({f'ref_{k}': v for k, v in res.items()})
    print(f'  bop scene: matches={len(res["idx"])}')
    np.savez_compressed(os.path.join(GOLDEN, 'bop_scene.npz'), **blob)


def golden_crops(ref):
    import cv2
    import torch
    import torchvision.transforms.functional as TF
    from bpc_baseline_b200 import synth
    H, W = 520, 648
    image = synth.make_images(1, seed=synth.SEED + 3, width=W, height=H)[0]
    rng = np.random.default_rng([synth.SEED, 3])
    boxes = []
    for _ in range(10):                                    # random sides, both regimes 1 and 3
        w, h = rng.integers(24, 420, 2)
        x1 = int(rng.integers(0, W - w + 1)); y1 = int(rng.integers(0, H - h + 1))
        boxes.append((x1, y1, x1 + int(w), y1 + int(h)))
    boxes += [(10, 20, 10 + 448, 20 + 448), (3, 5, 3 + 448, 5 + 224), (100, 7, 100 + 224, 7 + 224),
              (5, 5, 5 + 512, 5 + 512), (7, 300, 7 + 512, 300 + 96), (640, 0, 648, 400), (0, 0, 9, 8),
              (11, 13, 11 + 225, 13 + 223), (200, 100, 200 + 223, 100 + 224), (0, 0, 648, 520)]
    blob = {'image': image, 'boxes': np.array(boxes, np.int32)}
    for T in (224, 256, 64):
        canv, geom = [], []
        for (x1, y1, x2, y2) in (boxes if T != 256 else boxes[::2]):
            crop = image[y1:y2, x1:x2]
            letter, scale, dx, dy = ref.du.letterbox_preserving_aspect_ratio(
                crop, target_size=T, fill_color=(255, 255, 255))
            canv.append(letter)
            geom.append((scale, dx, dy))
        blob[f'canvas_T{T}'] = np.stack(canv)
        blob[f'geom_T{T}'] = np.array(geom, np.float64)
    # normalisation table from torchvision's own calls (process_pose.py:206-209)
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)       # "RGB" image
    lut = TF.normalize(TF.to_tensor(ramp), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    blob['lut'] = lut.numpy().reshape(3, 256)
    # a few complete network inputs, the inline crop code of process_pose.py:199-209
    tens = []
    for (x1, y1, x2, y2) in boxes[:4]:
        crop = image[y1:y2, x1:x2]
        letter, _, _, _ = ref.du.letterbox_preserving_aspect_ratio(crop, target_size=64, fill_color=(255, 255, 255))
        rgb = cv2.cvtColor(letter, cv2.COLOR_BGR2RGB)
        t = TF.normalize(TF.to_tensor(rgb), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
        tens.append(t.numpy())
    blob['tensor_T64'] = np.stack(tens)
    blob['versions'] = np.array([f'cv2 {cv2.__version__}', f'torch {torch.__version__}', f'numpy {np.__version__}'])
    np.savez_compressed(os.path.join(GOLDEN, 'crops.npz'), **blob)
    print(f'  crops: {len(boxes)} boxes x T in (224, 64), every other box at T=256')


def golden_dataset(ref):
    """train_crops.npz: BOPSingleObjDataset.__getitem__ (data_utils.py:233-298) on a temp train_pbr directory.

    Recorded per sample: bbox_visib, the three jitter draws (re-drawn from the same seed by the same calls), and --
    through a pass-through recorder around the reference's letterbox_preserving_aspect_ratio -- the shape of the
    window the dataset actually cropped and the uint8 canvas it got back, for the original and the augmented crop;
    plus orig_img_t (float32) for a few samples.  The augmented tensor itself goes through ColorJitter (torch RNG,
    PIL arithmetic) and is not recorded.
    """
    import random
    import cv2
    from bpc_baseline_b200 import synth
    H, W, T, obj = 540, 720, 64, 8
    image = synth.make_images(1, seed=synth.SEED + 11, width=W, height=H)[0]
    rng = np.random.default_rng([synth.SEED, 11])
    boxes = []
    for _ in range(10):
        w, h = (int(v) for v in rng.integers(40, 300, 2))
        boxes.append([int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1)), w, h])
    boxes += [[W - 120, H - 90, 120, 90], [0, 0, 200, 150], [W - 61, 10, 60, 200], [3, H - 75, 333, 70],   # clamps fire
              [100, 100, 250, 250], [0, 200, 45, 37]]
    with tempfile.TemporaryDirectory() as root:
        scene = os.path.join(root, 'train_pbr', '000000')
        os.makedirs(os.path.join(scene, 'rgb_cam1'))
        cv2.imwrite(os.path.join(scene, 'rgb_cam1', '000000.png'), image)
        ident = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
        info = {'0': [{'bbox_visib': b, 'visib_fract': 1.0, 'px_count_all': 5000, 'px_count_valid': 5000} for b in boxes]}
        gt = {'0': [{'obj_id': obj, 'cam_R_m2c': ident, 'cam_t_m2c': [0.0, 0.0, 1000.0]} for _ in boxes]}
        cam = {'0': {'cam_K': [1000.0, 0.0, 360.0, 0.0, 1000.0, 270.0, 0.0, 0.0, 1.0], 'depth_scale': 1.0}}
        for name, blob in (('scene_gt_info_cam1', info), ('scene_gt_cam1', gt), ('scene_camera_cam1', cam)):
            with open(os.path.join(scene, name + '.json'), 'w') as fh:
                json.dump(blob, fh)
        ds = quiet(ref.du.BOPSingleObjDataset, root, ['000000'], ['cam1'], obj, target_size=T, augment=True,
                   split='train', train_ratio=1.0)
        assert len(ds) == len(boxes)
        calls = []
        real = ref.du.letterbox_preserving_aspect_ratio

        def recorder(img, target_size=256, fill_color=(255, 255, 255)):
            out = real(img, target_size=target_size, fill_color=fill_color)
            calls.append((img.shape[:2], out[0].copy()))
            return out

        ref.du.letterbox_preserving_aspect_ratio = recorder
        try:
            rec = {k: [] for k in ('bbox', 'scale', 'shift', 'orig_hw', 'aug_hw', 'orig_canvas', 'aug_canvas', 'orig_t')}
            for i in range(len(ds)):
                x, y, w, h = ds.samples[i]['bbox_visib']
                random.seed(1000 + i)
                scale = 1.0 + 0.2 * random.random()                # the same three calls as data_utils.py:257-263
                shift = (random.randint(-int(0.1 * w), int(0.1 * w)), random.randint(-int(0.1 * h), int(0.1 * h)))
                random.seed(1000 + i)
                del calls[:]
                orig_t, _aug_t, _labels, _meta = ds[i]
                assert len(calls) == 2
                rec['bbox'].append([x, y, w, h]); rec['scale'].append(scale); rec['shift'].append(shift)
                rec['orig_hw'].append(calls[0][0]); rec['aug_hw'].append(calls[1][0])
                rec['orig_canvas'].append(calls[0][1]); rec['aug_canvas'].append(calls[1][1])
                if i < 4:
                    rec['orig_t'].append(orig_t.numpy())
        finally:
            ref.du.letterbox_preserving_aspect_ratio = real
    np.savez_compressed(os.path.join(GOLDEN, 'train_crops.npz'), image=image, T=np.int32(T),
                        bbox=np.array(rec['bbox'], np.int32), scale=np.array(rec['scale'], np.float64),
                        shift=np.array(rec['shift'], np.int32), orig_hw=np.array(rec['orig_hw'], np.int32),
                        aug_hw=np.array(rec['aug_hw'], np.int32), orig_canvas=np.stack(rec['orig_canvas']),
                        aug_canvas=np.stack(rec['aug_canvas']), orig_t=np.stack(rec['orig_t']))
    print(f'  dataset: {len(boxes)} samples at T={T} (order after the constructor shuffle)')


def golden_rotation(ref):
    """rotation.npz: the decode + pose-assembly tail of PoseEstimator._estimate_rotation (process_pose.py:211-239)
    for the three output heads, computed with the reference's own decoders (bpc/pose/models/losses.py:26-84).
    The network itself is replaced by recorded raw outputs (the reference hard-codes .to('cuda'), :210)."""
    import torch
    import bpc.pose.models.losses as L
    from bpc_baseline_b200 import synth
    rng = np.random.default_rng([synth.SEED, 77])
    n = 48
    blob = {}
    # world -> camera rotations with float32-rounded entries widened to float64, as Capture holds them
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q.T
    Rc = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
                   np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
                   np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)
    Rc = Rc.astype(np.float32).astype(np.float64)
    t = rng.normal(size=(n, 3)) * 500.0
    blob['cam_R'] = Rc
    blob['t'] = t
    raws = {'euler': (rng.normal(size=(n, 3)) * 4.0).astype(np.float32),           # beyond +-pi: the wrap of :215 matters
            'quat': (rng.normal(size=(n, 4)) * rng.uniform(0.01, 10.0, (n, 1))).astype(np.float32),
            '6d': rng.normal(size=(n, 6)).astype(np.float32)}
    for mode, raw in raws.items():
        finals, poses = [], []
        for r in range(n):
            raw_pred = raw[r]
            if mode == 'euler':                                    # :213-217
                wrapped_pred = ((raw_pred + np.pi) % (2 * np.pi)) - np.pi
                euler_tensor = torch.tensor(wrapped_pred, dtype=torch.float32).unsqueeze(0)
                rot_mat = L.rotmat_from_euler(euler_tensor).squeeze(0).numpy()
            elif mode == 'quat':                                   # :218-223
                quat_tensor = torch.tensor(raw_pred, dtype=torch.float32).unsqueeze(0)
                quat_tensor = quat_tensor / (quat_tensor.norm(dim=1, keepdim=True) + 1e-8)
                rot_mat = L.quat_to_rotmat(quat_tensor).squeeze(0).numpy()
            else:                                                  # :224-227
                rep6d_tensor = torch.tensor(raw_pred, dtype=torch.float32).unsqueeze(0)
                rot_mat = L.rotmat_from_6d(rep6d_tensor).squeeze(0).numpy()
            final_rot = Rc[r].T @ rot_mat                          # :233
            finals.append(final_rot)
            poses.append(ref.du.calc_pose_matrix(final_rot, t[r]))  # :239
        blob[f'raw_{mode}'] = raw
        blob[f'final_{mode}'] = np.stack(finals)
        blob[f'pose_{mode}'] = np.stack(poses)
    np.savez_compressed(os.path.join(GOLDEN, 'rotation.npz'), **blob)
    print(f'  rotation: {n} raw outputs per head (euler, quat, 6d)')


def golden_detect(ref):
    """detect.npz: PoseEstimator._detect (process_pose.py:113-142) called UNBOUND with a fake `yolo` callable that returns
    seeded boxes / confidences / classes -- no weights needed.  Covers negative coordinates (int() truncates toward
    zero), confidences exactly at the threshold, wrong classes, and cameras without any raw detection."""
    import torch
    rng = np.random.default_rng([20250131, 5])
    S, N, thresh = 6, 40, 0.1
    xyxy = (rng.random((S, 3, N, 4)) * np.array([3840, 2160, 3840, 2160])).astype(np.float32)
    xyxy[..., :2] -= (rng.random((S, 3, N, 2)) * 40).astype(np.float32)
    xyxy[0, 0, 0] = [-0.5, -1.5, 10.9, 20.1]                      # int(): 0, -1, 10, 20
    conf = rng.random((S, 3, N)).astype(np.float32)
    conf[0, 0, :6] = np.float32(thresh)                             # exactly at the threshold: kept (>=)
    conf[0, 1, :6] = np.nextafter(np.float32(thresh), np.float32(0))   # one ulp below: dropped
    cls = rng.integers(0, 3, (S, 3, N)).astype(np.float32)
    nraw = rng.integers(1, N + 1, (S, 3)).astype(np.int32)
    nraw[1, 2] = 0                                                  # len(results.boxes) == 0 -> []
    nraw[2, :] = 0
    cls[3, 0] = 1.0                                                 # detections, none of class 0

    class _Boxes:
        def __init__(self, b, c, k):
            self.xyxy, self.conf, self.cls = torch.from_numpy(b), torch.from_numpy(c), torch.from_numpy(k)

        def __len__(self):
            return int(self.xyxy.shape[0])

    blob = {'xyxy': xyxy, 'conf': conf, 'cls': cls, 'nraw': nraw, 'thresh': np.float64(thresh)}
    kept_max = 0
    for s in range(S):
        calls = iter(range(3))

        def yolo(image, imgsz=1280, _s=s, _calls=calls):
            c = next(_calls)
            n = nraw[_s, c]
            return [SimpleNamespace(boxes=_Boxes(xyxy[_s, c, :n].copy(), conf[_s, c, :n].copy(), cls[_s, c, :n].copy()))]

        me = SimpleNamespace(yolo=yolo, params=ref.pp.PoseEstimatorParams(yolo_conf_thresh=thresh))
        cap = SimpleNamespace(images=[np.zeros((4, 4, 3), np.uint8)] * 3)
        out = quiet(ref.pp.PoseEstimator._detect, me, cap)
        for c in range(3):
            dets = out[c]
            kept_max = max(kept_max, len(dets))
            blob[f'bbox_{s}_{c}'] = np.array([d['bbox'] for d in dets], np.int64).reshape(len(dets), 4)
            blob[f'center_{s}_{c}'] = np.array([d['bb_center'] for d in dets], np.float64).reshape(len(dets), 2)
    np.savez_compressed(os.path.join(GOLDEN, 'detect.npz'), **blob)
    print(f'  detect: {S} scenes x 3 cameras x <= {N} raw detections, at most {kept_max} kept')


def golden_skew(ref):
    """skew.npz: scenes whose intrinsics have skew and unequal focal lengths, so np.linalg.inv(K) (float32 LAPACK
    sgetrf / sgetri) is not the closed pinhole form: compute_fundamental_matrix, compute_cost_matrix, match list, X."""
    from bpc_baseline_b200 import synth
    batch = synth.make_scenes(6, 8, seed=synth.SEED + 21)
    rng = np.random.default_rng([synth.SEED, 21])
    blob, names = {}, []
    for s in range(6):
        Ks = batch.Ks[s].copy()
        for c in range(3):
            Ks[c, 0, 1] = np.float32(rng.uniform(-8.0, 8.0))          # skew
            Ks[c, 1, 1] = np.float32(Ks[c, 1, 1] * rng.uniform(0.97, 1.03))
            if s >= 4:                                               # a general (non upper-triangular) matrix as well
                Ks[c, 1, 0] = np.float32(rng.uniform(-2.0, 2.0))
                Ks[c, 2, 0] = np.float32(rng.uniform(-1e-5, 1e-5))
        RTs = batch.RTs[s]
        res = run_match(ref, [Ks[c] for c in range(3)], [RTs[c] for c in range(3)], batch.detections(s))
        tag = f'skew_{s}'
        names.append(tag)
        blob[f'{tag}/Ks'] = Ks
        blob[f'{tag}/RTs'] = RTs
        blob[f'{tag}/boxes'] = batch.boxes[s]
        blob[f'{tag}/centers'] = batch.centers[s]
        blob[f'{tag}/counts'] = batch.counts[s]
        blob[f'{tag}/ref_Kinv'] = np.stack([np.linalg.inv(Ks[c]) for c in range(3)])
        for k, v in res.items():
            blob[f'{tag}/ref_{k}'] = v
        print(f'  {tag}: matches={len(res["idx"])}')
    blob['names'] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, 'skew.npz'), **blob)


def main():
    if not os.path.isdir(REFERENCE):
        raise SystemExit('needs /root/reference (authoring container only)')
    sys.path.insert(0, ROOT)
    os.makedirs(GOLDEN, exist_ok=True)
    ref = import_reference()
    only = set(sys.argv[1:])
    for name, fn in (('geometry', golden_geometry), ('bop_scene', golden_bop_scene), ('crops', golden_crops),
                     ('dataset', golden_dataset), ('rotation', golden_rotation), ('detect', golden_detect),
                     ('skew', golden_skew)):
        if not only or name in only:
            print(name)
            fn(ref)


if __name__ == '__main__':
    main()
