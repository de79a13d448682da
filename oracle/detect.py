"""Oracle (test infrastructure): restatement of the detector post-processing of PoseEstimator._detect,
reference bpc/inference/process_pose.py:123-141 (class / confidence filter, int() truncation, centres).
Pinned by tests/golden/detect.npz: oracle/make_golden.py calls the reference's PoseEstimator._detect unbound
with a fake `yolo` callable returning seeded boxes / confidences / classes (no weights needed), and
tests/test_oracle_golden.py::test_detect_restatement checks this restatement against those outputs."""
from __future__ import annotations

import numpy as np


def detections_from_yolo(boxes_xyxy, confs, clss, conf_thresh):
    """One camera: arrays as ultralytics returns them (float32) -> list of {'bbox', 'bb_center'} dicts."""
    boxes = np.asarray(boxes_xyxy)
    confs = np.asarray(confs)
    clss = np.asarray(clss)
    if len(boxes) == 0:                                   # process_pose.py:126-128
        return []
    valid = (clss == 0) & (confs >= conf_thresh)          # :130
    boxes = boxes[valid]
    preds_cam = []
    for box in boxes:                                     # :133-141
        x1, y1, x2, y2 = map(int, box)
        cx = 0.5 * (x1 + x2)
        cy = 0.5 * (y1 + y2)
        preds_cam.append({'bbox': (x1, y1, x2, y2), 'bb_center': (cx, cy)})
    return preds_cam
