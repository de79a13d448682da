"""Oracle (test infrastructure): restatement of the detector post-processing of PoseEstimator._detect,
reference bpc/inference/process_pose.py:123-141 (class / confidence filter, int() truncation, centres).
Parity unpinned by reference outputs: the lines are inline in a method that needs YOLO weights to run, so
this restatement is checked by reading only (it is ten lines)."""
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
