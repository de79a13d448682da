"""Oracle (test infrastructure): restatement of the reference ROI crop -> network-input path.

Follows bpc/utils/data_utils.py:34-44 (letterbox) and bpc/inference/process_pose.py:199-209
(crop slice, letterbox, BGR->RGB, to_tensor, normalize).  ``*_ref`` functions make the same four
library calls the reference makes (cv2.resize, cv2.cvtColor, torchvision to_tensor / normalize);
``*_spec`` functions use the pure-NumPy models instead so a failing kernel can be bisected.
"""
from __future__ import annotations

import numpy as np

from .area_spec import resize_area_u8

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def letterbox_geometry(h, w, target_size):
    """(scale, new_w, new_h, dx, dy) -- data_utils.py:35-38,41-42.  Python round = half-to-even."""
    scale = float(target_size) / max(h, w)
    new_w = int(round(w * scale))
    new_h = int(round(h * scale))
    dx = (target_size - new_w) // 2
    dy = (target_size - new_h) // 2
    return scale, new_w, new_h, dx, dy


def letterbox_ref(img, target_size=256, fill_color=(255, 255, 255)):
    """letterbox_preserving_aspect_ratio with the reference's own cv2 call -- data_utils.py:34-44."""
    import cv2
    h, w = img.shape[:2]
    scale, new_w, new_h, dx, dy = letterbox_geometry(h, w, target_size)
    resized = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_AREA)
    canvas = np.full((target_size, target_size, 3), fill_color, dtype=np.uint8)
    canvas[dy:dy + new_h, dx:dx + new_w] = resized
    return canvas, scale, dx, dy


def letterbox_spec(img, target_size=256, fill_color=(255, 255, 255)):
    """Same, with the NumPy INTER_AREA model instead of cv2."""
    h, w = img.shape[:2]
    scale, new_w, new_h, dx, dy = letterbox_geometry(h, w, target_size)
    resized = resize_area_u8(img, new_w, new_h)
    canvas = np.full((target_size, target_size, 3), fill_color, dtype=np.uint8)
    canvas[dy:dy + new_h, dx:dx + new_w] = resized
    return canvas, scale, dx, dy


def normalise_lut(mean=MEAN, std=STD):
    """f32 [3, 256]: LUT[c][v] == normalize(to_tensor(v))[c] bit-for-bit.

    to_tensor: uint8 -> float32, true division by 255 (torchvision functional.py ``.div(255)``);
    normalize: ``tensor.sub_(mean).div_(std)`` with float32 mean/std tensors.
    """
    import torch
    v = torch.arange(256, dtype=torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32)
    s = torch.as_tensor(std, dtype=torch.float32)
    return ((v[None, :] - m[:, None]) / s[:, None]).numpy()


def crop_tensor_ref(image, box, target_size=256, swap_rb=True):
    """One network input [3, T, T] f32 with the reference's four library calls.

    process_pose.py:199-209; ``swap_rb=False`` gives the training-dataset variant
    (data_utils.py:243-252,282) which skips BGR->RGB.
    """
    import cv2
    import torchvision.transforms.functional as TF
    x1, y1, x2, y2 = [int(v) for v in box]
    crop = image[y1:y2, x1:x2]
    letter_img, _, _, _ = letterbox_ref(crop, target_size=target_size, fill_color=(255, 255, 255))
    if swap_rb:
        letter_img = cv2.cvtColor(letter_img, cv2.COLOR_BGR2RGB)
    tens = TF.to_tensor(letter_img)
    tens = TF.normalize(tens, list(MEAN), list(STD))
    return tens.numpy()


def crop_u8_ref(image, box, target_size=256):
    """The uint8 letterboxed crop [T, T, 3] (BGR order, before colour conversion)."""
    x1, y1, x2, y2 = [int(v) for v in box]
    return letterbox_ref(image[y1:y2, x1:x2], target_size=target_size)[0]


def crop_tensor_spec(image, box, target_size=256, swap_rb=True):
    """Same as crop_tensor_ref but through area_spec + the LUT (no cv2 / torchvision calls)."""
    x1, y1, x2, y2 = [int(v) for v in box]
    canvas = letterbox_spec(image[y1:y2, x1:x2], target_size=target_size)[0]
    if swap_rb:
        canvas = canvas[..., ::-1]
    lut = normalise_lut()
    return np.stack([lut[c][canvas[..., c]] for c in range(3)])


def crops_ref(images, rois, target_size=256, swap_rb=True):
    """Batch form: images u8 [B,H,W,3], rois i32 [R,5]=(img,x1,y1,x2,y2) -> f32 [R,3,T,T]."""
    out = np.empty((len(rois), 3, target_size, target_size), np.float32)
    for r, (b, x1, y1, x2, y2) in enumerate(np.asarray(rois)):
        out[r] = crop_tensor_ref(images[b], (x1, y1, x2, y2), target_size, swap_rb)
    return out


def dataset_window(bbox_visib, W, H, scale=None, shift=(0, 0)):
    """Crop window (x1, y1, x2, y2) of BOPSingleObjDataset.__getitem__ -- data_utils.py:242-246 (original,
    scale=None) and :257-268 (augmented, given the draws scale_factor and (shift_x, shift_y))."""
    x, y, w, h = (int(v) for v in bbox_visib)
    if scale is None:
        return x, y, min(x + w, W), min(y + h, H)              # NumPy slicing clamps the ends
    aug_w = int(round(w * float(scale)))
    aug_h = int(round(h * float(scale)))
    aug_x = max(0, min(x - int(shift[0]), W - 1))
    aug_y = max(0, min(y - int(shift[1]), H - 1))
    aug_w = min(aug_w, W - aug_x)
    aug_h = min(aug_h, H - aug_y)
    return aug_x, aug_y, aug_x + aug_w, aug_y + aug_h


def dataset_tensor_ref(image, window, target_size=256):
    """orig_img_t of the dataset: letterbox, HWC->CHW, /255, normalize -- BGR order kept (data_utils.py:249-252, 282)."""
    return crop_tensor_ref(image, window, target_size=target_size, swap_rb=False)
