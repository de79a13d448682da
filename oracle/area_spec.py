"""Oracle (test infrastructure): restatement of ``cv2.resize(src, (dw, dh), INTER_AREA)`` for 8UC3.

The reference resizes every crop with OpenCV at bpc/utils/data_utils.py:39 (pinned
opencv-python==4.8.0.74, docker/requirements.txt:4; this image has opencv-python-headless
4.13.0).  OpenCV's source is not under /root/reference, so its published algorithm
(modules/imgproc/src/resize.cpp: ``computeResizeAreaTab``, ``ResizeArea_Invoker``,
``ResizeAreaFast_Invoker``, and the ``area_mode`` branch of the generic linear resize) is
restated here; ``tests/test_area_spec.py`` pins it bit-for-bit against the installed cv2.

Three regimes:
  1. both axes shrink, non-integer ratio : true area, float32 taps, sequential float32
     accumulation (multiply and add rounded separately), cvRound (half-to-even) at the end;
  2. both axes shrink by exact integers  : integer box sum; 2x2 -> (sum+2)>>2, otherwise
     cvRound(float(sum) * (1.f/area));
  3. either axis grows                   : 11-bit fixed-point bilinear with area-mode source
     coordinates.
"""
from __future__ import annotations

import math

import numpy as np

_EPS = np.finfo(np.float64).eps


def regime(sw, sh, dw, dh):
    """1 = area, 2 = integer-ratio fast area, 3 = fixed-point bilinear (up-scaling)."""
    scale_x = 1.0 / (dw / float(sw))
    scale_y = 1.0 / (dh / float(sh))
    if scale_x >= 1 and scale_y >= 1:
        ix, iy = int(np.rint(scale_x)), int(np.rint(scale_y))
        if abs(scale_x - ix) < _EPS and abs(scale_y - iy) < _EPS:
            return 2
        return 1
    return 3


def area_tab(ssize, dsize, scale):
    """computeResizeAreaTab: list of (di, si, alpha f32) in emission order."""
    tab = []
    for d in range(dsize):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((d, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((d, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((d, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _padded_tab(tab, dsize):
    per = [[] for _ in range(dsize)]
    for d, s, a in tab:
        per[d].append((s, a))
    nt = max(len(p) for p in per)
    si = np.zeros((nt, dsize), np.int64)
    al = np.zeros((nt, dsize), np.float32)
    ok = np.zeros((nt, dsize), bool)
    for d, p in enumerate(per):
        for t, (s, a) in enumerate(p):
            si[t, d], al[t, d], ok[t, d] = s, a, True
    return si, al, ok


def _resize_area(src, dw, dh, scale_x, scale_y):
    sh, sw, cn = src.shape
    xs, xa, xok = _padded_tab(area_tab(sw, dw, scale_x), dw)
    ytab = area_tab(sh, dh, scale_y)
    # horizontal pass of every source row that some y tap uses
    S = src.astype(np.float32)
    hbuf = {}
    for _, sy, _ in ytab:
        if sy in hbuf:
            continue
        buf = np.zeros((dw, cn), np.float32)
        row = S[sy]
        for t in range(xs.shape[0]):
            term = row[xs[t]] * xa[t][:, None]                  # f32 multiply, rounded
            buf = np.where(xok[t][:, None], buf + term, buf)    # f32 add, rounded
        hbuf[sy] = buf
    out = np.zeros((dh, dw, cn), np.uint8)
    acc = {}
    for dy, sy, beta in ytab:
        term = beta * hbuf[sy]
        acc[dy] = term if dy not in acc else acc[dy] + term
    for dy in range(dh):
        out[dy] = np.clip(np.rint(acc[dy]), 0, 255).astype(np.uint8)
    return out


def _resize_area_fast(src, dw, dh, isx, isy):
    sh, sw, cn = src.shape
    blk = src[:dh * isy, :dw * isx].astype(np.int32).reshape(dh, isy, dw, isx, cn)
    s = blk.sum(axis=(1, 3))
    if isx == 2 and isy == 2:
        return ((s + 2) >> 2).astype(np.uint8)
    scale = np.float32(1.0) / np.float32(isx * isy)
    return np.clip(np.rint(s.astype(np.float32) * scale), 0, 255).astype(np.uint8)


def _linear_coeffs(ssize, dsize, scale, inv_scale):
    """area_mode coordinates of the generic linear resize: (ofs, w0, w1, dmax)."""
    ofs = np.zeros(dsize, np.int64)
    w0 = np.zeros(dsize, np.int32)
    w1 = np.zeros(dsize, np.int32)
    dmax = dsize
    for d in range(dsize):
        s = math.floor(d * scale)
        f = np.float32((d + 1) - (s + 1) * inv_scale)
        f = np.float32(0) if f <= 0 else np.float32(f - np.float32(math.floor(f)))
        if s < 0:
            f, s = np.float32(0), 0
        if s + 1 >= ssize:
            dmax = min(dmax, d)
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        ofs[d] = s
        w0[d] = int(np.rint((np.float32(1) - f) * np.float32(2048)))
        w1[d] = int(np.rint(f * np.float32(2048)))
    return ofs, w0, w1, dmax


def _resize_linear_area_mode(src, dw, dh, scale_x, scale_y, inv_x, inv_y):
    sh, sw, cn = src.shape
    xo, a0, a1, xmax = _linear_coeffs(sw, dw, scale_x, inv_x)
    yo, b0, b1, _ = _linear_coeffs(sh, dh, scale_y, inv_y)
    S = src.astype(np.int32)
    xo1 = np.minimum(xo + 1, sw - 1)
    H = S[:, xo] * a0[None, :, None] + S[:, xo1] * a1[None, :, None]
    if xmax < dw:
        H[:, xmax:] = S[:, xo[xmax:]] * 2048
    r0 = yo
    r1 = np.minimum(yo + 1, sh - 1)
    v = (((b0[:, None, None] * (H[r0] >> 4)) >> 16) + ((b1[:, None, None] * (H[r1] >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def resize_area_u8(src, dw, dh):
    """Model of cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for HxWx3 uint8."""
    src = np.asarray(src)
    sh, sw, _ = src.shape
    if dw < 1 or dh < 1 or sw < 1 or sh < 1:
        raise ValueError("empty source or destination")
    inv_x = dw / float(sw)
    inv_y = dh / float(sh)
    scale_x = 1.0 / inv_x
    scale_y = 1.0 / inv_y
    r = regime(sw, sh, dw, dh)
    if r == 2:
        return _resize_area_fast(src, dw, dh, int(np.rint(scale_x)), int(np.rint(scale_y)))
    if r == 1:
        return _resize_area(src, dw, dh, scale_x, scale_y)
    return _resize_linear_area_mode(src, dw, dh, scale_x, scale_y, inv_x, inv_y)
