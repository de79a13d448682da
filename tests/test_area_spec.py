"""oracle.area_spec (INTER_AREA restatement) pinned bit-for-bit against the installed cv2."""
import cv2
import numpy as np
import pytest

from oracle.area_spec import regime, resize_area_u8
from oracle.crop import letterbox_geometry


def _img(rng, h, w, smooth):
    if not smooth:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    g = (127.5 + 60 * np.sin(xx * 0.07) + 50 * np.cos(yy * 0.05))[..., None] + rng.normal(0, 5, (h, w, 3))
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)


def _cases():
    rng = np.random.default_rng(11)
    out = []
    for T in (224, 256):
        for _ in range(25):
            h, w = rng.integers(8, 420, 2)
            out.append((int(h), int(w), T))
        out += [(T, T, T), (2 * T, 2 * T, T), (2 * T, T, T), (3 * T, 2 * T, T), (T, T // 2, T), (T + 1, T, T),
                (T - 1, T, T), (900, 300, T), (8, 8, T), (8, 400, T), (400, 9, T), (3 * T, 3 * T, T),
                (100, T, T), (T, 100, T)]
    return out


@pytest.mark.parametrize('h,w,T', _cases())
def test_against_cv2(h, w, T):
    rng = np.random.default_rng(h * 1000 + w)
    src = _img(rng, h, w, smooth=(h + w) % 2 == 0)
    _, nw, nh, _, _ = letterbox_geometry(h, w, T)
    ref = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_AREA)
    got = resize_area_u8(src, nw, nh)
    assert np.array_equal(ref, got), (regime(w, h, nw, nh), int(np.abs(ref.astype(int) - got).max()))


def test_all_regimes_covered():
    seen = {regime(w, h, *letterbox_geometry(h, w, T)[1:3]) for (h, w, T) in _cases()}
    assert seen == {1, 2, 3}
