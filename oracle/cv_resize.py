"""CPU restatement of OpenCV's bilinear resize for uint8 images.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference's pre-processing is ``Compose([Resize((224, 224)), Normalize(mean, std), ToTensor()])`` on the HWC uint8
array ``load_image`` returns (demo/image_classification/predict-resnet.py:49-56).  ``Resize`` lives in the third-party
``tensorlayerx`` package (requirements/requirements.txt:1, ``tensorlayerx>=0.5.8``, not vendored), whose numpy-image path
calls ``cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)`` [recalled from tensorlayerx/vision/functional_cv2.py];
the arithmetic is therefore OpenCV's (``opencv-python`` 4.13.0 in this image): ``modules/imgproc/src/resize.cpp``,
``resize()`` coefficient set-up + ``HResizeLinear`` / ``VResizeLinear<uchar, int, short, ...>``:

  * per destination column: ``fx = float((dx + 0.5) * scale_x - 0.5)``, ``sx = floor(fx)``, ``fx -= sx``; columns falling
    left of the image take ``(sx, fx) = (0, 0)``, columns at or beyond the last source column ``(W_src - 1, 0)``;
    coefficients are 11-bit fixed point: ``a0 = round_half_even((1 - fx) * 2048)``, ``a1 = round_half_even(fx * 2048)``;
  * rows: the same with ``scale_y`` but WITHOUT the fraction clamp - the two source rows ``sy, sy + 1`` are clamped into
    the image instead;
  * horizontal pass (32-bit): ``D = S[sx] * a0 + S[sx + 1] * a1``;
  * vertical pass: ``dst = ((b0 * (D0 >> 4) >> 16) + (b1 * (D1 >> 4) >> 16) + 2) >> 2``.

Pinned bit-for-bit against ``cv2.resize`` itself wherever ``cv2`` imports (tests/test_oracle.py) and against the committed
vectors ``tests/golden/cv_resize.npz`` (minted from cv2 by tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def coefficients(dst: int, src: int, clamp_fraction: bool):
    """(i0, i1, c0, c1): source indices and 11-bit weights per destination index."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo, hi = s < 0, s >= src - 1
        f[lo], s[lo] = 0, 0
        f[hi], s[hi] = 0, src - 1
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int64)
    c1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), c0, c1


def resize_u8(img: np.ndarray, height: int, width: int) -> np.ndarray:
    """``cv2.resize(img, (width, height), interpolation=cv2.INTER_LINEAR)`` for an (H, W, C) uint8 array."""
    if img.dtype != np.uint8 or img.ndim != 3:
        raise TypeError("resize_u8 expects an (H, W, C) uint8 image")
    hs, ws = img.shape[:2]
    if (hs, ws) == (height, width):
        return img.copy()
    x0, x1, a0, a1 = coefficients(width, ws, True)
    y0, y1, b0, b1 = coefficients(height, hs, False)
    src = img.astype(np.int64)
    top = src[y0][:, x0] * a0[None, :, None] + src[y0][:, x1] * a1[None, :, None]
    bot = src[y1][:, x0] * a0[None, :, None] + src[y1][:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (top >> 4)) >> 16) + ((b1[:, None, None] * (bot >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
