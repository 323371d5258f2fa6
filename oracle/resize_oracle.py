"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the foreground `Resize` of the reference's training pipeline.

The reference's config ends its geometric pipeline with ``dict(type='Resize', scale=(224, 224), keep_ratio=False)``
(configs/ucf101/bgmix_plus_randAug/bgmix_seed_1000_inc_10_stages_bgmix_plus_randAug.py:136).  That transform is third-party
code absent from /root/reference: mmaction2 0.x ``Resize`` -> mmcv 1.x ``imresize`` ->
``cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)`` (no version pinned by the reference; cv2 4.13.0 here).
This file restates OpenCV's published 8-bit linear-resize algorithm (modules/imgproc/src/resize.cpp: ``resizeGeneric_``
with ``HResizeLinear<uchar,int,short,2048>`` / ``VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>``) in numpy.

Pinned: tests/test_oracle_resize.py checks it bit for bit against cv2.resize itself (importable on both boxes) over a few
hundred shapes and against tests/golden/resize_reference.npz (outputs of cv2.resize recorded by oracle/gen_golden_resize.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def linear_coeffs(src: int, dst: int, clamp: bool):
    """Tap index and the two 11-bit weights of every destination sample.

    cv::resize: ``inv_scale = dst / src`` (double), ``scale = 1 / inv_scale``; per sample
    ``f = float((d + 0.5) * scale - 0.5)``, ``s = floor(f)``, ``f -= s``.  Columns (``clamp``) snap the taps that fall
    outside the image onto the border pixel with weight (1, 0); rows keep their fraction and clamp the two tap rows instead.
    Weights are ``saturate_cast<short>(c * 2048)`` = round-half-even of an exact float product."""
    scale = 1.0 / (float(dst) / float(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        s[lo], f[lo] = 0, 0.0
        hi = s >= src - 1
        s[hi], f[hi] = src - 1, 0.0
    c0 = np.clip(np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int64)
    c1 = np.clip(np.rint(f * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int64)
    return s, c0, c1


def resize_linear_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """``cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR)`` for uint8 ``[h, w]`` / ``[h, w, c]``."""
    img = np.asarray(img)
    assert img.dtype == np.uint8
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    h, w, _ = img.shape
    sx, a0, a1 = linear_coeffs(w, dst_w, True)
    sy, b0, b1 = linear_coeffs(h, dst_h, False)
    src = img.astype(np.int64)
    x1 = np.minimum(sx + 1, w - 1)
    rows = src[:, sx, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]        # horizontal pass, x2048
    r0 = rows[np.clip(sy, 0, h - 1)] >> 4
    r1 = rows[np.clip(sy + 1, 0, h - 1)] >> 4
    out = (((b0[:, None, None] * r0) >> 16) + ((b1[:, None, None] * r1) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_clip(clip_thwc: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """Every frame of a uint8 ``[T, h, w, 3]`` clip, as mmaction's Resize does (one imresize per frame)."""
    return np.stack([resize_linear_u8(f, dst_h, dst_w) for f in clip_thwc])


def multiscale_crop_sizes(base: int = 256, input_size: int = 224, scales=(1, 0.875, 0.75, 0.66), max_wh_scale_gap: int = 1):
    """(crop_w, crop_h) pairs mmaction's MultiScaleCrop can produce (config :129-135): the shapes Resize sees."""
    sizes = [int(base * s) for s in scales]
    crop_h = [input_size if abs(s - input_size) < 3 else s for s in sizes]
    crop_w = list(crop_h)
    return [(crop_w[j], crop_h[i]) for i in range(len(sizes)) for j in range(len(sizes)) if abs(i - j) <= max_wh_scale_gap]
