"""CPU oracle for the BG-mix augmentation path.

TEST INFRASTRUCTURE ONLY (see oracle/median_oracle.py for the rule): imported by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs, never by the product.

Restates ``BackgroundMixDataset._mix_background`` and the pieces around it:

    libs/loader/comix_loader.py:72-75    bg_pipeline = Resize(256) -> RandomCrop(224) -> Normalize
    libs/loader/comix_loader.py:105-124  prepare_train_frames (the gate and the bg_idx codes)
    libs/loader/comix_loader.py:126-136  _get_bg_image (RNG draw of the pool index)
    libs/loader/comix_loader.py:138-145  _mix_background:  imgs*(1-alpha) + bg*alpha

Third-party arithmetic that the reference reaches but does not contain:

* foreground normalisation = mmaction ``Normalize`` -> ``mmcv.imnormalize_`` (mmcv-full 1.x,
  not installed, version not pinned by the reference).  Its published algorithm is
  ``cv2.subtract(img_f32, float64(mean)); cv2.multiply(img_f32, 1/float64(std))``, which for a
  uint8-valued pixel x evaluates to ``f32( f64( f32(x) - f32(mean) ) * (1/f64(std)) )``
  (checked against cv2 4.13 in :func:`fg_lut_cv2`; tests compare the two).
* background ``Resize`` / ``RandomCrop`` / ``Normalize`` = torchvision (installed, 0.26):
  called directly here, exactly as the reference calls them.

Parity status: PINNED by ``tests/golden/bgmix_*.npz`` -- outputs of the reference class
itself (imported with stub ``mmaction`` modules by ``oracle/gen_golden.py``).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

DEFAULT_MEAN = (123.675, 116.28, 103.53)   # comix_loader.py:29, config img_norm_cfg
DEFAULT_STD = (58.395, 57.12, 57.375)      # comix_loader.py:30


# --------------------------------------------------------------------------- #
# foreground normalisation (mmcv.imnormalize_ semantics) as a 3x256 table
# --------------------------------------------------------------------------- #
def fg_lut(mean: Sequence[float] = DEFAULT_MEAN, std: Sequence[float] = DEFAULT_STD) -> np.ndarray:
    """[3,256] fp32 table: value of the normalised pixel for every uint8 input, per channel."""
    mean32 = np.asarray(mean, dtype=np.float64).astype(np.float32)
    stdinv64 = 1.0 / np.asarray(std, dtype=np.float64)
    x = np.arange(256, dtype=np.float32)
    d32 = (x[None, :] - mean32[:, None]).astype(np.float32)          # cv2.subtract on f32 data
    return (d32.astype(np.float64) * stdinv64[:, None]).astype(np.float32)   # cv2.multiply by f64 scalar


def fg_lut_cv2(mean: Sequence[float] = DEFAULT_MEAN, std: Sequence[float] = DEFAULT_STD) -> np.ndarray:
    """Same table produced by the actual cv2 calls mmcv makes (needs cv2; used to pin fg_lut)."""
    import cv2
    img = np.arange(256, dtype=np.float32).reshape(1, 256, 1).repeat(3, axis=2).copy()
    mean64 = np.float64(np.asarray(mean, dtype=np.float64).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std, dtype=np.float64).reshape(1, -1))
    cv2.subtract(img, mean64, img)
    cv2.multiply(img, stdinv, img)
    return np.ascontiguousarray(img[0].T)


def fg_normalize(fg_u8_thwc: np.ndarray, lut: np.ndarray) -> np.ndarray:
    """uint8 [T,H,W,3] -> fp32 [T,3,H,W]  (Normalize + FormatShape('NCHW'), config :137-138)."""
    fg = np.asarray(fg_u8_thwc)
    out = np.empty((fg.shape[0], 3) + fg.shape[1:3], dtype=np.float32)
    for c in range(3):
        out[:, c] = lut[c][fg[..., c]]
    return out


# --------------------------------------------------------------------------- #
# background pipeline (torchvision ops, called like the reference calls them)
# --------------------------------------------------------------------------- #
def bg_resize(bg_chw, size: int = 256):
    """``Resize(size)`` on a float [3,h,w] tensor (comix_loader.py:72); torchvision defaults."""
    import torch
    from torchvision.transforms import Resize
    t = bg_chw if isinstance(bg_chw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bg_chw))
    t = t.float()
    return Resize(size)(t)


def resized_hw(h: int, w: int, size: int = 256) -> Tuple[int, int]:
    """Output (h, w) of torchvision ``Resize(int)``: short edge -> size, long edge truncated."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def draw_bg_params(n_bg: int, h: int, w: int, crop: Tuple[int, int]) -> Tuple[int, int, int]:
    """RNG draws in the reference's order, on torch's global generator.

    1. ``torch.randint(len(bg_files), (1,))``        comix_loader.py:129
    2. ``RandomCrop.get_params``: nothing if (h, w) == crop, else
       ``i = torch.randint(0, h-th+1)`` (top) then ``j = torch.randint(0, w-tw+1)`` (left).
    ``h, w`` are the sizes AFTER Resize.
    """
    import torch
    th, tw = crop
    bg_idx = int(torch.randint(n_bg, (1,)).item())
    if h < th or w < tw:
        raise ValueError(f"Required crop size {(th, tw)} is larger than input image size {(h, w)}")
    if h == th and w == tw:
        return bg_idx, 0, 0
    top = int(torch.randint(0, h - th + 1, size=(1,)).item())
    left = int(torch.randint(0, w - tw + 1, size=(1,)).item())
    return bg_idx, top, left


def bg_normalize(bg_crop: np.ndarray, mean: Sequence[float] = DEFAULT_MEAN,
                 std: Sequence[float] = DEFAULT_STD) -> np.ndarray:
    """torchvision ``Normalize`` on fp32 [3,H,W]: ``(x - f32(mean)) / f32(std)`` per channel."""
    m = np.asarray(mean, dtype=np.float64).astype(np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float64).astype(np.float32).reshape(3, 1, 1)
    x = np.asarray(bg_crop, dtype=np.float32)
    return ((x - m).astype(np.float32) / s).astype(np.float32)


# --------------------------------------------------------------------------- #
# the blend (comix_loader.py:142)
# --------------------------------------------------------------------------- #
def blend(fg_norm_tchw: np.ndarray, bg_norm_chw: np.ndarray, alpha: float) -> np.ndarray:
    """``imgs * (1 - alpha) + bg.view(1,3,H,W) * alpha`` with torch's scalar semantics.

    ``tensor * python_float`` multiplies by the scalar rounded to fp32; each of the two
    multiplies and the add is a separately rounded fp32 operation.
    """
    w_fg = np.float32(1 - alpha)
    w_bg = np.float32(alpha)
    a = (np.asarray(fg_norm_tchw, np.float32) * w_fg).astype(np.float32)
    b = (np.asarray(bg_norm_chw, np.float32)[None] * w_bg).astype(np.float32)
    return (a + b).astype(np.float32)


def mix_clip(fg_u8_thwc: np.ndarray, bg_resized_chw: np.ndarray, top: int, left: int,
             crop: Tuple[int, int] = (224, 224), alpha: float = 0.5, apply: bool = True,
             mean: Sequence[float] = DEFAULT_MEAN, std: Sequence[float] = DEFAULT_STD,
             bg_mean=None, bg_std=None) -> np.ndarray:
    """One sample end to end from uint8 foreground + resized fp32 background -> fp32 [T,3,H,W]."""
    lut = fg_lut(mean, std)
    fg = fg_normalize(fg_u8_thwc, lut)
    if not apply:
        return fg
    th, tw = crop
    bgc = np.asarray(bg_resized_chw, np.float32)[:, top:top + th, left:left + tw]
    bgn = bg_normalize(bgc, bg_mean if bg_mean is not None else mean, bg_std if bg_std is not None else std)
    return blend(fg, bgn, alpha)


def mix_batch(fg_u8_bthwc: np.ndarray, pool_resized: np.ndarray, bg_idx: Sequence[int],
              top: Sequence[int], left: Sequence[int], apply: Sequence[int],
              crop: Tuple[int, int] = (224, 224), alpha: float = 0.5,
              mean: Sequence[float] = DEFAULT_MEAN, std: Sequence[float] = DEFAULT_STD,
              layout: str = "NTCHW") -> np.ndarray:
    """Batch form the CUDA op implements; default layout is what default_collate gives the
    reference: [B, T, 3, H, W] (libs/cil/cil.py:203-210).  'NCTHW' is the permuted variant."""
    outs = []
    for b in range(len(fg_u8_bthwc)):
        outs.append(mix_clip(fg_u8_bthwc[b], pool_resized[int(bg_idx[b])] if apply[b] else None,
                             int(top[b]), int(left[b]), crop, alpha, bool(apply[b]), mean, std))
    out = np.stack(outs, 0)
    if layout == "NCTHW":
        out = np.ascontiguousarray(out.transpose(0, 2, 1, 3, 4))
    elif layout != "NTCHW":
        raise ValueError(layout)
    return out


def mix_clip_like_reference(imgs, bg_u8_chw, alpha: float = 0.5, bg_resize_size: int = 256, crop=(224, 224),
                            mean: Sequence[float] = DEFAULT_MEAN, std: Sequence[float] = DEFAULT_STD):
    """``_mix_background`` with the libraries the reference calls (comix_loader.py:72-75,139-142): torchvision
    ``Compose([Resize, RandomCrop, Normalize])`` on the float background, then the torch blend.  ``imgs``: fp32
    torch tensor ``[T,3,H,W]`` (already normalised); consumes the torch RNG for the crop.  Used as the CPU arm
    of bench.py's BG-mix leg and to cross-check :func:`mix_clip`."""
    import torch
    from torchvision.transforms import Compose, Normalize, RandomCrop, Resize
    pipe = Compose([Resize(bg_resize_size), RandomCrop(tuple(crop)), Normalize(mean=list(mean), std=list(std))])
    bg = pipe(torch.as_tensor(bg_u8_chw).float())
    bg = bg.view(1, 3, crop[0], crop[1])
    return imgs * (1 - alpha) + bg * alpha


def gate(with_randAug: bool, randAug_flag: bool, prob: float, rand_value: float) -> bool:
    """Whether a sample is mixed (comix_loader.py:111-116); ``rand_value`` = ``random.random()``
    (drawn only when ``with_randAug`` is False)."""
    if with_randAug:
        return not randAug_flag
    return rand_value < prob
