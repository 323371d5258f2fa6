"""CPU oracle for the simulated-camera-motion ("type C") background extraction.

TEST INFRASTRUCTURE ONLY (see median_oracle.py for the rules): imported by ``tests/`` and by
``__graft_entry__.smoke()`` as the checker, never by ``bgdebias_b200``.

Restates ``sim_cam_motion_bg_extract`` of the reference (cil_tools/extract_background.py:78-99):

    image_files = sorted(data_path.glob('*'))                                  :82-83
    for i, f in enumerate(image_files[:-1:interval]):  stop at i == max_frames  :86-88
        frame = read_image(f).float()                                          :89
        frame = RandomResizedCrop(size=100)(frame).permute(1, 2, 0).numpy()    :81,90
        frame[frame == 0] = nan                                                :91
    ave = nanmedian / nanmean over the frames, .astype(uint8)                  :94-98
    cv2.imwrite(dest, cv2.cvtColor(ave, cv2.COLOR_BGR2RGB))                    :99

The crop + resize is third-party arithmetic (torchvision ``RandomResizedCrop``: torch RNG draws and
an antialiased bilinear resize); the oracle executes the installed torchvision for it, exactly as the
reference does, and restates the part the CUDA kernel replaces -- the NaN-masked temporal reduction:

    median: valid values sorted, n of them:  f32((s[(n-1)//2] + s[n//2]) / 2)   (np.nanmedian on float32:
            the two middles are summed in float32, then divided by 2)
    mean:   float32 running sum over the valid values in frame order, then f32(f64(sum) / f64(n))
            (np.nansum along a strided axis adds frame by frame; _divide_by_count divides by an int64 count)
    n == 0: NaN; the uint8 cast of NaN is 0 on x86-64 (cvttss2si), values 0 <= v < 256 truncate.

Parity status: PINNED.  ``oracle/gen_golden_simcam.py`` runs the reference function itself on seeded PNG
folders and commits the inputs, the transformed frames and the reference's output image under
``tests/golden/simcam_reference.npz``.
"""
from __future__ import annotations

from typing import List

import numpy as np

CROP = 100


def select_file_indices(n_files: int, interval: int, max_frames: int) -> List[int]:
    """``image_files[:-1:interval]`` cut at ``max_frames`` items (extract_background.py:86-88)."""
    return list(range(0, max(n_files - 1, 0), interval))[:max_frames]


def transform_frames(frames_rgb_u8: np.ndarray) -> np.ndarray:
    """``[T, H, W, 3]`` uint8 RGB -> ``[T, 100, 100, 3]`` float32 with zeros turned into NaN: the body of
    the reference's loop, one ``RandomResizedCrop(size=100)`` call per frame in order (torch global RNG)."""
    import torch
    from torchvision.transforms import Compose, RandomResizedCrop
    pipe = Compose([RandomResizedCrop(size=CROP)])
    out = []
    for f in frames_rgb_u8:
        x = torch.from_numpy(np.ascontiguousarray(f)).permute(2, 0, 1).float()
        y = pipe(x).permute(1, 2, 0).numpy()
        y[y == 0] = np.nan
        out.append(y)
    return np.stack(out).astype(np.float32)


def cast_u8(x: np.ndarray) -> np.ndarray:
    """``ndarray.astype(uint8)`` of float32 as the reference's platform does it: truncation for
    0 <= v < 256, 0 for NaN."""
    y = np.where(np.isnan(x), np.float32(0), x)
    return np.trunc(y).astype(np.int64).astype(np.uint8)


def nan_temporal_reduce(frames: np.ndarray, avg_method: int) -> np.ndarray:
    """``frames``: float32 ``[T, ...]`` with NaN for missing; returns the float32 reduction over axis 0
    (NaN where no frame is valid).  avg_method 0 = median, 1 = mean (extract_background.py:94-98)."""
    fr = np.asarray(frames, dtype=np.float32)
    T = fr.shape[0]
    flat = fr.reshape(T, -1)
    valid = ~np.isnan(flat)
    n = valid.sum(0)
    if avg_method == 0:
        s = np.sort(flat, axis=0)                         # NaN sorts last
        cols = np.arange(flat.shape[1])
        lo = s[np.clip((n - 1) // 2, 0, T - 1), cols]
        hi = s[np.clip(n // 2, 0, T - 1), cols]
        res = ((lo + hi).astype(np.float32) / np.float32(2)).astype(np.float32)
    else:
        acc = np.zeros(flat.shape[1], np.float32)
        for t in range(T):
            acc = (acc + np.where(valid[t], flat[t], np.float32(0))).astype(np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            res = (acc.astype(np.float64) / n.astype(np.float64)).astype(np.float32)
    res = np.where(n == 0, np.float32(np.nan), res).astype(np.float32)
    return res.reshape(fr.shape[1:])


def sim_cam_background(frames_rgb_u8: np.ndarray, interval: int, max_frames: int, avg_method: int) -> np.ndarray:
    """The uint8 ``[100, 100, 3]`` image the reference hands to ``cv2.cvtColor`` / ``cv2.imwrite``
    for a folder whose sorted images decode to ``frames_rgb_u8``.  Consumes the torch global RNG."""
    idx = select_file_indices(len(frames_rgb_u8), interval, max_frames)
    tf = transform_frames(frames_rgb_u8[idx])
    return cast_u8(nan_temporal_reduce(tf, avg_method))
