#!/usr/bin/env python
"""Generate tests/golden/simcam_reference.npz by running the REFERENCE's own
``sim_cam_motion_bg_extract`` (cil_tools/extract_background.py:78-99), unmodified.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden_simcam.py

Per case the fixture holds the input frames (written as PNG folders for the reference, so the decode is
lossless), the torch seed, the arguments, the frames after the reference's RandomResizedCrop (captured by
re-running the same torchvision transform under the same seed; they pin the kernel's input independently of
the torchvision version on the test box), the uint8 image the reference passed to cv2.imwrite and the JPEG
bytes it wrote.
"""
from __future__ import annotations

import pathlib
import sys
import tempfile
import warnings

import cv2
import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle import _ref_import, simcam_oracle as so  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def make_frames(T, H, W, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 256, (H, W, 3), dtype=np.uint8)
    fr = np.clip(base.astype(np.int16) + rng.integers(-30, 31, (T, H, W, 3)), 1, 255).astype(np.uint8)
    fr[:, : H // 3] = 0                       # a black band: NaN wherever a crop shows it, all-NaN pixels possible
    fr[:, :, -W // 6:, 1] = 0                 # one channel missing on the right
    fr[::3, H // 2:, : W // 4] = 0            # missing in every third frame only
    return fr


def main():
    ref = _ref_import.load_extract_background()
    out = {}
    cases = [("median_i1", 8, 60, 80, 1, 500, 0, 11), ("mean_i1", 8, 60, 80, 1, 500, 1, 11),
             ("median_i2_cap3", 11, 48, 64, 2, 3, 0, 12), ("mean_single", 2, 40, 40, 1, 500, 1, 13),
             ("median_even", 7, 64, 48, 1, 6, 0, 14)]
    for name, T, H, W, interval, max_frames, avg, seed in cases:
        frames = make_frames(T, H, W, seed)
        with tempfile.TemporaryDirectory() as tmp:
            d = pathlib.Path(tmp) / "video"
            d.mkdir()
            for i, f in enumerate(frames):
                assert cv2.imwrite(str(d / f"img_{i + 1:05d}.png"), cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
            captured = {}
            real = cv2.imwrite
            ref.cv2.imwrite = lambda p, img, *a: (captured.__setitem__("img", img.copy()), real(p, img, *a))[1]
            try:
                torch.manual_seed(seed)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ret = ref.sim_cam_motion_bg_extract(d, pathlib.Path(tmp) / "o.jpg", False, interval, max_frames, avg)
            finally:
                ref.cv2.imwrite = real
            assert ret is None
            jpeg = np.frombuffer((pathlib.Path(tmp) / "o.jpg").read_bytes(), np.uint8)
        written = captured["img"]                                   # = cvtColor(ave_frame, BGR2RGB)
        ave = cv2.cvtColor(written, cv2.COLOR_BGR2RGB)              # the swap is an involution
        torch.manual_seed(seed)
        tf = so.transform_frames(frames[so.select_file_indices(T, interval, max_frames)])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(so.cast_u8(so.nan_temporal_reduce(tf, avg)), ave), name
        out[f"{name}/frames"] = frames
        out[f"{name}/params"] = np.array([interval, max_frames, avg, seed], np.int64)
        out[f"{name}/transformed"] = tf
        out[f"{name}/expected"] = ave
        out[f"{name}/jpeg"] = jpeg
        print(name, frames.shape, tf.shape, "nan frac", float(np.isnan(tf).mean()), "all-nan px", int(np.isnan(tf).all(0).sum()))
    np.savez_compressed(GOLDEN / "simcam_reference.npz", **out)
    print("wrote", GOLDEN / "simcam_reference.npz")


if __name__ == "__main__":
    main()
