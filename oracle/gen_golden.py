#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own functions (unmodified).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference, cv2 with FFV1):

    python oracle/gen_golden.py            # rewrites tests/golden/

What is executed from the reference:
* ``cil_tools/extract_background.py:42-75``   ``bg_extraction_tmf`` (video mode) on lossless
  FFV1 ``.avi`` files written from seeded synthetic frames (FFV1 round-trips bit-exactly);
* ``libs/loader/comix_loader.py:148-164``     the rawframes ``bg_extraction_tmf`` on PNG folders;
* ``libs/loader/comix_loader.py:16-145``      ``BackgroundMixDataset`` (ctor, ``prepare_train_frames``,
  ``_get_bg_image``, ``_mix_background``) with stub mmaction base classes (oracle/_ref_import.py).
  The JPEG decode inside ``_get_bg_image`` (``torchvision.io.read_image``) is replaced by a
  lookup into seeded synthetic images -- decode is outside the kernel boundary.
* the foreground normalisation the reference gets from mmcv is produced with the cv2 calls
  mmcv makes (``cv2.subtract`` / ``cv2.multiply``), because mmcv itself is not installed.

The fixtures hold inputs AND reference outputs, so tests on the GPU box (no reference tree)
can check the oracle and the CUDA path against them.
"""
from __future__ import annotations

import hashlib
import os
import pathlib
import random
import sys
import tempfile

import cv2
import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle import _ref_import  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


# --------------------------------------------------------------------------- #
def make_frames(pattern: str, T: int, H: int, W: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if pattern == "random":
        return rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    if pattern == "constant":
        return np.full((T, H, W, 3), 137, np.uint8)
    if pattern == "two_valued":
        return rng.choice(np.array([3, 250], np.uint8), size=(T, H, W, 3))
    if pattern == "saturated":
        return rng.choice(np.array([0, 255], np.uint8), size=(T, H, W, 3))
    if pattern == "sorted":
        base = np.linspace(0, 255, T).astype(np.uint8).reshape(T, 1, 1, 1)
        return np.broadcast_to(base, (T, H, W, 3)).copy()
    if pattern == "reverse_sorted":
        base = np.linspace(255, 0, T).astype(np.uint8).reshape(T, 1, 1, 1)
        return np.broadcast_to(base, (T, H, W, 3)).copy()
    if pattern == "moving_block":   # static background + moving bright block (what TMF is for)
        bg = rng.integers(40, 200, (H, W, 3), dtype=np.uint8)
        fr = np.broadcast_to(bg, (T, H, W, 3)).copy()
        for t in range(T):
            x = (t * 3) % max(1, W - 6)
            fr[t, 2:8, x:x + 6] = 255 - (t % 7)
        return fr
    if pattern == "near_ties":      # values within +-2 of a base: many equal keys, narrow spread
        base = rng.integers(2, 254, (1, H, W, 3))
        return (base + rng.integers(-2, 3, (T, H, W, 3))).astype(np.uint8)
    raise ValueError(pattern)


def write_ffv1(path: str, frames: np.ndarray) -> None:
    T, H, W, _ = frames.shape
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 25, (W, H))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()
    cap = cv2.VideoCapture(path)     # make sure the container round-trips bit-exactly
    got = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        got.append(f)
    assert np.array_equal(np.array(got), frames), "FFV1 round trip is not lossless here"


MEDIAN_CASES = [
    # name,            pattern,        T_total, H,  W, interval, max_frames
    ("t1",             "random",          1,   16, 24, 1, 500),
    ("t2",             "random",          2,   16, 24, 1, 500),
    ("t3",             "random",          3,   16, 24, 1, 500),
    ("t4",             "random",          4,   16, 24, 1, 500),
    ("t7",             "random",          7,   16, 24, 1, 500),
    ("t8",             "random",          8,   16, 24, 1, 500),
    ("t64",            "random",         64,   16, 24, 1, 500),
    ("t65",            "random",         65,   16, 24, 1, 500),
    ("t180_odd_w",     "random",        180,   10, 14, 1, 500),   # W*3 = 42: not a multiple of 4/16
    ("constant",       "constant",       17,   16, 24, 1, 500),
    ("two_valued_even","two_valued",     20,   16, 24, 1, 500),
    ("saturated_even", "saturated",      32,   16, 24, 1, 500),
    ("sorted",         "sorted",         33,   16, 24, 1, 500),
    ("reverse_sorted", "reverse_sorted", 34,   16, 24, 1, 500),
    ("moving_block",   "moving_block",   48,   16, 48, 1, 500),
    ("near_ties_even", "near_ties",      96,   16, 24, 1, 500),
    ("interval3",      "random",         50,   16, 24, 3, 500),   # keeps counts 0,3,...,48 -> 17 frames
    ("max_frames4",    "random",         20,   16, 24, 1, 4),     # off-by-one: 5 frames used
    ("interval2_max5", "random",         40,   16, 24, 2, 5),     # 6 frames: counts 0,2,..,10
    ("t520_max500",    "random",        520,    8,  8, 1, 500),   # default cap: 501 frames used
]


def gen_median(ref_extract, ref_comix, tmp: str) -> None:
    data = {}
    for name, pattern, T, H, W, interval, max_frames in MEDIAN_CASES:
        seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
        frames = make_frames(pattern, T, H, W, seed)
        vid = os.path.join(tmp, name + ".avi")
        write_ffv1(vid, frames)
        dest = pathlib.Path(tmp) / (name + ".jpg")
        out = ref_extract.bg_extraction_tmf(pathlib.Path(vid), dest, True, interval, max_frames, 0)
        assert out.dtype == np.uint8 and out.shape == (H, W, 3)
        data[name + "/frames"] = frames
        data[name + "/expected"] = out
        data[name + "/params"] = np.array([interval, max_frames], np.int64)
        # rawframes variant (comix_loader.py:148-164): all files, no interval, no cap
        if T <= 65:
            d = pathlib.Path(tmp) / (name + "_frames")
            d.mkdir()
            for t, f in enumerate(frames):
                cv2.imwrite(str(d / f"img_{t + 1:05}.png"), f)
            out2 = ref_comix.bg_extraction_tmf(d, pathlib.Path(tmp) / (name + "_e2.jpg"))
            data[name + "/expected_rawframes"] = out2
    np.savez_compressed(GOLDEN / "median_reference.npz", **data)
    print("median_reference.npz:", len(MEDIAN_CASES), "cases,",
          (GOLDEN / "median_reference.npz").stat().st_size, "bytes")


# --------------------------------------------------------------------------- #
def mmcv_style_normalize(fg_u8_thwc: np.ndarray, mean, std) -> np.ndarray:
    """mmaction Normalize(to_bgr=False) + FormatShape('NCHW'): the cv2 calls mmcv.imnormalize_ makes."""
    imgs = np.empty(fg_u8_thwc.shape, dtype=np.float32)
    for i, img in enumerate(fg_u8_thwc):
        imgs[i] = img
    mean64 = np.float64(np.asarray(mean, np.float64).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std, np.float64).reshape(1, -1))
    for img in imgs:
        cv2.subtract(img, mean64, img)
        cv2.multiply(img, stdinv, img)
    return np.ascontiguousarray(imgs.transpose(0, 3, 1, 2))


MIX_CASES = [
    # name,        T, crop,      bg_resize, bg (h, w),  n_bg, alpha, with_randAug, prob, seed
    ("small_a05",  4, (32, 32),  40,        (36, 48),   5,    0.5,   True,         0.25, 11),
    ("small_a03",  4, (32, 32),  40,        (36, 48),   5,    0.3,   True,         0.25, 12),
    ("small_a07",  3, (24, 40),  48,        (60, 50),   3,    0.7,   True,         0.25, 13),   # portrait bg
    ("exact_size", 2, (32, 32),  32,        (32, 32),   4,    0.25,  True,         0.25, 14),   # RandomCrop draws nothing
    ("prob_gate",  2, (16, 16),  20,        (30, 40),   6,    0.5,   False,        0.5,  15),   # random.random() gate
    ("down_scale", 2, (32, 32),  36,        (90, 120),  2,    0.5,   True,         0.25, 16),   # antialias matters
]


def gen_mix(ref_comix, tmp: str) -> None:
    data = {}
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    for name, T, crop, bg_resize, (bh, bw), n_bg, alpha, with_ra, prob, seed in MIX_CASES:
        rng = np.random.default_rng(seed)
        n_samples = 6
        fg = rng.integers(0, 256, (n_samples, T, crop[0], crop[1], 3), dtype=np.uint8)
        ra_flags = rng.integers(0, 2, n_samples).astype(bool)
        pool = rng.integers(0, 256, (n_bg, 3, bh, bw), dtype=np.uint8)
        bg_dir = pathlib.Path(tmp) / ("bg_" + name)
        bg_dir.mkdir()
        names = [f"v{i:03d}" for i in range(n_bg)]
        for n in names:
            (bg_dir / (n + ".jpg")).write_bytes(b"stub")     # existence is all the ctor checks (:93)
        video_infos = [dict(frame_dir=f"/nowhere/{names[i % n_bg]}", total_frames=T, label=i, sample=i)
                       for i in range(n_samples)]

        def pipeline(info, _fg=fg, _ra=ra_flags):
            i = info["sample"]
            return dict(imgs=torch.from_numpy(mmcv_style_normalize(_fg[i], mean, std)),
                        label=torch.tensor([info["label"]]), randAug=bool(_ra[i]))

        ds = ref_comix.BackgroundMixDataset(video_infos, pipeline, bg_dir=str(bg_dir),
                                            bg_resize=bg_resize, bg_crop_size=crop, alpha=alpha,
                                            prob=prob, with_randAug=with_ra)
        # map_bg_to_video=True: one pool entry per video_info whose <bg_dir>/<name>.jpg exists (:88-94)
        assert len(ds.bg_files) == n_samples
        path_to_idx = {str(bg_dir / (n + ".jpg")): i for i, n in enumerate(names)}
        ref_comix.read_image = lambda p, mode=None, _m=path_to_idx, _pool=pool: torch.from_numpy(_pool[_m[p]])

        random.seed(seed)
        torch.manual_seed(seed)
        outs, bg_idx = [], []
        for i in range(n_samples):
            r = ds.prepare_train_frames(i)
            outs.append(r["imgs"].numpy())
            bg_idx.append(int(r["bg_idx"]))
        data[name + "/fg"] = fg
        data[name + "/randAug"] = ra_flags
        data[name + "/pool_u8"] = pool
        data[name + "/bg_files_order"] = np.array([path_to_idx[p] for p in ds.bg_files], np.int64)
        data[name + "/expected"] = np.stack(outs)
        data[name + "/bg_idx"] = np.array(bg_idx, np.int64)
        data[name + "/params"] = np.array([crop[0], crop[1], bg_resize, alpha, float(with_ra), prob, seed],
                                          np.float64)
    np.savez_compressed(GOLDEN / "bgmix_reference.npz", **data)
    print("bgmix_reference.npz:", len(MIX_CASES), "cases,",
          (GOLDEN / "bgmix_reference.npz").stat().st_size, "bytes")

    # one full-size sample (config 5 shape) pinned by digest only; inputs are re-derived from seeds
    T, crop, n_bg = 8, (224, 224), 3
    rng = np.random.default_rng(2024)
    fg = rng.integers(0, 256, (1, T, 224, 224, 3), dtype=np.uint8)
    pool = rng.integers(0, 256, (n_bg, 3, 240, 320), dtype=np.uint8)
    bg_dir = pathlib.Path(tmp) / "bg_full"
    bg_dir.mkdir()
    names = [f"v{i:03d}" for i in range(n_bg)]
    for n in names:
        (bg_dir / (n + ".jpg")).write_bytes(b"stub")
    infos = [dict(frame_dir=f"/nowhere/{names[0]}", total_frames=T, label=0, sample=0)]
    ds = ref_comix.BackgroundMixDataset(
        infos, lambda info: dict(imgs=torch.from_numpy(mmcv_style_normalize(fg[0], mean, std)),
                                 label=torch.tensor([0]), randAug=False),
        bg_dir=str(bg_dir), with_randAug=True)
    path_to_idx = {str(bg_dir / (n + ".jpg")): i for i, n in enumerate(names)}
    ref_comix.read_image = lambda p, mode=None: torch.from_numpy(pool[path_to_idx[p]])
    torch.manual_seed(7)
    r = ds.prepare_train_frames(0)
    out = r["imgs"].numpy()
    np.savez_compressed(GOLDEN / "bgmix_fullsize_digest.npz",
                        sha256=np.frombuffer(hashlib.sha256(out.tobytes()).digest(), np.uint8),
                        bg_idx=np.int64(r["bg_idx"]), seed_data=np.int64(2024), seed_torch=np.int64(7),
                        bg_files_order=np.array([path_to_idx[p] for p in ds.bg_files], np.int64),
                        sample_values=out[::3, ::2, ::37, ::41].copy())
    print("bgmix_fullsize_digest.npz written; bg_idx", int(r["bg_idx"]))


def main() -> None:
    if not _ref_import.available():
        raise SystemExit("reference tree not found at " + _ref_import.REFERENCE_ROOT)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    ref_extract = _ref_import.load_extract_background()
    ref_comix = _ref_import.load_comix_loader()
    with tempfile.TemporaryDirectory() as tmp:
        gen_median(ref_extract, ref_comix, tmp)
        gen_mix(ref_comix, tmp)
    (GOLDEN / "README.md").write_text(
        "# Golden vectors\n\nWritten by `python oracle/gen_golden.py` in the build container, by running the\n"
        "reference's own `bg_extraction_tmf` (both variants) and `BackgroundMixDataset` from\n"
        "`/root/reference` on seeded synthetic inputs.  Each `.npz` holds the inputs and the\n"
        "reference's outputs.  Versions at generation time: numpy %s, cv2 %s, torch %s, torchvision %s.\n"
        % (np.__version__, cv2.__version__, torch.__version__, __import__("torchvision").__version__))


if __name__ == "__main__":
    main()
