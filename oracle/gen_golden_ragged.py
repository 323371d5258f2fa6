#!/usr/bin/env python
"""Generate tests/golden/bgmix_ragged_reference.npz by running the REFERENCE's ``BackgroundMixDataset`` (unmodified).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden_ragged.py

Covers what tests/golden/bgmix_reference.npz does not:

* pools whose images differ in size (HMDB51 / Sth-Sth-v2 style: one height, several widths, plus portrait and
  down-scaled images): the reference resizes and crops each drawn image at its own size
  (libs/loader/comix_loader.py:72-75,126-131,139-141);
* the random-frame mode ``back_ground_from_bg_dir=False`` ("type A", :133-136;
  configs/ucf101/predefined_background/seed_1000_inc_10_stages_bgmix_plus_randAug_type_a_bg.py:197): the background is
  a random frame of a random video, ``bg_idx`` = -2;
* one full-size mixed-width case (240x320 / 240x427 / 240x352 -> Resize(256) -> 224x224 crop), pinned by digest.

As in gen_golden.py the JPEG decode inside ``_get_bg_image`` (``torchvision.io.read_image``) is replaced by a lookup
into seeded synthetic images -- decode is outside the kernel boundary -- and the foreground normalisation is produced
with the cv2 calls mmcv makes.  Everything else is the reference's code and the installed torchvision.
"""
from __future__ import annotations

import hashlib
import pathlib
import random
import sys
import tempfile

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle import _ref_import  # noqa: E402
from oracle.gen_golden import mmcv_style_normalize  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]

POOL_CASES = [
    # name,          T, crop,     bg_resize, bg sizes (h, w) cycled over the pool,                      n_bg, alpha, with_randAug, prob, seed
    ("mixed_widths", 3, (32, 32), 40,        [(36, 48), (36, 64), (36, 54), (36, 48), (36, 72)],         7,    0.5,   True,         0.25, 21),
    ("portrait_mix", 2, (24, 24), 32,        [(60, 50), (30, 50), (40, 40), (33, 47)],                   6,    0.3,   True,         0.25, 22),
    ("down_and_up",  2, (32, 32), 36,        [(90, 120), (36, 48), (100, 250), (30, 31)],                5,    0.7,   False,        0.6,  23),   # prob gate; many taps
    ("skipped_axes", 2, (16, 16), 16,        [(16, 16), (20, 30), (16, 40)],                             4,    0.5,   True,         0.25, 24),   # Resize(16) keeps one or both axes: ATen skips those passes
]

FRAME_CASES = [
    # name,       T, crop,     bg_resize, frame sizes per video,                 frames per video, alpha, with_randAug, prob, seed
    ("type_a",    3, (32, 32), 40,        [(36, 48), (36, 64), (48, 40)],        [5, 3, 4],        0.5,   True,         0.25, 31),
    ("type_a_p",  2, (24, 24), 30,        [(30, 40), (50, 30)],                  [2, 6],           0.4,   False,        0.5,  32),
]


def run_pool_case(ref_comix, tmp, name, T, crop, bg_resize, sizes, n_bg, alpha, with_ra, prob, seed, n_samples=8):
    rng = np.random.default_rng(seed)
    fg = rng.integers(0, 256, (n_samples, T, crop[0], crop[1], 3), dtype=np.uint8)
    ra_flags = rng.integers(0, 2, n_samples).astype(bool)
    pool = [rng.integers(0, 256, (3,) + tuple(sizes[i % len(sizes)]), dtype=np.uint8) for i in range(n_bg)]
    bg_dir = pathlib.Path(tmp) / ("bg_" + name)
    bg_dir.mkdir()
    names = [f"v{i:03d}" for i in range(n_bg)]
    for n in names:
        (bg_dir / (n + ".jpg")).write_bytes(b"stub")
    video_infos = [dict(frame_dir=f"/nowhere/{names[i % n_bg]}", total_frames=T, label=i, sample=i) for i in range(n_samples)]

    def pipeline(info):
        i = info["sample"]
        return dict(imgs=torch.from_numpy(mmcv_style_normalize(fg[i], MEAN, STD)), label=torch.tensor([info["label"]]),
                    randAug=bool(ra_flags[i]))

    ds = ref_comix.BackgroundMixDataset(video_infos, pipeline, bg_dir=str(bg_dir), bg_resize=bg_resize,
                                        bg_crop_size=crop, alpha=alpha, prob=prob, with_randAug=with_ra)
    path_to_idx = {str(bg_dir / (n + ".jpg")): i for i, n in enumerate(names)}
    ref_comix.read_image = lambda p, mode=None: torch.from_numpy(pool[path_to_idx[p]])
    random.seed(seed)
    torch.manual_seed(seed)
    outs, bg_idx = [], []
    for i in range(n_samples):
        r = ds.prepare_train_frames(i)
        outs.append(r["imgs"].numpy())
        bg_idx.append(int(r["bg_idx"]))
    data = {name + "/fg": fg, name + "/randAug": ra_flags, name + "/expected": np.stack(outs),
            name + "/bg_idx": np.array(bg_idx, np.int64),
            name + "/bg_files_order": np.array([path_to_idx[p] for p in ds.bg_files], np.int64),
            name + "/params": np.array([crop[0], crop[1], bg_resize, alpha, float(with_ra), prob, seed], np.float64),
            name + "/n_bg": np.int64(n_bg)}
    for i, im in enumerate(pool):
        data[f"{name}/pool_{i:03d}"] = im
    return data


def run_frame_case(ref_comix, tmp, name, T, crop, bg_resize, sizes, n_frames, alpha, with_ra, prob, seed, n_samples=8):
    rng = np.random.default_rng(seed)
    fg = rng.integers(0, 256, (n_samples, T, crop[0], crop[1], 3), dtype=np.uint8)
    ra_flags = rng.integers(0, 2, n_samples).astype(bool)
    n_vid = len(sizes)
    frames = {}                                              # path -> uint8 [3, h, w]
    video_infos = []
    for v in range(n_vid):
        d = f"/nowhere/{name}/vid{v:02d}"
        for k in range(1, n_frames[v] + 1):
            frames[f"{d}/img_{k:05}.jpg"] = rng.integers(0, 256, (3,) + tuple(sizes[v]), dtype=np.uint8)
    for i in range(n_samples):
        v = i % n_vid
        video_infos.append(dict(frame_dir=f"/nowhere/{name}/vid{v:02d}", total_frames=n_frames[v], label=i, sample=i))

    def pipeline(info):
        i = info["sample"]
        return dict(imgs=torch.from_numpy(mmcv_style_normalize(fg[i], MEAN, STD)), label=torch.tensor([info["label"]]),
                    randAug=bool(ra_flags[i]))

    bg_dir = pathlib.Path(tmp) / ("bg_" + name)
    ds = ref_comix.BackgroundMixDataset(video_infos, pipeline, bg_dir=str(bg_dir), back_ground_from_bg_dir=False,
                                        bg_resize=bg_resize, bg_crop_size=crop, alpha=alpha, prob=prob, with_randAug=with_ra)
    assert ds.bg_files == []
    drawn = []
    ref_comix.read_image = lambda p, mode=None: (drawn.append(p), torch.from_numpy(frames[p]))[1]
    random.seed(seed)
    torch.manual_seed(seed)
    outs, bg_idx = [], []
    for i in range(n_samples):
        r = ds.prepare_train_frames(i)
        outs.append(r["imgs"].numpy())
        bg_idx.append(int(r["bg_idx"]))
    paths = sorted(frames)
    data = {name + "/fg": fg, name + "/randAug": ra_flags, name + "/expected": np.stack(outs),
            name + "/bg_idx": np.array(bg_idx, np.int64),
            name + "/drawn": np.array([paths.index(p) for p in drawn], np.int64),
            name + "/frame_paths": np.array(paths),
            name + "/video_frames": np.array(n_frames, np.int64),
            name + "/video_of_sample": np.array([i % n_vid for i in range(n_samples)], np.int64),
            name + "/params": np.array([crop[0], crop[1], bg_resize, alpha, float(with_ra), prob, seed], np.float64)}
    for i, p in enumerate(paths):
        data[f"{name}/frame_{i:03d}"] = frames[p]
    return data


def run_fullsize(ref_comix, tmp):
    """240x320 / 240x427 / 240x352 backgrounds -> Resize(256) -> RandomCrop(224): pinned by digest and sampled values."""
    T, crop, sizes = 8, (224, 224), [(240, 320), (240, 427), (240, 352), (256, 256)]
    rng = np.random.default_rng(2025)
    n_samples = 4
    fg = rng.integers(0, 256, (n_samples, T, 224, 224, 3), dtype=np.uint8)
    pool = [rng.integers(0, 256, (3,) + s, dtype=np.uint8) for s in sizes]
    bg_dir = pathlib.Path(tmp) / "bg_full_ragged"
    bg_dir.mkdir()
    names = [f"v{i:03d}" for i in range(len(pool))]
    for n in names:
        (bg_dir / (n + ".jpg")).write_bytes(b"stub")
    infos = [dict(frame_dir=f"/nowhere/{names[i % len(pool)]}", total_frames=T, label=i, sample=i) for i in range(n_samples)]
    ds = ref_comix.BackgroundMixDataset(
        infos, lambda info: dict(imgs=torch.from_numpy(mmcv_style_normalize(fg[info["sample"]], MEAN, STD)), label=torch.tensor([0]), randAug=False),
        bg_dir=str(bg_dir), with_randAug=True)
    path_to_idx = {str(bg_dir / (n + ".jpg")): i for i, n in enumerate(names)}
    ref_comix.read_image = lambda p, mode=None: torch.from_numpy(pool[path_to_idx[p]])
    torch.manual_seed(9)
    outs, idx = [], []
    for i in range(n_samples):
        r = ds.prepare_train_frames(i)
        outs.append(r["imgs"].numpy())
        idx.append(int(r["bg_idx"]))
    out = np.stack(outs)
    return {"fullsize/sha256": np.frombuffer(hashlib.sha256(out.tobytes()).digest(), np.uint8),
            "fullsize/bg_idx": np.array(idx, np.int64), "fullsize/seed_data": np.int64(2025), "fullsize/seed_torch": np.int64(9),
            "fullsize/sizes": np.array(sizes, np.int64),
            "fullsize/bg_files_order": np.array([path_to_idx[p] for p in ds.bg_files], np.int64),
            "fullsize/sample_values": out[:, ::3, :, ::37, ::41].copy()}


def main() -> None:
    if not _ref_import.available():
        raise SystemExit("reference tree not found at " + _ref_import.REFERENCE_ROOT)
    ref_comix = _ref_import.load_comix_loader()
    data = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case in POOL_CASES:
            data.update(run_pool_case(ref_comix, tmp, *case))
        for case in FRAME_CASES:
            data.update(run_frame_case(ref_comix, tmp, *case))
        data.update(run_fullsize(ref_comix, tmp))
    np.savez_compressed(GOLDEN / "bgmix_ragged_reference.npz", **data)
    print("bgmix_ragged_reference.npz:", len(POOL_CASES), "pool cases,", len(FRAME_CASES), "random-frame cases + fullsize,",
          (GOLDEN / "bgmix_ragged_reference.npz").stat().st_size, "bytes; torch", torch.__version__,
          "torchvision", __import__("torchvision").__version__)


if __name__ == "__main__":
    main()
