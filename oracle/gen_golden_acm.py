#!/usr/bin/env python
"""Generate tests/golden/acm_reference.npz by running the REFERENCE's own ``ActorCutMixDataset.actor_cut_mix``
(libs/loader/actor_cut_mix_loader.py:135-152), unmodified, with stub pipelines that feed it seeded frames.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/gen_golden_acm.py
"""
from __future__ import annotations

import pathlib
import random
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle import _ref_import, acm_oracle as ao  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def main():
    mod = _ref_import.load_actor_cut_mix_loader()
    cls = mod.ActorCutMixDataset
    out = {}
    for name, T, H, W, seed, boxes in [("boxes", 4, 40, 56, 1, 2), ("no_person", 3, 32, 32, 2, 0), ("all_person", 2, 24, 40, 3, -1)]:
        rng = np.random.default_rng(seed)
        actor = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
        scene = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
        masks = np.zeros((T, H, W, 3), np.uint8)
        if boxes < 0:
            masks[:] = 1                                   # box.py:187: no detections -> the whole frame is foreground
        for t in range(T):
            for _ in range(max(boxes, 0)):
                y0, x0 = int(rng.integers(0, H - 4)), int(rng.integers(0, W - 4))
                masks[t, y0:y0 + int(rng.integers(2, H // 2)), x0:x0 + int(rng.integers(2, W // 2)), :] = 1
        ds = cls.__new__(cls)
        ds.video_infos = [dict(frame_dir=f"/x/v{i}", total_frames=T, label=10 + i) for i in range(5)]
        ds.filename_tmpl, ds.modality, ds.start_index = "img_{:05}.jpg", "RGB", 1
        ds.action_pipeline = lambda r, a=actor, m=masks: dict(r, imgs=[f.copy() for f in a], human_mask=[k.copy() for k in m])
        ds.scene_pipeline = lambda r, s=scene: dict(r, imgs=[f.copy() for f in s])
        random.seed(seed)
        res = ds.actor_cut_mix(ds._prepare_frames(0))
        mixed = np.stack(res["imgs"])
        assert np.array_equal(mixed, ao.cut_mix(actor, masks, scene)), name
        assert float(res["foreground_ratio"]) == ao.foreground_ratio(masks), name
        out[f"{name}/actor"], out[f"{name}/scene"], out[f"{name}/mask"] = actor, scene, masks
        out[f"{name}/expected"] = mixed
        out[f"{name}/foreground_ratio"] = np.float64(res["foreground_ratio"])
        out[f"{name}/background_label"] = np.int64(res["background_label"])
        out[f"{name}/seed"] = np.int64(seed)
        print(name, mixed.shape, float(res["foreground_ratio"]), int(res["background_label"]))
    np.savez_compressed(GOLDEN / "acm_reference.npz", **out)


if __name__ == "__main__":
    main()
