"""Writes tests/golden/resize_reference.npz: outputs of cv2.resize(..., INTER_LINEAR) -- the call mmaction's Resize
(config ..._bgmix_plus_randAug.py:136) bottoms out in -- on seeded uint8 images.  Run in the build container:
    python oracle/gen_golden_resize.py
"""
import pathlib

import cv2
import numpy as np

OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "resize_reference.npz"

# (src_h, src_w, dst_h, dst_w): MultiScaleCrop shapes at reduced scale, the real 168x192 -> 224x224, edge cases
CASES = [(168, 192, 224, 224), (256, 224, 224, 224), (64, 48, 56, 56), (42, 56, 56, 56), (48, 48, 56, 56), (112, 112, 56, 56),
         (1, 1, 8, 8), (1, 7, 5, 9), (9, 1, 4, 6), (2, 2, 7, 7), (37, 53, 24, 40), (100, 37, 224, 224), (31, 31, 31, 31)]


def main():
    rng = np.random.default_rng(20240607)
    data = {"versions": np.array([f"cv2 {cv2.__version__}", f"numpy {np.__version__}"])}
    for i, (sh, sw, dh, dw) in enumerate(CASES):
        if i % 2 == 0:
            img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        else:                                             # smooth ramps + noise: neighbouring levels, rounding cases
            yy, xx = np.mgrid[0:sh, 0:sw]
            img = np.stack([(yy * 3 + xx * 2) % 256, (xx * 5) % 256, (yy * 7 + 11) % 256], -1).astype(np.uint8)
            img ^= rng.integers(0, 4, img.shape, dtype=np.uint8)
        data[f"case{i}/src"] = img
        data[f"case{i}/dst"] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(OUT, **data)
    print(OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
