"""The resize oracle (oracle/resize_oracle.py) against cv2.resize -- the call the reference pipeline's
`Resize(scale=(224, 224), keep_ratio=False)` (config ..._bgmix_plus_randAug.py:136) bottoms out in -- and against the
outputs recorded in tests/golden/resize_reference.npz."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import resize_oracle as ro

_NPZ = np.load(GOLDEN / "resize_reference.npz")
_CASES = sorted({k.split("/")[0] for k in _NPZ.files if k.startswith("case")}, key=lambda s: int(s[4:]))


@pytest.mark.parametrize("case", _CASES)
def test_oracle_matches_recorded_cv2_outputs(case):
    src, dst = _NPZ[case + "/src"], _NPZ[case + "/dst"]
    np.testing.assert_array_equal(ro.resize_linear_u8(src, dst.shape[0], dst.shape[1]), dst)


def test_oracle_matches_cv2_on_multiscale_crop_shapes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    shapes = ro.multiscale_crop_sizes()
    assert (256, 256) in shapes and (168, 192) in shapes and (256, 168) not in shapes
    for cw, ch in shapes:
        img = rng.integers(0, 256, (ch, cw, 3), dtype=np.uint8)
        np.testing.assert_array_equal(ro.resize_linear_u8(img, 224, 224), cv2.resize(img, (224, 224), interpolation=cv2.INTER_LINEAR))


def test_oracle_matches_cv2_on_random_shapes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(150):
        sh, sw, dh, dw = (int(v) for v in rng.integers(1, 300, 4))
        cn = int(rng.choice([1, 3]))
        img = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
        exp = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR).reshape(dh, dw, cn)
        np.testing.assert_array_equal(ro.resize_linear_u8(img, dh, dw), exp, err_msg=f"{(sh, sw)} -> {(dh, dw)} x{cn}")


def test_integer_downscales_and_identity():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for sh, sw, dh, dw in [(448, 448, 224, 224), (672, 448, 224, 224), (896, 672, 224, 224), (224, 224, 224, 224), (112, 56, 224, 224)]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        np.testing.assert_array_equal(ro.resize_linear_u8(img, dh, dw), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
    img = rng.integers(0, 256, (31, 17, 3), dtype=np.uint8)
    np.testing.assert_array_equal(ro.resize_linear_u8(img, 31, 17), img)


def test_coefficients_sum_and_range():
    for src, dst in [(256, 224), (168, 224), (1, 9), (300, 7)]:
        for clamp in (True, False):
            s, c0, c1 = ro.linear_coeffs(src, dst, clamp)
            assert ((c0 >= 0) & (c1 >= 0) & (c0 <= 2048) & (c1 <= 2048)).all()
            assert (np.abs(c0 + c1 - 2048) <= 1).all()
            assert s.min() >= (0 if clamp else -1) and s.max() <= src - 1
