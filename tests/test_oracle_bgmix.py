"""The BG-mix oracle against outputs of the reference's BackgroundMixDataset
(libs/loader/comix_loader.py:16-145), stored in tests/golden/bgmix_reference.npz by
oracle/gen_golden.py."""
import hashlib
import random

import numpy as np
import pytest
import torch

from oracle import bgmix_oracle as bo
from oracle import c_oracle

from conftest import GOLDEN, median_case_names

_NPZ = np.load(GOLDEN / "bgmix_reference.npz")
CASES = median_case_names(_NPZ)


def _replay(name):
    """Re-run the reference's RNG draws with the oracle and rebuild every sample."""
    fg = _NPZ[name + "/fg"]
    ra = _NPZ[name + "/randAug"]
    pool = _NPZ[name + "/pool_u8"]
    order = _NPZ[name + "/bg_files_order"]
    ch, cw, bg_resize, alpha, with_ra, prob, seed = _NPZ[name + "/params"]
    crop = (int(ch), int(cw))
    random.seed(int(seed))
    torch.manual_seed(int(seed))
    outs, idxs = [], []
    for i in range(len(fg)):
        rv = None if with_ra else random.random()
        if bo.gate(bool(with_ra), bool(ra[i]), float(prob), rv):
            h, w = bo.resized_hw(pool.shape[2], pool.shape[3], int(bg_resize))
            bg_idx, top, left = bo.draw_bg_params(len(order), h, w, crop)
            resized = bo.bg_resize(pool[order[bg_idx]], int(bg_resize)).numpy()
            assert resized.shape[1:] == (h, w)
            outs.append(bo.mix_clip(fg[i], resized, top, left, crop, float(alpha), True))
            idxs.append(bg_idx)
        else:
            outs.append(bo.mix_clip(fg[i], None, 0, 0, crop, float(alpha), False))
            idxs.append(-1)
    return np.stack(outs), np.array(idxs)


@pytest.mark.parametrize("name", CASES)
def test_oracle_bit_exact_vs_reference(name):
    got, idx = _replay(name)
    np.testing.assert_array_equal(idx, _NPZ[name + "/bg_idx"])
    exp = _NPZ[name + "/expected"]
    assert got.dtype == np.float32 and got.shape == exp.shape
    np.testing.assert_array_equal(got.view(np.uint32), exp.view(np.uint32))   # bit-exact


def test_some_samples_mixed_and_some_not():
    mixed = [(_NPZ[n + "/bg_idx"] >= 0).sum() for n in CASES]
    plain = [(_NPZ[n + "/bg_idx"] == -1).sum() for n in CASES]
    assert sum(mixed) > 0 and sum(plain) > 0


def test_fg_lut_matches_cv2():
    np.testing.assert_array_equal(bo.fg_lut().view(np.uint32), bo.fg_lut_cv2().view(np.uint32))
    m, s = (10.5, 200.25, 0.0), (1.0, 33.3, 255.0)
    np.testing.assert_array_equal(bo.fg_lut(m, s).view(np.uint32), bo.fg_lut_cv2(m, s).view(np.uint32))


def test_c_blend_matches_numpy_blend():
    rng = np.random.default_rng(3)
    fg = rng.integers(0, 256, (3, 9, 11, 3), dtype=np.uint8)
    bg = rng.uniform(0, 255, (3, 9, 11)).astype(np.float32)
    for alpha in (0.5, 0.3, 0.7, 0.25):
        a = bo.blend(bo.fg_normalize(fg, bo.fg_lut()), bo.bg_normalize(bg), alpha)
        b = c_oracle.bgmix_clip(fg, bg, bo.fg_lut(), bo.DEFAULT_MEAN, bo.DEFAULT_STD, alpha, True)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    c = c_oracle.bgmix_clip(fg, bg, bo.fg_lut(), bo.DEFAULT_MEAN, bo.DEFAULT_STD, 0.5, False)
    np.testing.assert_array_equal(c, bo.fg_normalize(fg, bo.fg_lut()))


def test_fullsize_digest():
    """Config-5-shaped sample [8,3,224,224]: inputs re-derived from seeds, output pinned by sha256."""
    d = np.load(GOLDEN / "bgmix_fullsize_digest.npz")
    rng = np.random.default_rng(int(d["seed_data"]))
    fg = rng.integers(0, 256, (1, 8, 224, 224, 3), dtype=np.uint8)
    pool = rng.integers(0, 256, (3, 3, 240, 320), dtype=np.uint8)
    order = d["bg_files_order"]
    torch.manual_seed(int(d["seed_torch"]))
    h, w = bo.resized_hw(240, 320, 256)
    assert (h, w) == (256, 341)
    bg_idx, top, left = bo.draw_bg_params(len(order), h, w, (224, 224))
    assert bg_idx == int(d["bg_idx"])
    out = bo.mix_clip(fg[0], bo.bg_resize(pool[order[bg_idx]], 256).numpy(), top, left)
    np.testing.assert_array_equal(out[::3, ::2, ::37, ::41], d["sample_values"])
    assert hashlib.sha256(out.tobytes()).digest() == bytes(d["sha256"])


def test_layouts():
    rng = np.random.default_rng(8)
    fg = rng.integers(0, 256, (2, 3, 8, 8, 3), dtype=np.uint8)
    pool = rng.uniform(0, 255, (2, 3, 12, 12)).astype(np.float32)
    a = bo.mix_batch(fg, pool, [1, 0], [2, 0], [1, 3], [1, 0], crop=(8, 8))
    b = bo.mix_batch(fg, pool, [1, 0], [2, 0], [1, 3], [1, 0], crop=(8, 8), layout="NCTHW")
    assert a.shape == (2, 3, 3, 8, 8) and b.shape == (2, 3, 3, 8, 8)
    np.testing.assert_array_equal(a.transpose(0, 2, 1, 3, 4), b)


def test_restated_arithmetic_equals_the_library_calls_of_the_reference():
    """mix_clip (the arithmetic the CUDA kernel implements, restated operation by operation) is bit-identical to
    the reference's own expression evaluated with torchvision + torch (comix_loader.py:72-75,139-142)."""
    import numpy as np
    import torch
    from oracle import bgmix_oracle as bo
    rs = np.random.default_rng(3)
    fg_u8 = rs.integers(0, 256, (4, 224, 224, 3), dtype=np.uint8)
    bg = rs.integers(0, 256, (3, 240, 320), dtype=np.uint8)
    fgn = torch.from_numpy(bo.fg_normalize(fg_u8, bo.fg_lut()))
    for alpha in (0.5, 0.3):
        torch.manual_seed(7)
        a = bo.mix_clip_like_reference(fgn, torch.from_numpy(bg), alpha).numpy()
        torch.manual_seed(7)
        r = bo.bg_resize(bg, 256).numpy()
        top = int(torch.randint(0, r.shape[1] - 224 + 1, (1,)))
        left = int(torch.randint(0, r.shape[2] - 224 + 1, (1,)))
        b = bo.mix_clip(fg_u8, r, top, left, alpha=alpha)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
