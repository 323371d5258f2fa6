"""CPU tests of the background-Resize oracle (oracle/aa_resize_oracle.py), of the host-side weight tables of the C ABI
(bgd_aa_resize_table needs no GPU) and of the oracle's replay of the reference's ragged-pool / random-frame outputs
(tests/golden/bgmix_ragged_reference.npz, written by oracle/gen_golden_ragged.py from the unmodified reference)."""
import ctypes
import hashlib
import pathlib
import random

import numpy as np
import pytest
import torch

from oracle import aa_resize_oracle as ao, bgmix_oracle as bo

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "bgmix_ragged_reference.npz"

SIZES = [(240, 320, 256), (240, 427, 256), (240, 426, 256), (36, 48, 40), (36, 64, 40), (60, 50, 48), (90, 120, 36),
         (100, 250, 36), (256, 256, 256), (360, 480, 256), (480, 640, 256), (100, 176, 256), (16, 40, 16), (20, 30, 16),
         (33, 47, 32), (1080, 1920, 256), (50, 30, 30)]


@pytest.mark.parametrize("h,w,size", SIZES)
def test_oracle_equals_torchvision_resize_bit_for_bit(h, w, size):
    """The executable third-party reference: torchvision's Resize on read_image(...).float() (comix_loader.py:72,130)."""
    from torchvision.transforms import Resize
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    ref = Resize(size)(torch.from_numpy(img).float()).numpy()
    got = ao.aa_resize(img, size)
    assert got.shape == ref.shape == (3,) + ao.resized_hw(h, w, size)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def _cabi_table(cabi, n_in, n_out):
    taps = ctypes.c_int32()
    cabi.check(cabi.lib().bgd_aa_resize_table(n_in, n_out, ctypes.byref(taps), None, 0))
    K = taps.value
    words = np.empty(n_out * (2 + K), np.int32)
    cabi.check(cabi.lib().bgd_aa_resize_table(n_in, n_out, ctypes.byref(taps), words.ctypes.data, words.size))
    words = words.reshape(n_out, 2 + K)
    return words[:, 0].astype(np.int64), words[:, 1].astype(np.int64), words[:, 2:].copy().view(np.float32)


@pytest.mark.parametrize("n_in,n_out", [(320, 341), (427, 455), (240, 256), (120, 48), (90, 36), (1000, 256), (341, 100),
                                       (2000, 100), (48, 53), (7, 5), (5, 7), (1, 3), (3, 1), (250, 90)])
def test_cabi_weight_tables_equal_oracle_and_torch(n_in, n_out):
    """bgd_aa_resize_table (host code of the library) against the oracle's tables and against the weights torch itself
    applies: the impulse responses of F.interpolate(antialias=True) along one axis."""
    import __graft_entry__ as g
    g.build()
    from bgdebias_b200 import _cabi
    mins, sizes, W = _cabi_table(_cabi, n_in, n_out)
    omins, osizes, oW = ao.aa_tables(n_in, n_out)
    assert np.array_equal(mins, omins) and np.array_equal(sizes, osizes)
    assert W.shape == oW.shape and np.array_equal(W.view(np.uint32), oW.view(np.uint32))
    eye = torch.eye(n_in).view(n_in, 1, 1, n_in)
    Wt = torch.nn.functional.interpolate(eye, size=(1, n_out), mode="bilinear", antialias=True, align_corners=False).view(n_in, n_out).numpy()
    M = np.zeros((n_in, n_out), np.float32)
    for i in range(n_out):
        for j in range(sizes[i]):
            M[mins[i] + j, i] = W[i, j]
    assert np.array_equal(M.view(np.uint32), Wt.view(np.uint32))


def test_cabi_table_rejects_bad_sizes():
    import __graft_entry__ as g
    g.build()
    from bgdebias_b200 import _cabi
    taps = ctypes.c_int32()
    assert _cabi.lib().bgd_aa_resize_table(0, 4, ctypes.byref(taps), None, 0) == _cabi.BGD_ERR_INVALID
    words = np.empty(4, np.int32)
    assert _cabi.lib().bgd_aa_resize_table(10, 20, ctypes.byref(taps), words.ctypes.data, 4) == _cabi.BGD_ERR_INVALID


# --------------------------------------------------------------------------------------------------------------------
def replay_pool_case(d, name):
    """The reference's draws (comix_loader.py:105-145) replayed with the oracle's arithmetic."""
    fg, ra, order = d[name + "/fg"], d[name + "/randAug"], d[name + "/bg_files_order"]
    th, tw, size, alpha, with_ra, prob, seed = d[name + "/params"]
    crop, size, seed = (int(th), int(tw)), int(size), int(seed)
    pool = [d[f"{name}/pool_{i:03d}"] for i in range(int(d[name + "/n_bg"]))]
    random.seed(seed)
    torch.manual_seed(seed)
    outs, idxs = [], []
    for i in range(len(fg)):
        fgn = bo.fg_normalize(fg[i], bo.fg_lut())
        mixed = bo.gate(bool(with_ra), bool(ra[i]), prob, random.random() if not with_ra else 0.0)
        if not mixed:
            outs.append(fgn); idxs.append(-1)
            continue
        img = None
        bg_idx = int(torch.randint(len(order), (1,)).item())
        img = pool[order[bg_idx]]
        H, W = ao.resized_hw(img.shape[1], img.shape[2], size)
        top, left = (0, 0) if (H, W) == crop else (int(torch.randint(0, H - crop[0] + 1, (1,)).item()), int(torch.randint(0, W - crop[1] + 1, (1,)).item()))
        bgc = ao.aa_resize(img, size)[:, top:top + crop[0], left:left + crop[1]]
        outs.append(bo.blend(fgn, bo.bg_normalize(bgc), alpha)); idxs.append(bg_idx)
    return np.stack(outs), idxs


def replay_frame_case(d, name):
    fg, ra = d[name + "/fg"], d[name + "/randAug"]
    th, tw, size, alpha, with_ra, prob, seed = d[name + "/params"]
    crop, size, seed = (int(th), int(tw)), int(size), int(seed)
    paths = [str(p) for p in d[name + "/frame_paths"]]
    frames = {p: d[f"{name}/frame_{i:03d}"] for i, p in enumerate(paths)}
    n_frames, vid_of = d[name + "/video_frames"], d[name + "/video_of_sample"]
    infos = [dict(frame_dir=f"/nowhere/{name}/vid{int(v):02d}", total_frames=int(n_frames[v])) for v in vid_of]
    random.seed(seed)
    torch.manual_seed(seed)
    outs, idxs, drawn = [], [], []
    for i in range(len(fg)):
        fgn = bo.fg_normalize(fg[i], bo.fg_lut())
        mixed = bo.gate(bool(with_ra), bool(ra[i]), prob, random.random() if not with_ra else 0.0)
        if not mixed:
            outs.append(fgn); idxs.append(-1)
            continue
        video = random.choice(infos)                                    # comix_loader.py:133
        k = random.randint(1, video["total_frames"] - 1 + 1)            # :134, start_index = 1
        p = f"{video['frame_dir']}/img_{k:05}.jpg"
        drawn.append(paths.index(p))
        img = frames[p]
        H, W = ao.resized_hw(img.shape[1], img.shape[2], size)
        top, left = (0, 0) if (H, W) == crop else (int(torch.randint(0, H - crop[0] + 1, (1,)).item()), int(torch.randint(0, W - crop[1] + 1, (1,)).item()))
        bgc = ao.aa_resize(img, size)[:, top:top + crop[0], left:left + crop[1]]
        outs.append(bo.blend(fgn, bo.bg_normalize(bgc), alpha)); idxs.append(-2)
    return np.stack(outs), idxs, drawn


@pytest.mark.parametrize("name", ["mixed_widths", "portrait_mix", "down_and_up", "skipped_axes"])
def test_oracle_replays_reference_mixed_size_pools(name):
    d = np.load(GOLDEN)
    out, idxs = replay_pool_case(d, name)
    assert idxs == d[name + "/bg_idx"].tolist()
    assert np.array_equal(out.view(np.uint32), d[name + "/expected"].view(np.uint32))


@pytest.mark.parametrize("name", ["type_a", "type_a_p"])
def test_oracle_replays_reference_random_frame_mode(name):
    d = np.load(GOLDEN)
    out, idxs, drawn = replay_frame_case(d, name)
    assert idxs == d[name + "/bg_idx"].tolist() and drawn == d[name + "/drawn"].tolist()
    assert np.array_equal(out.view(np.uint32), d[name + "/expected"].view(np.uint32))


def fullsize_inputs(d):
    rng = np.random.default_rng(int(d["fullsize/seed_data"]))
    fg = rng.integers(0, 256, (4, 8, 224, 224, 3), dtype=np.uint8)
    pool = [rng.integers(0, 256, (3,) + tuple(int(v) for v in s), dtype=np.uint8) for s in d["fullsize/sizes"]]
    return fg, pool


def test_oracle_fullsize_mixed_width_digest():
    """240x320 / 240x427 / 240x352 / 256x256 backgrounds -> Resize(256) -> RandomCrop(224): the reference's bytes."""
    d = np.load(GOLDEN)
    fg, pool = fullsize_inputs(d)
    order = d["fullsize/bg_files_order"]
    torch.manual_seed(int(d["fullsize/seed_torch"]))
    outs, idxs = [], []
    for i in range(4):
        bg_idx = int(torch.randint(len(order), (1,)).item())
        img = pool[order[bg_idx]]
        H, W = ao.resized_hw(img.shape[1], img.shape[2], 256)
        top, left = (0, 0) if (H, W) == (224, 224) else (int(torch.randint(0, H - 223, (1,)).item()), int(torch.randint(0, W - 223, (1,)).item()))
        bgc = ao.aa_resize(img, 256)[:, top:top + 224, left:left + 224]
        outs.append(bo.blend(bo.fg_normalize(fg[i], bo.fg_lut()), bo.bg_normalize(bgc), 0.5)); idxs.append(bg_idx)
    out = np.stack(outs)
    assert idxs == d["fullsize/bg_idx"].tolist()
    assert np.array_equal(out[:, ::3, :, ::37, ::41].view(np.uint32), d["fullsize/sample_values"].view(np.uint32))
    assert hashlib.sha256(out.tobytes()).digest() == d["fullsize/sha256"].tobytes()
