"""Parity of the CUDA BG-mix blend (torch.ops.bgdebias.bgmix_blend*, the drop-in dataset and the
C-ABI host entry point) with the oracle and with the reference's own outputs
(tests/golden/bgmix_reference.npz = BackgroundMixDataset of libs/loader/comix_loader.py run on
seeded inputs).  Tolerance from BASELINE.json: 1e-6 relative (fp32); every fp32 operation in the
kernel is IEEE-rounded in the reference's order, so the tests also assert bit equality."""
import ctypes
import hashlib
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN, median_case_names            # noqa: E402
from oracle import bgmix_oracle as bo                     # noqa: E402

_NPZ = np.load(GOLDEN / "bgmix_reference.npz")
CASES = median_case_names(_NPZ)
RTOL = ATOL = 1e-6


@pytest.fixture(scope="module")
def env():
    import bgdebias_b200.ops as ops
    from bgdebias_b200 import _cabi, comix_loader, pool
    assert torch.cuda.is_available()
    _cabi.lib()
    return ops, _cabi, comix_loader, pool


def _assert_same(got, exp):
    np.testing.assert_allclose(got, exp, rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(got.view(np.uint32), exp.view(np.uint32))


def _params(name):
    ch, cw, bg_resize, alpha, with_ra, prob, seed = _NPZ[name + "/params"]
    return (int(ch), int(cw)), int(bg_resize), float(alpha), bool(with_ra), float(prob), int(seed)


@pytest.mark.parametrize("name", CASES)
def test_batch_op_matches_reference_outputs(env, name):
    """Replay the reference's RNG draws on the host, blend the whole batch with one launch."""
    ops, cabi, cl, pool_mod = env
    crop, bg_resize, alpha, with_ra, prob, seed = _params(name)
    fg, ra, pool_u8, order = (_NPZ[name + k] for k in ("/fg", "/randAug", "/pool_u8", "/bg_files_order"))
    pool = pool_mod.BackgroundPool.from_images([pool_u8[i] for i in order], None, bg_resize, "cuda")
    random.seed(seed)
    torch.manual_seed(seed)
    idx, top, left, app = [], [], [], []
    for i in range(len(fg)):
        rv = None if with_ra else random.random()
        if bo.gate(with_ra, bool(ra[i]), prob, rv):
            b, t, l = bo.draw_bg_params(len(order), *pool.hw, crop)
            idx.append(b); top.append(t); left.append(l); app.append(1)
        else:
            idx.append(0); top.append(0); left.append(0); app.append(0)
    dev = torch.device("cuda")
    t32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    out = torch.ops.bgdebias.bgmix_blend(
        torch.from_numpy(fg).to(dev), pool.tensor, t32(idx), t32(top), t32(left),
        torch.tensor(app, dtype=torch.uint8, device=dev), ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev),
        torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), alpha, "NTCHW").cpu().numpy()
    _assert_same(out, _NPZ[name + "/expected"])
    exp_idx = _NPZ[name + "/bg_idx"]
    np.testing.assert_array_equal(np.where(np.array(app) == 1, idx, -1), exp_idx)


@pytest.mark.parametrize("name", CASES)
def test_dropin_dataset_matches_reference_outputs(env, name):
    """Our BackgroundMixDataset, driven like the reference's (same kwargs, same seeds)."""
    ops, cabi, cl, pool_mod = env
    crop, bg_resize, alpha, with_ra, prob, seed = _params(name)
    fg, ra, pool_u8, order = (_NPZ[name + k] for k in ("/fg", "/randAug", "/pool_u8", "/bg_files_order"))
    import tempfile, pathlib
    with tempfile.TemporaryDirectory() as tmp:
        bg_dir = pathlib.Path(tmp) / "bg"
        bg_dir.mkdir()
        n_bg = len(pool_u8)
        names = [f"v{i:03d}" for i in range(n_bg)]
        for n in names:
            (bg_dir / (n + ".jpg")).write_bytes(b"stub")
        infos = [dict(frame_dir=f"/nowhere/{names[i % n_bg]}", total_frames=fg.shape[1], label=i, sample=i)
                 for i in range(len(fg))]
        lut = bo.fg_lut()

        def pipeline(info):
            i = info["sample"]
            return dict(imgs=torch.from_numpy(bo.fg_normalize(fg[i], lut)), label=torch.tensor([info["label"]]),
                        randAug=bool(ra[i]))

        path_to_idx = {str((bg_dir / (n + ".jpg")).resolve()): i for i, n in enumerate(names)}
        ds = cl.BackgroundMixDataset(infos, pipeline, bg_dir=str(bg_dir), bg_resize=bg_resize, bg_crop_size=crop,
                                     alpha=alpha, prob=prob, with_randAug=with_ra,
                                     bg_reader=lambda p: pool_u8[path_to_idx[str(pathlib.Path(p).resolve())]])
        assert [path_to_idx[str(pathlib.Path(p).resolve())] for p in ds.bg_files] == list(order)
        random.seed(seed)
        torch.manual_seed(seed)
        outs, idxs = [], []
        for i in range(len(fg)):
            r = ds.prepare_train_frames(i)
            assert r["imgs"].dtype == torch.float32 and tuple(r["imgs"].shape) == (fg.shape[1], 3) + crop
            outs.append(r["imgs"].numpy()); idxs.append(int(r["bg_idx"]))
    np.testing.assert_array_equal(idxs, _NPZ[name + "/bg_idx"])
    _assert_same(np.stack(outs), _NPZ[name + "/expected"])


def test_device_mix_collate(env):
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(5)
    T, H, W, n_bg, B = 4, 32, 32, 7, 9
    fg = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    pool_u8 = rng.integers(0, 256, (n_bg, 3, 36, 48), dtype=np.uint8)
    import tempfile, pathlib
    with tempfile.TemporaryDirectory() as tmp:
        bg_dir = pathlib.Path(tmp)
        names = [f"v{i:03d}" for i in range(n_bg)]
        for n in names:
            (bg_dir / (n + ".jpg")).write_bytes(b"stub")
        infos = [dict(frame_dir=f"/x/{names[i % n_bg]}", total_frames=T, label=i, sample=i) for i in range(B)]
        ra = rng.integers(0, 2, B).astype(bool)
        pipeline = lambda info: dict(imgs=torch.from_numpy(fg[info["sample"]]), label=torch.tensor([info["label"]]),
                                     randAug=bool(ra[info["sample"]]))
        path_to_idx = {str((bg_dir / (n + ".jpg")).resolve()): i for i, n in enumerate(names)}
        ds = cl.BackgroundMixDataset(infos, pipeline, bg_dir=str(bg_dir), bg_resize=40, bg_crop_size=(H, W),
                                     with_randAug=True, device_mix=True,
                                     bg_reader=lambda p: pool_u8[path_to_idx[str(pathlib.Path(p).resolve())]])
        torch.manual_seed(3)
        samples = [ds.prepare_train_frames(i) for i in range(B)]
        batch = ds.gpu_collate(samples)
    assert batch["imgs"].is_cuda and tuple(batch["imgs"].shape) == (B, T, 3, H, W)
    order = [path_to_idx[str(pathlib.Path(p).resolve())] for p in ds.bg_files]
    resized = np.stack([bo.bg_resize(pool_u8[i], 40).numpy() for i in order])
    exp = bo.mix_batch(fg, resized, [max(s["bg_idx"], 0) for s in samples], [s["bg_top"] for s in samples],
                       [s["bg_left"] for s in samples], [s["bg_apply"] for s in samples], crop=(H, W))
    _assert_same(batch["imgs"].cpu().numpy(), exp)
    assert [int(v) for v in batch["bg_idx"]] == [s["bg_idx"] for s in samples]
    assert all((s["bg_idx"] == -1) == bool(r) for s, r in zip(samples, ra))


def test_fullsize_digest(env):
    """Config-5-shaped sample, inputs from seeds, output pinned by the reference's sha256."""
    ops, cabi, cl, pool_mod = env
    d = np.load(GOLDEN / "bgmix_fullsize_digest.npz")
    rng = np.random.default_rng(int(d["seed_data"]))
    fg = rng.integers(0, 256, (1, 8, 224, 224, 3), dtype=np.uint8)
    pool_u8 = rng.integers(0, 256, (3, 3, 240, 320), dtype=np.uint8)
    order = d["bg_files_order"]
    pool = pool_mod.BackgroundPool.from_images([pool_u8[i] for i in order], None, 256, "cuda")
    assert pool.hw == (256, 341)
    torch.manual_seed(int(d["seed_torch"]))
    bg_idx, top, left = bo.draw_bg_params(len(order), 256, 341, (224, 224))
    dev = torch.device("cuda")
    t32 = lambda a: torch.tensor([a], dtype=torch.int32, device=dev)
    out = torch.ops.bgdebias.bgmix_blend(
        torch.from_numpy(fg).to(dev), pool.tensor, t32(bg_idx), t32(top), t32(left),
        torch.ones(1, dtype=torch.uint8, device=dev), ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev),
        torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), 0.5, "NTCHW")[0].cpu().numpy()
    assert hashlib.sha256(out.tobytes()).digest() == bytes(d["sha256"])


@pytest.mark.parametrize("layout", ["NTCHW", "NCTHW"])
@pytest.mark.parametrize("pool_dtype", ["f32", "u8"])
@pytest.mark.parametrize("hw", [(224, 224), (30, 27)])
def test_config5_shape_layouts_and_pools(env, layout, pool_dtype, hw):
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(4)
    B, T, (H, W) = 6, 8, hw
    Hb, Wb = H + 32, W + 117
    fg = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    if pool_dtype == "f32":
        pool = rng.uniform(0, 255, (5, 3, Hb, Wb)).astype(np.float32)
    else:
        pool = rng.integers(0, 256, (5, 3, Hb, Wb), dtype=np.uint8)
    idx = rng.integers(0, 5, B); top = rng.integers(0, 33, B); left = rng.integers(0, 118, B)
    app = (rng.uniform(size=B) < 0.6).astype(np.uint8); app[0] = 1; app[1] = 0
    exp = bo.mix_batch(fg, pool.astype(np.float32), idx, top, left, app, crop=(H, W), alpha=0.3, layout=layout)
    dev = torch.device("cuda")
    t32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    out = torch.ops.bgdebias.bgmix_blend(
        torch.from_numpy(fg).to(dev), torch.from_numpy(pool).to(dev), t32(idx), t32(top), t32(left),
        torch.from_numpy(app).to(dev), ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev),
        torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), 0.3, layout).cpu().numpy()
    _assert_same(out, exp)


def test_normfg_op_and_transform(env):
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(9)
    B, T, H, W = 3, 4, 28, 36
    fgn = rng.normal(0, 1.2, (B, T, 3, H, W)).astype(np.float32)
    pool = rng.uniform(0, 255, (4, 3, 40, 50)).astype(np.float32)
    idx, top, left, app = [3, 1, 0], [0, 12, 5], [14, 0, 7], [1, 1, 0]
    dev = torch.device("cuda")
    t32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    out = torch.ops.bgdebias.bgmix_blend_normfg(
        torch.from_numpy(fgn).to(dev), torch.from_numpy(pool).to(dev), t32(idx), t32(top), t32(left),
        torch.tensor(app, dtype=torch.uint8, device=dev), torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD),
        0.7, "NTCHW").cpu().numpy()
    for b in range(B):
        if app[b]:
            bgn = bo.bg_normalize(pool[idx[b]][:, top[b]:top[b] + H, left[b]:left[b] + W])
            _assert_same(out[b], bo.blend(fgn[b], bgn, 0.7))
        else:
            np.testing.assert_array_equal(out[b], fgn[b])
    # pipeline transform: same draws as the dataset's gate
    tr = cl.BackgroundMix([f"bg{i}" for i in range(4)], bg_resize=None, bg_crop_size=(H, W), alpha=0.7,
                          with_randAug=True, bg_reader=lambda p: pool[int(p[2:])])
    torch.manual_seed(1)
    res = tr(dict(imgs=torch.from_numpy(fgn[0]), randAug=False))
    torch.manual_seed(1)
    bi = int(torch.randint(4, (1,)).item()); tp, lf = cl.draw_crop(40, 50, (H, W))
    assert res["bg_idx"] == bi
    _assert_same(res["imgs"].numpy(), bo.blend(fgn[0], bo.bg_normalize(pool[bi][:, tp:tp + H, lf:lf + W]), 0.7))
    assert tr(dict(imgs=torch.from_numpy(fgn[0]), randAug=True))["bg_idx"] == -1


def test_batch_equals_per_sample(env):
    """Config 5 batch size (64 x 8 x 224 x 224): every sample of the batched launch equals its B=1 launch."""
    ops, cabi, cl, pool_mod = env
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    B, T, H, W, P = 64, 8, 224, 224, 16
    fg = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
    pool = torch.rand((P, 3, 256, 341), device=dev, generator=g) * 255
    idx = torch.randint(0, P, (B,), device=dev, generator=g).int()
    top = torch.randint(0, 33, (B,), device=dev, generator=g).int()
    left = torch.randint(0, 118, (B,), device=dev, generator=g).int()
    app = (torch.rand(B, device=dev, generator=g) < 0.25).to(torch.uint8)
    lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
    m, s = torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD)
    full = torch.ops.bgdebias.bgmix_blend(fg, pool, idx, top, left, app, lut, m, s, 0.5, "NTCHW")
    for b in (0, 1, 17, 63):
        one = torch.ops.bgdebias.bgmix_blend(fg[b:b + 1], pool, idx[b:b + 1], top[b:b + 1], left[b:b + 1],
                                             app[b:b + 1], lut, m, s, 0.5, "NTCHW")
        assert torch.equal(one[0], full[b])
    exp0 = bo.mix_clip(fg[0].cpu().numpy(), pool[int(idx[0])].cpu().numpy(), int(top[0]), int(left[0]),
                       apply=bool(app[0]))
    _assert_same(full[0].cpu().numpy(), exp0)


def test_host_entry_point_and_errors(env):
    ops, cabi, cl, pool_mod = env
    L = cabi.lib()
    rng = np.random.default_rng(2)
    B, T, H, W, P, Hb, Wb = 4, 3, 16, 20, 3, 24, 28
    fg = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    pool = rng.uniform(0, 255, (P, 3, Hb, Wb)).astype(np.float32)
    idx = np.array([2, 0, 1, 1], np.int32); top = np.array([0, 8, 3, 1], np.int32); left = np.array([8, 0, 2, 5], np.int32)
    app = np.array([1, 1, 0, 1], np.uint8)
    dev = torch.device("cuda")
    d_pool = torch.from_numpy(pool).to(dev)
    d_lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
    d_out = torch.empty((B, T, 3, H, W), dtype=torch.float32, device=dev)
    chk = ctypes.c_double()
    p = lambda a: a.ctypes.data
    args = lambda top_: (p(fg), B, T, H, W, d_pool.data_ptr(), P, Hb, Wb, p(idx), p(top_), p(left), p(app),
                         d_lut.data_ptr(), cabi.f32x3(bo.DEFAULT_MEAN), cabi.f32x3(bo.DEFAULT_STD), 0.5, 0,
                         d_out.data_ptr(), ctypes.byref(chk), 0)
    cabi.check(L.bgd_bgmix_blend_f32_host(*args(top)))
    exp = bo.mix_batch(fg, pool, idx, top, left, app, crop=(H, W))
    _assert_same(d_out.cpu().numpy(), exp)
    assert abs(chk.value - float(exp.astype(np.float64).sum())) < 1e-6 * max(1.0, abs(chk.value))
    bad_top = top.copy(); bad_top[0] = Hb      # crop leaves the image: the reference's RandomCrop would raise
    with pytest.raises(ValueError):
        cabi.check(L.bgd_bgmix_blend_f32_host(*args(bad_top)))
    with pytest.raises(ValueError):            # crop larger than the pool image
        torch.ops.bgdebias.bgmix_blend(torch.zeros((1, 1, 40, 40, 3), dtype=torch.uint8, device=dev), d_pool,
                                       torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                                       torch.zeros(1, dtype=torch.int32, device=dev), torch.ones(1, dtype=torch.uint8, device=dev),
                                       d_lut, torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), 0.5, "NTCHW")
    with pytest.raises(ValueError):
        torch.ops.bgdebias.bgmix_blend(torch.zeros((1, 1, 8, 8, 3), dtype=torch.uint8, device=dev), d_pool,
                                       torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                                       torch.zeros(1, dtype=torch.int32, device=dev), torch.ones(1, dtype=torch.uint8, device=dev),
                                       d_lut, torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), 0.5, "NHWC")


def test_pool_lifecycle_across_tasks(env):
    """The trainer's rewrites of ``dataset.bg_files`` between CIL tasks -- keep_all_backgrounds
    (libs/cil/cil.py:193-195: ``_all_bg_files.update(bg_files); bg_files = list(_all_bg_files)``), cbf_full_bg
    (:156-160: union of two lists) and merge_bg_files (:390-393: ``bg_files.extend(other.bg_files)``, duplicates
    included) -- decode only the paths not seen before and blend exactly like a pool built from scratch."""
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(8)
    T, H, W, B = 2, 32, 32, 12
    fg = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    imgs = {f"/bg/v{i:02d}.jpg": rng.integers(0, 256, (3, 36, 48), dtype=np.uint8) for i in range(10)}
    paths = sorted(imgs)
    calls = []

    def reader(p):
        calls.append(p)
        return imgs[p]

    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        infos = [dict(frame_dir=f"/x/v{i:02d}", total_frames=T, label=i, sample=i) for i in range(B)]
        pipeline = lambda info: dict(imgs=torch.from_numpy(fg[info["sample"]]), label=torch.tensor([0]), randAug=False)
        ds = cl.BackgroundMixDataset(infos, pipeline, bg_dir=tmp, bg_resize=40, bg_crop_size=(H, W), with_randAug=True,
                                     device_mix=True, bg_reader=reader, extract_bg_if_not_found=False)

        def run(bg_files, seed):
            ds.bg_files = bg_files
            torch.manual_seed(seed)
            samples = [ds.prepare_train_frames(i) for i in range(B)]
            got = ds.gpu_collate(samples)["imgs"].cpu().numpy()
            resized = np.stack([bo.bg_resize(imgs[p], 40).numpy() for p in bg_files])
            exp = bo.mix_batch(fg, resized, [s["bg_idx"] for s in samples], [s["bg_top"] for s in samples],
                               [s["bg_left"] for s in samples], [1] * B, crop=(H, W))
            _assert_same(got, exp)
            assert all(0 <= s["bg_idx"] < len(bg_files) for s in samples)

        task0 = paths[:4]
        run(list(task0), 1)
        assert set(calls) == set(task0) and ds._store.decoded == 4
        probes = len(calls) - 4                                          # the size probe of _pool_hw, before the pool exists
        assert probes <= 1
        # task 1 with keep_all_backgrounds: set union, arbitrary order
        all_bg = set(task0)
        all_bg.update(paths[3:7])
        run(list(all_bg), 2)
        assert set(calls) == set(paths[:7]) and len(calls) == 7 + probes  # v03 was not decoded again
        # merge_bg_files: extend with an exemplar set's list (duplicates keep their draw probability)
        merged = list(all_bg)
        merged.extend(paths[5:9])
        run(merged, 3)
        assert len(calls) == 9 + probes
        # cbf_full_bg: union of the train list and the merged exemplar list; nothing new to decode
        run(list(set(task0) | set(merged)), 4)
        assert len(calls) == 9 + probes
        assert len(ds._store) == 9 and ds._store.decoded == 9


class _U8Pipeline:
    """Picklable stand-in for the mmaction pipeline up to (not including) Normalize: uint8 [T,H,W,3] clips."""
    def __init__(self, fg, ra):
        self.fg, self.ra = fg, ra

    def __call__(self, info):
        return dict(imgs=torch.from_numpy(self.fg[info["sample"]]), label=torch.tensor([info["label"]]),
                    randAug=bool(self.ra[info["sample"]]))


class _Reader:
    def __init__(self, table):
        self.table = table

    def __call__(self, path):
        return self.table[path]


def test_dataloader_workers_then_device_finish(env, tmp_path):
    """device_mix=True behind a real torch DataLoader with worker processes: workers draw and stack on the
    host (no CUDA there), the training process blends once per batch.  Output = oracle blend of the draws."""
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(17)
    T, H, W, n_bg, n = 2, 32, 32, 5, 12
    fg = rng.integers(0, 256, (n, T, H, W, 3), dtype=np.uint8)
    ra = rng.integers(0, 2, n).astype(bool)
    names = [f"v{i:02d}" for i in range(n_bg)]
    for nm in names:
        (tmp_path / (nm + ".jpg")).write_bytes(b"stub")
    table = {str((tmp_path / (nm + ".jpg")).resolve()): rng.integers(0, 256, (3, 36, 48), dtype=np.uint8) for nm in names}
    infos = [dict(frame_dir=f"/x/{names[i % n_bg]}", total_frames=T, label=i, sample=i) for i in range(n)]
    ds = cl.BackgroundMixDataset(infos, _U8Pipeline(fg, ra), bg_dir=str(tmp_path), bg_resize=40, bg_crop_size=(H, W),
                                 with_randAug=True, device_mix=True, bg_reader=_Reader(table))
    ds._pool_hw()                                             # size probe once, before the workers fork
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, num_workers=2, collate_fn=ds.host_collate,
                                         pin_memory=True)
    seen = 0
    for batch in loader:
        assert batch["imgs"].dtype == torch.uint8 and not batch["imgs"].is_cuda
        out = ds.device_finish(batch)
        B = batch["imgs"].shape[0]
        idx = batch["bg_idx"].tolist()
        resized = np.stack([bo.bg_resize(table[p], 40).numpy() for p in ds.bg_files])
        exp = bo.mix_batch(fg[seen:seen + B], resized, [max(i, 0) for i in idx], batch["bg_top"].tolist(),
                           batch["bg_left"].tolist(), batch["bg_apply"].tolist(), crop=(H, W))
        _assert_same(out["imgs"].cpu().numpy(), exp)
        assert out["imgs"].is_cuda and [(i == -1) for i in idx] == list(ra[seen:seen + B])
        assert set(out) == {"imgs", "bg_idx", "label", "randAug"}
        seen += B
    assert seen == n
