"""Parity of the foreground pipeline tail on the GPU (torch.ops.bgdebias.resize_bilinear / bgmix_resize_blend) with
cv2.resize(..., INTER_LINEAR) -- what the reference pipeline's `Resize(scale=(224, 224), keep_ratio=False)`
(configs/ucf101/bgmix_plus_randAug/..._bgmix_plus_randAug.py:136) evaluates through mmaction/mmcv -- as recorded in
tests/golden/resize_reference.npz and as restated by oracle/resize_oracle.py.  Integer work: bit-exact.  The fused
Resize -> Normalize -> FormatShape -> blend must equal the oracle blend of the oracle-resized clip bit for bit
(tolerance of the path: 1e-6 relative, asserted as well)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN                               # noqa: E402
from oracle import bgmix_oracle as bo                     # noqa: E402
from oracle import resize_oracle as ro                    # noqa: E402

_NPZ = np.load(GOLDEN / "resize_reference.npz")
_CASES = sorted({k.split("/")[0] for k in _NPZ.files if k.startswith("case")}, key=lambda s: int(s[4:]))
RTOL = ATOL = 1e-6


@pytest.fixture(scope="module")
def env():
    import bgdebias_b200.ops as ops
    from bgdebias_b200 import _cabi, comix_loader, pool
    assert torch.cuda.is_available()
    _cabi.lib()
    return ops, _cabi, comix_loader, pool


def _resize(ops, clips, H, W):
    buf, geom = ops.pack_clips(clips)
    return torch.ops.bgdebias.resize_bilinear(buf.cuda(), geom, int(clips[0].shape[0]), H, W).cpu().numpy()


@pytest.mark.parametrize("case", _CASES)
def test_recorded_cv2_outputs(env, case):
    ops = env[0]
    src, dst = _NPZ[case + "/src"], _NPZ[case + "/dst"]
    got = _resize(ops, [src[None]], dst.shape[0], dst.shape[1])
    np.testing.assert_array_equal(got[0, 0], dst)


def test_multiscale_crop_batch(env):
    """One ragged batch holding every crop shape MultiScaleCrop (config :129-135) can hand to Resize, T = 8."""
    ops = env[0]
    rng = np.random.default_rng(3)
    clips = [rng.integers(0, 256, (8, ch, cw, 3), dtype=np.uint8) for cw, ch in ro.multiscale_crop_sizes()]
    got = _resize(ops, clips, 224, 224)
    for b, clip in enumerate(clips):
        np.testing.assert_array_equal(got[b], ro.resize_clip(clip, 224, 224), err_msg=str(clip.shape))


def test_random_shapes_against_cv2(env):
    cv2 = pytest.importorskip("cv2")
    ops = env[0]
    rng = np.random.default_rng(4)
    for _ in range(40):
        dh, dw = (int(v) for v in rng.integers(1, 120, 2))
        clips = [rng.integers(0, 256, (2, int(rng.integers(1, 150)), int(rng.integers(1, 150)), 3), dtype=np.uint8)
                 for _ in range(5)]
        got = _resize(ops, clips, dh, dw)
        for b, clip in enumerate(clips):
            for t in range(2):
                exp = cv2.resize(clip[t], (dw, dh), interpolation=cv2.INTER_LINEAR)
                np.testing.assert_array_equal(got[b, t], exp, err_msg=f"{clip.shape} -> {(dh, dw)}")


def test_integer_ratios_and_identity(env):
    ops = env[0]
    rng = np.random.default_rng(5)
    for sh, sw, dh, dw in [(448, 448, 224, 224), (672, 448, 224, 224), (224, 224, 224, 224), (112, 56, 224, 224), (1, 1, 5, 5)]:
        clip = rng.integers(0, 256, (1, sh, sw, 3), dtype=np.uint8)
        np.testing.assert_array_equal(_resize(ops, [clip], dh, dw)[0], ro.resize_clip(clip, dh, dw))


def test_crops_addressed_inside_full_frames(env):
    """The geometry table can point a crop into a larger decoded frame (row stride > 3w, odd byte offsets)."""
    ops = env[0]
    rng = np.random.default_rng(6)
    T, FH, FW = 3, 64, 85
    frames = rng.integers(0, 256, (2, T, FH, FW, 3), dtype=np.uint8)
    boxes = [(5, 7, 42, 56), (0, 29, 64, 56)]                       # top, left, h, w
    flat = torch.from_numpy(frames.reshape(-1))
    flat = torch.cat([flat, flat.new_zeros(4 + (-flat.numel()) % 4)])
    geom = torch.tensor([[((b * T * FH + top) * FW + left) * 3, h, w, FW * 3, FH * FW * 3]
                         for b, (top, left, h, w) in enumerate(boxes)], dtype=torch.int64)
    got = torch.ops.bgdebias.resize_bilinear(flat.cuda(), geom, T, 56, 56).cpu().numpy()
    for b, (top, left, h, w) in enumerate(boxes):
        np.testing.assert_array_equal(got[b], ro.resize_clip(frames[b, :, top:top + h, left:left + w], 56, 56))


@pytest.mark.parametrize("layout", ["NTCHW", "NCTHW"])
@pytest.mark.parametrize("pool_dtype", [torch.float32, torch.uint8])
def test_fused_resize_blend(env, layout, pool_dtype):
    ops = env[0]
    rng = np.random.default_rng(7)
    T, H, W, P = 4, 56, 56, 6
    shapes = [(64, 64), (56, 56), (48, 56), (42, 48), (56, 64), (64, 56), (48, 48), (42, 42)]
    clips = [rng.integers(0, 256, (T, h, w, 3), dtype=np.uint8) for h, w in shapes]
    B = len(clips)
    pool = rng.integers(0, 256, (P, 3, 64, 85)).astype(np.float32)
    idx = rng.integers(0, P, B); top = rng.integers(0, 64 - H + 1, B); left = rng.integers(0, 85 - W + 1, B)
    app = (rng.random(B) < 0.6).astype(np.uint8); app[0], app[1] = 1, 0
    buf, geom = ops.pack_clips(clips)
    dev = torch.device("cuda")
    i32 = lambda a: torch.as_tensor(a, dtype=torch.int32, device=dev)     # noqa: E731
    lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
    d_pool = torch.from_numpy(pool).to(dev, pool_dtype)
    args = (d_pool, i32(idx), i32(top), i32(left), torch.as_tensor(app, device=dev), lut, torch.tensor(bo.DEFAULT_MEAN),
            torch.tensor(bo.DEFAULT_STD), 0.5, layout)
    got = torch.ops.bgdebias.bgmix_resize_blend(buf.to(dev), geom, T, H, W, *args)
    resized = np.stack([ro.resize_clip(c, H, W) for c in clips])
    exp = bo.mix_batch(resized, pool, idx, top, left, app, crop=(H, W), layout=layout)
    np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(got.cpu().numpy().view(np.uint32), exp.view(np.uint32))
    # and it is the two separate launches, exactly
    two = torch.ops.bgdebias.bgmix_blend(torch.ops.bgdebias.resize_bilinear(buf.to(dev), geom, T, H, W), *args)
    assert torch.equal(got, two)


def test_full_size_batch_properties(env):
    """Config-5 sized batch (64 clips x 8 frames -> 224x224): identity where the crop already is 224x224, the u8 resize and the
    fused launch agree everywhere, constant clips stay constant."""
    ops = env[0]
    g = torch.Generator().manual_seed(8)
    sizes = ro.multiscale_crop_sizes()
    clips = []
    for b in range(64):
        cw, ch = sizes[b % len(sizes)]
        clips.append(torch.randint(0, 256, (8, ch, cw, 3), dtype=torch.uint8, generator=g) if b % 7 else
                     torch.full((8, ch, cw, 3), 3 * b, dtype=torch.uint8))
    buf, geom = ops.pack_clips(clips)
    dev = torch.device("cuda")
    d_buf = buf.to(dev)
    out = torch.ops.bgdebias.resize_bilinear(d_buf, geom, 8, 224, 224)
    for b, c in enumerate(clips):
        if tuple(c.shape[1:3]) == (224, 224):
            assert torch.equal(out[b].cpu(), c)
        if b % 7 == 0:
            assert bool((out[b] == 3 * b).all())
    lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
    z = torch.zeros(64, dtype=torch.int32, device=dev)
    none = torch.zeros(64, dtype=torch.uint8, device=dev)
    pool = torch.zeros((1, 3, 224, 224), dtype=torch.uint8, device=dev)
    mean, std = torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD)
    fused = torch.ops.bgdebias.bgmix_resize_blend(d_buf, geom, 8, 224, 224, pool, z, z, z, none, lut, mean, std, 0.5, "NTCHW")
    exp = lut[torch.arange(3, device=dev)[None, None, :, None, None], out.permute(0, 1, 4, 2, 3).long()]
    assert torch.equal(fused, exp)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_entry_point(env, pinned):
    """bgd_bgmix_resize_blend_f32_host: packed crops and draws in host memory (pageable or pinned) -> training tensor on the
    device, equal to the device-resident op, plus the checksum; draws outside the pool are refused."""
    import ctypes
    ops, cabi = env[0], env[1]
    rng = np.random.default_rng(12)
    T, H, W, P = 3, 40, 40, 4
    clips = [rng.integers(0, 256, (T, h, w, 3), dtype=np.uint8) for h, w in [(48, 48), (40, 40), (30, 36), (44, 33), (36, 30)]]
    B = len(clips)
    buf, geom = ops.pack_clips(clips, pin=pinned)
    dev = torch.device("cuda")
    pool = torch.from_numpy(rng.integers(0, 256, (P, 3, 48, 64)).astype(np.float32)).to(dev)
    idx = rng.integers(0, P, B).astype(np.int32); top = rng.integers(0, 9, B).astype(np.int32)
    left = rng.integers(0, 25, B).astype(np.int32); app = np.array([1, 0, 1, 1, 0], np.uint8)
    lut = ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev)
    out = torch.empty((B, T, 3, H, W), dtype=torch.float32, device=dev)
    chk = ctypes.c_double()
    gptr = ctypes.cast(geom.data_ptr(), ctypes.POINTER(ctypes.c_int64))
    call = lambda idx_: cabi.lib().bgd_bgmix_resize_blend_f32_host(                     # noqa: E731
        buf.data_ptr(), buf.numel(), gptr, B, T, H, W, pool.data_ptr(), P, 48, 64, idx_.ctypes.data, top.ctypes.data,
        left.ctypes.data, app.ctypes.data, lut.data_ptr(), cabi.f32x3(bo.DEFAULT_MEAN), cabi.f32x3(bo.DEFAULT_STD), 0.5, 0,
        out.data_ptr(), ctypes.byref(chk), 0)
    cabi.check(call(idx))
    t = lambda a: torch.from_numpy(a).to(dev)                                           # noqa: E731
    exp = torch.ops.bgdebias.bgmix_resize_blend(buf.to(dev), geom, T, H, W, pool, t(idx), t(top), t(left), t(app), lut,
                                                torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), 0.5, "NTCHW")
    assert torch.equal(out, exp)
    assert abs(chk.value - float(exp.double().sum())) <= 1e-6 * max(1.0, abs(chk.value))
    bad = idx.copy(); bad[0] = P
    with pytest.raises(ValueError):
        cabi.check(call(bad))


def test_errors(env):
    ops = env[0]
    clip = torch.zeros((1, 4, 4, 3), dtype=torch.uint8)
    buf, geom = ops.pack_clips([clip])
    with pytest.raises(ValueError):
        torch.ops.bgdebias.resize_bilinear(buf[:49].cuda(), geom, 1, 8, 8)            # not a multiple of 4 bytes
    bad = geom.clone(); bad[0, 1] = 40                                              # crop runs past the buffer
    with pytest.raises(ValueError):
        torch.ops.bgdebias.resize_bilinear(buf.cuda(), bad, 1, 8, 8)
    bad = geom.clone(); bad[0, 3] = 5                                               # row stride below 3*w
    with pytest.raises(ValueError):
        torch.ops.bgdebias.resize_bilinear(buf.cuda(), bad, 1, 8, 8)
    with pytest.raises(NotImplementedError):
        torch.ops.bgdebias.resize_bilinear(buf, geom, 1, 8, 8)                        # CPU tensors: no fallback
    assert torch.ops.bgdebias.resize_bilinear(buf.cuda(), geom[:0], 1, 8, 8).shape == (0, 1, 8, 8, 3)


class _CropPipeline:
    """Stand-in for the mmaction pipeline up to MultiScaleCrop: uint8 clips whose size differs per sample."""
    def __init__(self, clips, ra):
        self.clips, self.ra = clips, ra

    def __call__(self, info):
        return dict(imgs=torch.from_numpy(self.clips[info["sample"]]), label=torch.tensor([info["label"]]),
                    randAug=bool(self.ra[info["sample"]]))


class _Reader:
    def __init__(self, table):
        self.table = table

    def __call__(self, path):
        return self.table[path]


def test_dataset_ships_unresized_crops(env, tmp_path):
    """device_mix=True with a pipeline that stops before Resize: host_collate packs the ragged crops, device_finish returns
    what Resize -> Normalize -> FormatShape -> _mix_background (comix_loader.py:138-145) gives."""
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(9)
    T, H, W, n_bg, n = 2, 32, 32, 4, 8
    shapes = [(36, 36), (32, 32), (28, 32), (24, 28), (32, 36), (36, 32), (28, 28), (24, 24)]
    clips = [rng.integers(0, 256, (T, h, w, 3), dtype=np.uint8) for h, w in shapes]
    ra = rng.integers(0, 2, n).astype(bool)
    names = [f"v{i:02d}" for i in range(n_bg)]
    for nm in names:
        (tmp_path / (nm + ".jpg")).write_bytes(b"stub")
    table = {str((tmp_path / (nm + ".jpg")).resolve()): rng.integers(0, 256, (3, 36, 48), dtype=np.uint8) for nm in names}
    infos = [dict(frame_dir=f"/x/{names[i % n_bg]}", total_frames=T, label=i, sample=i) for i in range(n)]
    ds = cl.BackgroundMixDataset(infos, _CropPipeline(clips, ra), bg_dir=str(tmp_path), bg_resize=40, bg_crop_size=(H, W),
                                 with_randAug=True, device_mix=True, bg_reader=_Reader(table))
    ds._pool_hw()
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, num_workers=2, collate_fn=ds.host_collate,
                                         pin_memory=True)
    resized_pool = np.stack([bo.bg_resize(table[p], 40).numpy() for p in ds.bg_files])
    seen = 0
    for batch in loader:
        assert batch["imgs"].dim() == 1 and batch["fg_geom"].shape == (4, 5) and not batch["imgs"].is_cuda
        out = ds.device_finish(batch)
        idx = batch["bg_idx"].tolist()
        fg = np.stack([ro.resize_clip(c, H, W) for c in clips[seen:seen + 4]])
        exp = bo.mix_batch(fg, resized_pool, [max(i, 0) for i in idx], batch["bg_top"].tolist(), batch["bg_left"].tolist(),
                           batch["bg_apply"].tolist(), crop=(H, W))
        np.testing.assert_array_equal(out["imgs"].cpu().numpy().view(np.uint32), exp.view(np.uint32))
        assert set(out) == {"imgs", "bg_idx", "label", "randAug"}
        seen += 4
    # a stacked batch of one other size takes the same path
    same = [dict(imgs=torch.from_numpy(clips[0]), bg_idx=-1, bg_top=0, bg_left=0, bg_apply=0) for _ in range(2)]
    out = ds.device_finish(ds.host_collate(same))
    exp = bo.mix_batch(np.stack([ro.resize_clip(clips[0], H, W)] * 2), resized_pool, [0, 0], [0, 0], [0, 0], [0, 0], crop=(H, W))
    np.testing.assert_array_equal(out["imgs"].cpu().numpy().view(np.uint32), exp.view(np.uint32))


def test_unresized_crops_with_a_mixed_size_pool(env, tmp_path):
    """The same pipeline (crops of mixed sizes, Resize on the device) over a pool of mixed image sizes: the foreground
    Resize runs as its own launch, then the ragged-pool blend -- equal to Resize -> Normalize -> _mix_background with each
    background resized and cropped at its own size (comix_loader.py:72-75,138-145)."""
    from oracle import aa_resize_oracle as ao
    ops, cabi, cl, pool_mod = env
    rng = np.random.default_rng(19)
    T, H, W, n = 2, 32, 32, 8
    shapes = [(36, 36), (32, 32), (28, 32), (24, 28), (32, 36), (36, 32), (28, 28), (24, 24)]
    clips = [rng.integers(0, 256, (T, h, w, 3), dtype=np.uint8) for h, w in shapes]
    ra = np.zeros(n, bool)
    names = [f"v{i:02d}" for i in range(4)]
    for nm in names:
        (tmp_path / (nm + ".jpg")).write_bytes(b"stub")
    bg_sizes = [(36, 48), (36, 64), (50, 40), (80, 100)]
    table = {str((tmp_path / (nm + ".jpg")).resolve()): rng.integers(0, 256, (3,) + bg_sizes[i], dtype=np.uint8) for i, nm in enumerate(names)}
    infos = [dict(frame_dir=f"/x/{names[i % 4]}", total_frames=T, label=i, sample=i) for i in range(n)]
    ds = cl.BackgroundMixDataset(infos, _CropPipeline(clips, ra), bg_dir=str(tmp_path), bg_resize=40, bg_crop_size=(H, W),
                                 with_randAug=True, device_mix=True, bg_reader=_Reader(table))
    torch.manual_seed(5)
    samples = [ds.prepare_train_frames(i) for i in range(n)]
    batch = ds.host_collate(samples)
    assert batch["imgs"].dim() == 1 and batch["fg_geom"].shape == (n, 5)
    out = ds.device_finish(batch)["imgs"].cpu().numpy()
    assert ds.device_pool().tensor is None
    exp = np.stack([bo.mix_clip(ro.resize_clip(clips[i], H, W), ao.aa_resize(table[ds.bg_files[s["bg_idx"]]], 40), s["bg_top"], s["bg_left"],
                                (H, W), 0.5, True) for i, s in enumerate(samples)])
    np.testing.assert_array_equal(out.view(np.uint32), exp.view(np.uint32))
