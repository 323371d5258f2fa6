"""Parity of the CUDA temporal median (through torch.ops.bgdebias and the C ABI) with the oracle.

Bit-exact (uint8).  Covers: the reference's own outputs (tests/golden/median_reference.npz),
every kernel variant, odd/even T, T from 1 to 576, ragged column counts, varlen batches,
host-buffer entry points, and full-size shapes from BASELINE.json's configs (checked against the
C oracle and through order-invariance / idempotence properties).
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN, median_case_names            # noqa: E402
from oracle import c_oracle, median_oracle as mo         # noqa: E402

_NPZ = np.load(GOLDEN / "median_reference.npz")
CASES = median_case_names(_NPZ)


@pytest.fixture(scope="module")
def bgd():
    import bgdebias_b200.ops as ops
    from bgdebias_b200 import _cabi
    assert torch.cuda.is_available()
    _cabi.lib()
    yield ops, _cabi
    _cabi.set_median_variant(_cabi.MEDIAN_AUTO)


def _gpu_median(ops, frames_np):
    t = torch.from_numpy(np.ascontiguousarray(frames_np)).cuda()
    return torch.ops.bgdebias.temporal_median(t).cpu().numpy()


VARIANTS = {"auto": 0, "swar": 1, "bitsliced": 2, "colplane": 3, "ldsm": 4}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", CASES)
def test_golden_reference_outputs(bgd, name, variant):
    ops, cabi = bgd
    frames = _NPZ[name + "/frames"]
    interval, max_frames = (int(v) for v in _NPZ[name + "/params"])
    used = frames[mo.select_frame_indices(len(frames), interval, max_frames)]
    N = int(np.prod(used.shape[1:]))
    cabi.set_median_variant(VARIANTS[variant])
    if variant in ("bitsliced", "colplane", "ldsm") and N % 16 != 0:
        with pytest.raises(cabi.BgdError):
            _gpu_median(ops, used)
        return
    got = _gpu_median(ops, used)
    np.testing.assert_array_equal(got, _NPZ[name + "/expected"])


T_VALUES = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 96, 100, 127, 128, 129,
            179, 180, 181, 239, 240, 255, 256, 257, 300, 383, 384, 385, 500, 501, 511, 512, 513, 527, 528, 543, 544, 576]


@pytest.mark.parametrize("variant", ["swar", "bitsliced", "colplane", "ldsm"])
@pytest.mark.parametrize("T", T_VALUES)
def test_random_all_T(bgd, T, variant):
    ops, cabi = bgd
    if (variant == "bitsliced" and T > 528) or (variant in ("colplane", "ldsm") and T > 544):
        pytest.skip("the TMA variants hold at most ~500 rows per thread group; AUTO falls back to the generic variant")
    cabi.set_median_variant(VARIANTS[variant])
    rng = np.random.default_rng(1000 + T)
    N = 16 * int(rng.integers(1, 90))                     # ragged tile tails
    fr = rng.integers(0, 256, (T, N), dtype=np.uint8)
    np.testing.assert_array_equal(_gpu_median(ops, fr), c_oracle.temporal_median(fr))


@pytest.mark.parametrize("pattern", ["constant", "two_valued", "saturated", "sorted", "reverse_sorted",
                                     "near_ties", "narrow_high", "all255", "all0", "static_scene"])
@pytest.mark.parametrize("T", [6, 37, 96, 180, 300])
def test_patterns(bgd, pattern, T):
    ops, cabi = bgd
    rng = np.random.default_rng(hash((pattern, T)) % (2 ** 32))
    N = 2048 + 48
    if pattern == "constant":
        fr = np.full((T, N), 77, np.uint8)
    elif pattern == "two_valued":
        fr = rng.choice(np.array([3, 250], np.uint8), (T, N))
    elif pattern == "saturated":
        fr = rng.choice(np.array([0, 255], np.uint8), (T, N))
    elif pattern == "sorted":
        fr = np.sort(rng.integers(0, 256, (T, N), dtype=np.uint8), axis=0)
    elif pattern == "reverse_sorted":
        fr = np.sort(rng.integers(0, 256, (T, N), dtype=np.uint8), axis=0)[::-1].copy()
    elif pattern == "near_ties":
        fr = (rng.integers(2, 254, (1, N)) + rng.integers(-2, 3, (T, N))).astype(np.uint8)
    elif pattern == "static_scene":
        # a fixed background, a few noisy columns and a block passing through: most warps never part the two middles of an
        # even T, some part them late in a few lanes only
        fr = np.repeat(rng.integers(0, 256, (1, N), dtype=np.uint8), T, axis=0)
        noisy = rng.random(N) < 0.02
        fr[:, noisy] = (fr[:, noisy].astype(np.int16) + rng.integers(-1, 2, (T, int(noisy.sum())))).clip(0, 255).astype(np.uint8)
        for t in range(T):
            fr[t, (37 * t) % (N - 64):(37 * t) % (N - 64) + 64] = 200 + t % 7
    elif pattern == "narrow_high":
        fr = rng.integers(250, 256, (T, N), dtype=np.uint8)
    elif pattern == "all255":
        fr = np.full((T, N), 255, np.uint8)
    else:
        fr = np.zeros((T, N), np.uint8)
    exp = mo.temporal_median_np(fr)
    for variant in (1, 2, 3, 4):
        cabi.set_median_variant(variant)
        np.testing.assert_array_equal(_gpu_median(ops, fr), exp)


def test_4d_input_and_unaligned_columns(bgd):
    ops, cabi = bgd
    cabi.set_median_variant(0)
    rng = np.random.default_rng(7)
    for shape in [(9, 10, 14, 3), (12, 5, 7, 3), (8, 3), (5, 1), (4, 240, 427, 3)]:
        fr = rng.integers(0, 256, shape, dtype=np.uint8)
        got = _gpu_median(ops, fr)
        assert got.shape == shape[1:]
        np.testing.assert_array_equal(got, mo.temporal_median_np(fr))


def test_varlen_mixed_lengths(bgd):
    ops, cabi = bgd
    rng = np.random.default_rng(11)
    Ts = [1, 2, 180, 64, 65, 33, 240, 96, 7, 128, 501, 30, 31, 200]
    offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    N = 4096 + 16 * 5
    fr = rng.integers(0, 256, (int(offs[-1]), N), dtype=np.uint8)
    exp = c_oracle.temporal_median_varlen(fr, offs)
    d = torch.from_numpy(fr).cuda()
    for variant in (0, 1, 2, 3, 4):
        cabi.set_median_variant(variant)
        got = torch.ops.bgdebias.temporal_median_varlen(d, torch.from_numpy(offs)).cpu().numpy()
        np.testing.assert_array_equal(got, exp)
    got32 = torch.ops.bgdebias.temporal_median_varlen(d, torch.from_numpy(offs.astype(np.int32))).cpu().numpy()
    np.testing.assert_array_equal(got32, exp)


def test_errors(bgd):
    ops, cabi = bgd
    cabi.set_median_variant(0)
    with pytest.raises(ValueError):
        torch.ops.bgdebias.temporal_median(torch.zeros((0, 16), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):      # a video without frames inside a batch
        torch.ops.bgdebias.temporal_median_varlen(torch.zeros((4, 16), dtype=torch.uint8, device="cuda"),
                                                  torch.tensor([0, 2, 2, 4]))
    with pytest.raises(ValueError):
        torch.ops.bgdebias.temporal_median(torch.zeros((3, 16), dtype=torch.float32, device="cuda"))
    with pytest.raises(NotImplementedError):   # no CPU implementation exists
        torch.ops.bgdebias.temporal_median(torch.zeros((3, 16), dtype=torch.uint8))
    empty_cols = torch.ops.bgdebias.temporal_median(torch.zeros((3, 0), dtype=torch.uint8, device="cuda"))
    assert empty_cols.shape == (0,)


def test_config1_shape_vs_c_oracle(bgd):
    """BASELINE config 1: UCF101-shaped 64x240x320x3, random and structured."""
    ops, cabi = bgd
    cabi.set_median_variant(0)
    rng = np.random.default_rng(0)
    fr = rng.integers(0, 256, (64, 240, 320, 3), dtype=np.uint8)
    np.testing.assert_array_equal(_gpu_median(ops, fr), c_oracle.temporal_median(fr))
    bg = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    st = np.broadcast_to(bg, (65, 240, 320, 3)).copy()
    for t in range(65):
        st[t, 100:140, (t * 4) % 280:(t * 4) % 280 + 40] = 255
    got = _gpu_median(ops, st)
    np.testing.assert_array_equal(got, c_oracle.temporal_median(st))
    np.testing.assert_array_equal(got, bg)              # the moving block never covers a pixel half the time


@pytest.mark.parametrize("T,H,W", [(180, 240, 320), (96, 240, 320), (48, 240, 427)])
def test_full_size_properties(bgd, T, H, W):
    """Configs 2-4 shapes: order invariance, idempotence on duplicated frames, C-oracle equality."""
    ops, cabi = bgd
    cabi.set_median_variant(0)
    g = torch.Generator(device="cuda").manual_seed(T)
    fr = torch.randint(0, 256, (T, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.ops.bgdebias.temporal_median(fr)
    perm = torch.randperm(T, device="cuda", generator=g)
    assert torch.equal(out, torch.ops.bgdebias.temporal_median(fr[perm].contiguous()))
    # duplicating every frame keeps both middle order statistics' average unchanged for odd T only;
    # for any T, median(frames ++ frames) == (s[(T-1)//2] + s[T//2]) >> 1 of the original
    assert torch.equal(out, torch.ops.bgdebias.temporal_median(torch.cat([fr, fr], 0)))
    # a constant video is its own median
    const = fr[:1].expand(T, H, W, 3).contiguous()
    assert torch.equal(torch.ops.bgdebias.temporal_median(const), fr[0])
    np.testing.assert_array_equal(out.cpu().numpy(), c_oracle.temporal_median(fr.cpu().numpy()))


def test_host_entry_points(bgd):
    ops, cabi = bgd
    cabi.set_median_variant(0)
    L = cabi.lib()
    rng = np.random.default_rng(21)
    N = 240 * 320 * 3
    frames = [rng.integers(0, 256, (240, 320, 3), dtype=np.uint8) for _ in range(33)]   # separate allocations
    ptrs = (ctypes.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    out = np.empty((240, 320, 3), np.uint8)
    cabi.check(L.bgd_temporal_median_u8_host(ptrs, len(frames), N, out.ctypes.data, 0))
    np.testing.assert_array_equal(out, mo.temporal_median_np(frames))

    Ts = [5, 64, 17, 2, 90, 33]
    offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    N2 = 16 * 1000
    block = rng.integers(0, 256, (int(offs[-1]), N2), dtype=np.uint8)
    out2 = np.empty((len(Ts), N2), np.uint8)
    import os
    os.environ["BGD_STAGING_SLAB_MB"] = "2"            # force several chunks through the double buffer
    try:
        cabi.check(L.bgd_temporal_median_varlen_u8_host(block.ctypes.data, offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                                        len(Ts), N2, out2.ctypes.data, 0))
    finally:
        del os.environ["BGD_STAGING_SLAB_MB"]
    np.testing.assert_array_equal(out2, c_oracle.temporal_median_varlen(block, offs))


def test_kernels_were_launched(bgd):
    ops, cabi = bgd
    before = cabi.kernel_launch_count()
    _gpu_median(ops, np.zeros((3, 32), np.uint8))
    assert cabi.kernel_launch_count() > before


def test_one_call_mixes_every_kernel_family(bgd):
    """One varlen call whose videos land in the two-column, shared-half-group and one-column
    instantiations of the transposing-load kernel and in the column-plane kernel (T > 512)."""
    ops, cabi = bgd
    cabi.set_median_variant(0)
    rng = np.random.default_rng(11)
    N = 2048 + 48
    Ts = [5, 16, 37, 64, 100, 181, 240, 300, 501, 530]
    offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    fr = rng.integers(0, 256, (int(offs[-1]), N), dtype=np.uint8)
    out = torch.ops.bgdebias.temporal_median_varlen(torch.from_numpy(fr).cuda(), torch.from_numpy(offs)).cpu().numpy()
    for v, T in enumerate(Ts):
        np.testing.assert_array_equal(out[v], c_oracle.temporal_median(fr[offs[v]:offs[v + 1]]), err_msg=f"T={T}")


def test_long_videos_inside_a_mixed_batch_take_the_generic_kernel_alone(bgd):
    """AUTO classifies every video, not the batch: the rawframes variant has no frame cap (comix_loader.py:157-161), so a
    call can hold videos beyond the TMA kernels' 544 frames.  Those alone take the generic kernel (one launch per run of
    consecutive long videos); the others keep the fast path."""
    ops, cabi = bgd
    cabi.set_median_variant(0)
    rng = np.random.default_rng(12)
    N = 1024 + 32
    Ts = [530, 600, 1200, 64, 181, 700, 545, 544, 2]
    offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    fr = rng.integers(0, 256, (int(offs[-1]), N), dtype=np.uint8)
    d_fr = torch.from_numpy(fr).cuda()
    n0 = cabi.kernel_launch_count()
    out = torch.ops.bgdebias.temporal_median_varlen(d_fr, torch.from_numpy(offs)).cpu().numpy()
    launches = cabi.kernel_launch_count() - n0
    for v, T in enumerate(Ts):
        np.testing.assert_array_equal(out[v], c_oracle.temporal_median(fr[offs[v]:offs[v + 1]]), err_msg=f"T={T}")
    # fast classes: 530 + 544 (column-plane), 64, 181, 2 -> at most 5 launches; generic: runs {600, 1200} and {700, 545} -> 2
    assert 3 <= launches <= 7, launches
    # a batch of long videos only still works (everything generic)
    Ts2 = [600, 900]
    offs2 = np.concatenate([[0], np.cumsum(Ts2)]).astype(np.int64)
    out2 = torch.ops.bgdebias.temporal_median_varlen(d_fr[:1500], torch.from_numpy(offs2)).cpu().numpy()
    for v in range(2):
        np.testing.assert_array_equal(out2[v], c_oracle.temporal_median(fr[offs2[v]:offs2[v + 1]]))


def test_concurrent_host_threads_and_streams(bgd):
    """Four host threads, each on its own CUDA stream, call the op concurrently (ctypes releases the GIL during
    the C-ABI call): per-thread workspaces and stream-ordered launches must not interfere."""
    import threading
    ops, cabi = bgd
    cabi.set_median_variant(0)
    rng = np.random.default_rng(31)
    N = 4096 + 16
    jobs = []
    for i in range(4):
        Ts = rng.integers(1, 300, 6)
        offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
        fr = rng.integers(0, 256, (int(offs[-1]), N), dtype=np.uint8)
        jobs.append((torch.from_numpy(fr).cuda(), torch.from_numpy(offs), fr, offs))
    torch.cuda.synchronize()
    results, errors = [None] * 4, []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                outs = [torch.ops.bgdebias.temporal_median_varlen(jobs[i][0], jobs[i][1]) for _ in range(8)]
            s.synchronize()
            results[i] = [o.cpu().numpy() for o in outs]
        except Exception as e:      # surfaced below
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, (_, _, fr, offs) in enumerate(jobs):
        exp = np.stack([c_oracle.temporal_median(fr[offs[v]:offs[v + 1]]) for v in range(len(offs) - 1)])
        for o in results[i]:
            np.testing.assert_array_equal(o, exp)


def test_calls_do_not_block_the_host(bgd):
    """Stream-ordered means the host never waits for earlier kernels: eight varlen calls are enqueued behind a kernel that
    keeps the stream busy for ~0.2 s, and the host is back before that kernel has finished (the per-thread table workspace is
    a ring, so a call only ever waits for the call 16 calls before it)."""
    import time
    dev = torch.device("cuda")
    rng = np.random.default_rng(3)
    Ts = rng.integers(20, 200, 24)
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64))
    fr = torch.randint(0, 256, (int(offs[-1]), 4096), dtype=torch.uint8, device=dev)
    ref = torch.ops.bgdebias.temporal_median_varlen(fr, offs)
    torch.cuda.synchronize()
    gate = torch.cuda.Event()
    torch.cuda._sleep(int(0.2 * 1.9e9))                  # ~0.2 s of device time in front of the calls
    gate.record()
    t0 = time.perf_counter()
    outs = [torch.ops.bgdebias.temporal_median_varlen(fr, offs) for _ in range(8)]
    host_s = time.perf_counter() - t0
    still_busy = not gate.query()
    torch.cuda.synchronize()
    assert still_busy, f"the host spent {host_s:.3f} s in 8 calls and the stream had drained: a call blocked"
    assert host_s < 0.15
    for o in outs:
        assert torch.equal(o, ref)
