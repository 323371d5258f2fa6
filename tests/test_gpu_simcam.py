"""GPU parity of the NaN-masked temporal median / mean (torch.ops.bgdebias.nan_temporal_reduce*, C ABI
bgd_nan_temporal_reduce_f32) and of the sim_cam drop-in with the oracle and with the reference's own outputs
(tests/golden/simcam_reference.npz = cil_tools/extract_background.py:78-99 run on seeded PNG folders).
Bit-exact: float32 results compared as bit patterns, uint8 images and JPEG bytes compared for equality."""
import pathlib
import sys
import warnings

import cv2
import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from conftest import GOLDEN, median_case_names            # noqa: E402
from oracle import simcam_oracle as so                    # noqa: E402

pytestmark = pytest.mark.gpu
_NPZ = np.load(GOLDEN / "simcam_reference.npz")
CASES = median_case_names(_NPZ)


@pytest.fixture(scope="module")
def ops():
    import bgdebias_b200.ops as ops
    assert torch.cuda.is_available()
    return ops


def _bits(a):
    return np.nan_to_num(a, nan=0.0).view(np.uint32), np.isnan(a)


@pytest.mark.parametrize("name", CASES)
def test_kernel_on_reference_frames(ops, name):
    interval, max_frames, avg, seed = (int(v) for v in _NPZ[name + "/params"])
    tf = _NPZ[name + "/transformed"]                       # NaN already marks the zeros
    dev = torch.from_numpy(np.nan_to_num(tf, nan=0.0)).cuda()          # hand the kernel the zeros, as the drop-in does
    got = torch.ops.bgdebias.nan_temporal_reduce(dev, avg, True).cpu().numpy()
    np.testing.assert_array_equal(got, _NPZ[name + "/expected"])
    gotf = torch.ops.bgdebias.nan_temporal_reduce_f32(torch.from_numpy(tf).cuda(), avg, False).cpu().numpy()
    expf = so.nan_temporal_reduce(tf, avg)
    assert np.array_equal(_bits(gotf)[1], _bits(expf)[1])
    np.testing.assert_array_equal(_bits(gotf)[0], _bits(expf)[0])


@pytest.mark.parametrize("T", [1, 2, 3, 4, 7, 8, 31, 64, 65, 180, 333])
@pytest.mark.parametrize("avg", [0, 1])
def test_kernel_random(ops, T, avg):
    rng = np.random.default_rng(7 * T + avg)
    N = 1000 + 37
    x = (rng.random((T, N), dtype=np.float32) * 255).astype(np.float32)
    x[:, :50] = np.round(x[:, :50])                        # integral values: many ties
    x[:, 50:60] -= 128                                     # negative values order correctly
    x[rng.random(x.shape) < 0.25] = np.nan
    x[rng.random(x.shape) < 0.05] = 0.0                    # counted or not, by zero_is_missing
    x[:, 60] = np.nan
    x[1:, 61] = np.nan
    for zero_missing in (False, True):
        ref_in = x.copy()
        if zero_missing:
            ref_in[ref_in == 0] = np.nan
        expf = so.nan_temporal_reduce(ref_in, avg)
        gotf = torch.ops.bgdebias.nan_temporal_reduce_f32(torch.from_numpy(x).cuda(), avg, zero_missing).cpu().numpy()
        assert np.array_equal(_bits(gotf)[1], _bits(expf)[1])
        np.testing.assert_array_equal(_bits(gotf)[0], _bits(expf)[0])
        pos = ~(expf < 0)                                  # the uint8 cast is defined on 0 <= v < 256 (and NaN)
        got8 = torch.ops.bgdebias.nan_temporal_reduce(torch.from_numpy(x).cuda(), avg, zero_missing).cpu().numpy()
        np.testing.assert_array_equal(got8[pos], so.cast_u8(expf)[pos])


def test_long_columns_and_errors(ops):
    rng = np.random.default_rng(1)
    x = (rng.random((1200, 130), dtype=np.float32) * 255).astype(np.float32)    # beyond the shared-memory tile: scratch path
    x[rng.random(x.shape) < 0.2] = np.nan
    gotf = torch.ops.bgdebias.nan_temporal_reduce_f32(torch.from_numpy(x).cuda(), 0, False).cpu().numpy()
    np.testing.assert_array_equal(_bits(gotf)[0], _bits(so.nan_temporal_reduce(x, 0))[0])
    with pytest.raises(ValueError):
        torch.ops.bgdebias.nan_temporal_reduce(torch.zeros((0, 4), device="cuda"), 0, True)
    with pytest.raises(ValueError):
        torch.ops.bgdebias.nan_temporal_reduce(torch.zeros((2, 4), device="cuda"), 2, True)
    with pytest.raises(ValueError):
        torch.ops.bgdebias.nan_temporal_reduce(torch.zeros((2, 4), device="cuda", dtype=torch.float64), 0, True)


@pytest.mark.parametrize("name", CASES)
def test_dropin_writes_the_reference_jpeg(ops, name, tmp_path):
    """sim_cam_motion_bg_extract(data_path, dest, from_video, interval, max_frames, avg_method): same signature,
    returns None, writes the JPEG the reference wrote for the same folder and torch seed."""
    from bgdebias_b200 import extract_background as eb
    interval, max_frames, avg, seed = (int(v) for v in _NPZ[name + "/params"])
    d = tmp_path / "video"
    d.mkdir()
    for i, f in enumerate(_NPZ[name + "/frames"]):
        assert cv2.imwrite(str(d / f"img_{i + 1:05d}.png"), cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    torch.manual_seed(seed)
    assert eb.sim_cam_motion_bg_extract(d, tmp_path / "o.jpg", False, interval, max_frames, avg) is None
    assert (tmp_path / "o.jpg").read_bytes() == _NPZ[name + "/jpeg"].tobytes()
    torch.manual_seed(seed)
    np.testing.assert_array_equal(eb.sim_cam_background(d, interval, max_frames, avg), _NPZ[name + "/expected"])


def test_varlen_batch(ops):
    rng = np.random.default_rng(21)
    Ts = [1, 9, 40, 2, 77]
    offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    x = (rng.random((int(offs[-1]), 6, 7, 3), dtype=np.float32) * 255).astype(np.float32)
    x[rng.random(x.shape) < 0.3] = 0.0
    for avg in (0, 1):
        got = torch.ops.bgdebias.nan_temporal_reduce_varlen(torch.from_numpy(x).cuda(), torch.from_numpy(offs), avg, True).cpu().numpy()
        for v in range(len(Ts)):
            ref_in = x[offs[v]:offs[v + 1]].copy()
            ref_in[ref_in == 0] = np.nan
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                np.testing.assert_array_equal(got[v], so.cast_u8(so.nan_temporal_reduce(ref_in, avg)), err_msg=f"video {v} avg {avg}")
    with pytest.raises(ValueError):
        torch.ops.bgdebias.nan_temporal_reduce_varlen(torch.from_numpy(x).cuda(), torch.tensor([0, 3, 3]), 0, True)
