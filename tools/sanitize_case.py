"""Small end-to-end case for compute-sanitizer: every kernel family of the AUTO path on a few shapes, checked
against the oracle.  usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgdebias_b200.ops as ops
from oracle import c_oracle, bgmix_oracle as bo

dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
N = 2048 + 48                                            # a ragged last tile
Ts = [5, 16, 37, 64, 100, 181, 240, 300, 501, 530]       # ldsm two-column / shared half group / one-column, colplane
offs = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
fr = rng.integers(0, 256, (int(offs[-1]), N), dtype=np.uint8)
out = torch.ops.bgdebias.temporal_median_varlen(torch.from_numpy(fr).to(dev), torch.from_numpy(offs)).cpu().numpy()
for v, T in enumerate(Ts):
    assert np.array_equal(out[v], c_oracle.temporal_median(fr[offs[v]:offs[v + 1]])), f"median mismatch at T={T}"
B, T, H, W = 3, 4, 32, 32
fg = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8); pool = rng.uniform(0, 255, (5, 3, 40, 53)).astype(np.float32)
idx, top, left, app = [4, 0, 2], [3, 0, 8], [1, 21, 0], [1, 0, 1]
t = lambda a, dt: torch.tensor(a, dtype=dt, device=dev)
got = torch.ops.bgdebias.bgmix_blend(torch.from_numpy(fg).to(dev), torch.from_numpy(pool).to(dev), t(idx, torch.int32), t(top, torch.int32),
                                     t(left, torch.int32), t(app, torch.uint8), ops.make_fg_lut(bo.DEFAULT_MEAN, bo.DEFAULT_STD, dev),
                                     torch.tensor(bo.DEFAULT_MEAN), torch.tensor(bo.DEFAULT_STD), 0.5, "NTCHW").cpu().numpy()
np.testing.assert_allclose(got, bo.mix_batch(fg, pool, idx, top, left, app, crop=(H, W), alpha=0.5), rtol=1e-6, atol=1e-6)
torch.cuda.synchronize()
print("sanitize_case ok")
