"""ActorCutMix blend: the oracle against the reference's own outputs
(tests/golden/acm_reference.npz, libs/loader/actor_cut_mix_loader.py:135-163) on the CPU, the CUDA op and the
drop-in function against both on the GPU.  Bit-exact (uint8); the foreground ratio is compared as a float64."""
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from conftest import GOLDEN, median_case_names            # noqa: E402
from oracle import acm_oracle as ao                       # noqa: E402

_NPZ = np.load(GOLDEN / "acm_reference.npz")
CASES = median_case_names(_NPZ)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    a, m, s = _NPZ[name + "/actor"], _NPZ[name + "/mask"], _NPZ[name + "/scene"]
    np.testing.assert_array_equal(ao.cut_mix(a, m, s), _NPZ[name + "/expected"])
    assert ao.foreground_ratio(m) == float(_NPZ[name + "/foreground_ratio"])


def test_oracle_is_numpy_uint8_arithmetic():
    rng = np.random.default_rng(0)
    a, m, s = (rng.integers(0, 256, (5, 7, 3), dtype=np.uint8) for _ in range(3))     # masks outside {0,1} wrap like numpy
    np.testing.assert_array_equal(ao.cut_mix(a, m, s), a * m + s * (1 - m))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_op_and_dropin(name):
    import torch
    from bgdebias_b200 import actor_cut_mix as acm
    a, m, s = _NPZ[name + "/actor"], _NPZ[name + "/mask"], _NPZ[name + "/scene"]
    out, total = torch.ops.bgdebias.actor_cut_mix(torch.from_numpy(a).cuda(), torch.from_numpy(m).cuda(), torch.from_numpy(s).cuda())
    np.testing.assert_array_equal(out.cpu().numpy(), _NPZ[name + "/expected"])
    assert int(total) == int(m[..., 0].astype(np.uint64).sum())
    res = acm.actor_cut_mix(dict(imgs=[f.copy() for f in a], human_mask=[k.copy() for k in m]),
                            dict(imgs=[f.copy() for f in s], label=int(_NPZ[name + "/background_label"])))
    np.testing.assert_array_equal(np.stack(res["imgs"]), _NPZ[name + "/expected"])
    assert res["foreground_ratio"] == float(_NPZ[name + "/foreground_ratio"])
    assert res["background_label"] == int(_NPZ[name + "/background_label"])


@pytest.mark.gpu
def test_gpu_random_shapes_and_wrapping_masks():
    import torch
    rng = np.random.default_rng(4)
    for shape in [(1, 1, 1, 3), (2, 5, 7, 3), (8, 224, 224, 3), (3, 17, 33, 3)]:
        a, m, s = (rng.integers(0, 256, shape, dtype=np.uint8) for _ in range(3))
        out, total = torch.ops.bgdebias.actor_cut_mix(torch.from_numpy(a).cuda(), torch.from_numpy(m).cuda(), torch.from_numpy(s).cuda())
        np.testing.assert_array_equal(out.cpu().numpy(), ao.cut_mix(a, m, s))
        assert int(total) == int(m[..., 0].astype(np.uint64).sum())
    with pytest.raises(ValueError):
        torch.ops.bgdebias.actor_cut_mix(torch.zeros((2, 4, 4, 3), dtype=torch.uint8, device="cuda"),
                                         torch.zeros((2, 4, 4, 3), dtype=torch.uint8, device="cuda"),
                                         torch.zeros((2, 4, 5, 3), dtype=torch.uint8, device="cuda"))
