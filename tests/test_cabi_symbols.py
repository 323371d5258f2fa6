"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bgdebias.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from bgdebias_b200 import _cabi
    return _cabi


def _declared_functions():
    text = (ROOT / "include" / "bgdebias.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bgd_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(built_lib):
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(built_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(str(built_lib.LIB_PATH))
    for name in _declared_functions():
        assert hasattr(L, name), name
    assert built_lib.lib().bgd_abi_version() == 1


def test_no_device_is_a_loud_error(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc = built_lib.lib().bgd_device_info(0, None, None, None, None, None)
    assert rc == built_lib.BGD_ERR_NO_DEVICE
    assert b"no CUDA device" in built_lib.lib().bgd_last_error()
    with pytest.raises(built_lib.BgdError):
        built_lib.check(rc)


def test_ops_refuse_cpu_tensors(built_lib):
    import torch
    import bgdebias_b200.ops  # noqa: F401
    with pytest.raises(NotImplementedError):
        torch.ops.bgdebias.temporal_median(torch.zeros((3, 16), dtype=torch.uint8))


def test_product_never_imports_the_oracle():
    pkg = ROOT / "background-debiased-video-cil_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert "oracle" not in f.read_text().lower(), f
