import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isfile("/root/reference/cil_tools/extract_background.py")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_median():
    import numpy as np
    return np.load(GOLDEN / "median_reference.npz")


@pytest.fixture(scope="session")
def golden_bgmix():
    import numpy as np
    return np.load(GOLDEN / "bgmix_reference.npz")


def median_case_names(npz):
    return sorted({k.split("/")[0] for k in npz.files})
