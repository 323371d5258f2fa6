"""Import the reference's two hot-path modules from /root/reference, unmodified, with stubs.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container);
the GPU box has no reference tree, so nothing that runs there may import this module.

Stubs (nothing else is replaced):
* ``imutils``                           -- only used by the discarded ``short_side_resize``
                                           (extract_background.py:33-39,57)
* ``mmaction.datasets.RawframeDataset`` -- minimal base whose ``prepare_train_frames`` runs a
                                           caller-supplied pipeline callable
* ``mmaction.datasets.builder.DATASETS``-- ``register_module()`` -> identity decorator
* ``mmaction.datasets.pipelines.Compose``-- placeholder; the tests replace the dataset's pipelines with callables
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BGD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cil_tools", "extract_background.py"))


def _install_stubs() -> None:
    if "imutils" not in sys.modules:
        im = types.ModuleType("imutils")
        im.resize = lambda img, width=None, height=None: img
        sys.modules["imutils"] = im
    if "mmaction" not in sys.modules:
        mm = types.ModuleType("mmaction")
        ds = types.ModuleType("mmaction.datasets")
        bd = types.ModuleType("mmaction.datasets.builder")

        class _Registry:
            def register_module(self, *a, **k):
                return lambda cls: cls

        class RawframeDataset:  # the slice of mmaction's class the reference touches
            def __init__(self, ann_file, pipeline, data_prefix=None, test_mode=False,
                         filename_tmpl='img_{:05}.jpg', with_offset=False, multi_class=False,
                         num_classes=None, start_index=1, modality='RGB', sample_by_class=False,
                         power=0., dynamic_length=False, **kwargs):
                self.video_infos = list(ann_file) if isinstance(ann_file, (list, tuple)) else []
                self.pipeline = pipeline
                self.data_prefix = data_prefix
                self.test_mode = test_mode
                self.filename_tmpl = filename_tmpl
                self.start_index = start_index

            def prepare_train_frames(self, idx):
                import copy
                return self.pipeline(copy.deepcopy(self.video_infos[idx]))

        bd.DATASETS = _Registry()
        bd.PIPELINES = _Registry()
        ds.RawframeDataset = RawframeDataset
        ds.builder = bd
        ds.PIPELINES = bd.PIPELINES
        mm.datasets = ds
        pl = types.ModuleType("mmaction.datasets.pipelines")
        pl.Compose = lambda cfgs: (lambda results: results)      # never called by the tests: pipelines are replaced
        ds.pipelines = pl
        sys.modules.update({"mmaction": mm, "mmaction.datasets": ds, "mmaction.datasets.builder": bd,
                            "mmaction.datasets.pipelines": pl})


def _load(name: str, rel: str):
    _install_stubs()
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_extract_background():
    """The module at cil_tools/extract_background.py (``bg_extraction_tmf`` :42-75)."""
    return _load("_ref_extract_background", os.path.join("cil_tools", "extract_background.py"))


def load_actor_cut_mix_loader():
    """The module at libs/loader/actor_cut_mix_loader.py (``ActorCutMixDataset.actor_cut_mix`` :135-152)."""
    return _load("_ref_actor_cut_mix_loader", os.path.join("libs", "loader", "actor_cut_mix_loader.py"))


def load_comix_loader():
    """The module at libs/loader/comix_loader.py (``BackgroundMixDataset`` :16-145)."""
    return _load("_ref_comix_loader", os.path.join("libs", "loader", "comix_loader.py"))


def reference_median_of_frames():
    """``frames -> median frame`` through the REFERENCE's own ``bg_extraction_tmf`` (cil_tools/extract_background.py:42-75)
    with its file I/O stubbed out: ``cv2.VideoCapture`` is replaced by an in-memory capture that hands back the given
    frames and ``cv2.imwrite`` by a no-op, so what runs (and what bench.py's CPU arm times in the build container) is the
    reference's loop, its ``np.median(frames, axis=0).astype(np.uint8)`` and nothing of the codec."""
    import types
    ref = load_extract_background()

    class _Capture:
        def __init__(self, frames):
            self.frames, self.i = frames, 0

        def isOpened(self):
            return True

        def read(self):
            if self.i >= len(self.frames):
                return False, None
            f = self.frames[self.i]
            self.i += 1
            return True, f

    holder = {}
    fake_cv2 = types.SimpleNamespace(**{k: getattr(ref.cv2, k) for k in dir(ref.cv2) if not k.startswith("__")})
    fake_cv2.VideoCapture = lambda path: _Capture(holder["frames"])
    fake_cv2.imwrite = lambda path, img: True
    ref.cv2 = fake_cv2

    def run(frames):
        holder["frames"] = frames
        return ref.bg_extraction_tmf("in-memory", "nowhere.jpg", True, 1, max(len(frames), 1))
    return run
