"""bgdebias_b200 -- B200-native background-debiasing data path.

Drop-in for two pieces of NinV/Background-Debiased-Video-CIL:

* ``cil_tools/extract_background.py`` -> :mod:`bgdebias_b200.extract_background`
  (same CLI flags, same ``bg_extraction_tmf`` signature; the median runs on the GPU)
* ``libs/loader/comix_loader.py``     -> :mod:`bgdebias_b200.comix_loader`
  (``BackgroundMixDataset`` with the same kwargs/attributes, a ``BackgroundMix`` pipeline
  transform, and a batch-level GPU blend)

Compute goes through ``libbgdebias_b200.so`` (C ABI in ``include/bgdebias.h``), loaded with
ctypes by :mod:`bgdebias_b200._cabi` and exposed as ``torch.ops.bgdebias.*`` by
:mod:`bgdebias_b200.ops`.  There is no CPU fallback.
"""
__version__ = "0.1.0"

_SUBMODULES = ("_cabi", "ops", "extract_background", "comix_loader", "staging", "shard", "pool", "build")


def __getattr__(name):
    import importlib
    if name in _SUBMODULES:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
