"""ctypes loader for oracle/_build/libmedian_oracle.so (the plain-C oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/median_oracle.c.
"""
from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libmedian_oracle.so"
_lib = None


def build(force: bool = False) -> pathlib.Path:
    src = _HERE / "median_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        u8p, f32p, i64p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float),
                           ctypes.POINTER(ctypes.c_int64))
        L.oracle_temporal_median_u8.argtypes = [u8p, ctypes.c_int64, ctypes.c_int64, u8p]
        L.oracle_temporal_median_u8.restype = ctypes.c_int
        L.oracle_temporal_median_varlen_u8.argtypes = [u8p, i64p, ctypes.c_int64, ctypes.c_int64, u8p]
        L.oracle_temporal_median_varlen_u8.restype = ctypes.c_int
        L.oracle_bgmix_clip_f32.argtypes = [u8p, f32p, f32p, f32p, f32p, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, f32p]
        L.oracle_bgmix_clip_f32.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def temporal_median(frames: np.ndarray) -> np.ndarray:
    """frames [T, ...] uint8 -> [...] uint8 via the C counting-sort oracle."""
    fr = np.ascontiguousarray(frames, dtype=np.uint8)
    T = fr.shape[0]
    N = int(np.prod(fr.shape[1:], dtype=np.int64))
    out = np.empty(fr.shape[1:], np.uint8)
    rc = lib().oracle_temporal_median_u8(_p(fr, ctypes.c_uint8), T, N, _p(out, ctypes.c_uint8))
    if rc:
        raise ValueError("oracle_temporal_median_u8 rejected its arguments")
    return out


def temporal_median_varlen(frames: np.ndarray, offsets) -> np.ndarray:
    fr = np.ascontiguousarray(frames, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    V = len(off) - 1
    N = int(np.prod(fr.shape[1:], dtype=np.int64))
    out = np.empty((V,) + fr.shape[1:], np.uint8)
    rc = lib().oracle_temporal_median_varlen_u8(_p(fr, ctypes.c_uint8), _p(off, ctypes.c_int64), V, N,
                                                _p(out, ctypes.c_uint8))
    if rc:
        raise ValueError("oracle_temporal_median_varlen_u8 rejected its arguments")
    return out


def bgmix_clip(fg_thwc: np.ndarray, bg_crop_chw: np.ndarray, lut: np.ndarray, mean, std,
               alpha: float, apply: bool) -> np.ndarray:
    fg = np.ascontiguousarray(fg_thwc, np.uint8)
    T, H, W, _ = fg.shape
    bg = np.ascontiguousarray(bg_crop_chw, np.float32)
    lut = np.ascontiguousarray(lut, np.float32)
    m = np.asarray(mean, np.float64).astype(np.float32)
    s = np.asarray(std, np.float64).astype(np.float32)
    out = np.empty((T, 3, H, W), np.float32)
    rc = lib().oracle_bgmix_clip_f32(_p(fg, ctypes.c_uint8), _p(bg, ctypes.c_float), _p(lut, ctypes.c_float),
                                     _p(m, ctypes.c_float), _p(s, ctypes.c_float),
                                     ctypes.c_float(np.float32(1 - alpha)), ctypes.c_float(np.float32(alpha)),
                                     int(apply), T, H, W, _p(out, ctypes.c_float))
    if rc:
        raise ValueError("oracle_bgmix_clip_f32 rejected its arguments")
    return out
