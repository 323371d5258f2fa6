"""CPU oracle for the temporal-median background extraction path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and only as the checker or the
CPU arm -- never by ``bgdebias_b200``.

This file restates, in NumPy, what the reference does on the path

    cil_tools/extract_background.py:42-75   (``bg_extraction_tmf``, CLI variant, "E1")
    libs/loader/comix_loader.py:148-164     (``bg_extraction_tmf``, rawframes variant, "E2")

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the reference module
itself (with a stub ``imutils``) in the build container, runs its
``bg_extraction_tmf`` on lossless FFV1 videos and commits the input frames and
the reference's outputs under ``tests/golden/``; ``tests/test_oracle_median.py``
checks every function here against those vectors.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


# --------------------------------------------------------------------------- #
# frame selection (extract_background.py:48-60)
# --------------------------------------------------------------------------- #
def select_frame_indices(n_decodable: int, interval: int, max_frames: int) -> List[int]:
    """Indices (into the decode order) of the frames the reference keeps.

    Restates the ``while`` loop at extract_background.py:52-60:

    * the loop runs while ``len(frames) <= max_frames`` -- so up to
      ``max_frames + 1`` frames are kept (the reference's off-by-one, kept on purpose);
    * every iteration decodes one frame (``cap.read()``), also the skipped ones;
    * a frame is kept when ``count % interval == 0``;
    * a failed read on a *kept* position ends the loop (``break``), a failed read on a
      skipped position does not -- ``count`` just advances.  With ``cap.isOpened()`` still
      true that would spin until the next kept position, which then breaks; the outcome
      (the list of kept frames) is the same, so it is modelled as "stop at end of stream".
    """
    if interval <= 0:
        raise ZeroDivisionError("integer modulo by zero")  # same failure mode as `count % 0`
    kept: List[int] = []
    count = 0
    while len(kept) <= max_frames:
        ok = count < n_decodable
        if count % interval == 0:
            if ok:
                kept.append(count)
            else:
                break
        count += 1
    return kept


# --------------------------------------------------------------------------- #
# the median itself (extract_background.py:73, comix_loader.py:161)
# --------------------------------------------------------------------------- #
def temporal_median_np(frames: np.ndarray | Sequence[np.ndarray]) -> np.ndarray:
    """Exactly the reference expression: ``np.median(frames, axis=0).astype(np.uint8)``.

    ``frames`` is ``[T, ...]`` uint8 (or a list of T equal-shaped uint8 arrays).
    np.median sorts/partitions along T, takes the mean of the middle one (odd T) or two
    (even T) values in float64, and ``astype(uint8)`` truncates toward zero.
    """
    return np.median(frames, axis=0).astype(dtype=np.uint8)


def temporal_median_int(frames: np.ndarray) -> np.ndarray:
    """Integer restatement: ``(s[(T-1)//2] + s[T//2]) >> 1`` with ``s`` sorted along T.

    This is the form the CUDA kernel implements.  It is
    bit-identical to :func:`temporal_median_np` because the mean of two uint8 values is
    k or k + 0.5 exactly in float64 and the cast truncates.
    """
    frames = np.asarray(frames)
    if frames.dtype != np.uint8:
        raise TypeError("frames must be uint8")
    T = frames.shape[0]
    if T == 0:
        raise ValueError("median of zero frames")  # reference: np.median([]) -> nan -> imwrite raises
    s = np.sort(frames, axis=0)
    lo = s[(T - 1) // 2].astype(np.uint16)
    hi = s[T // 2].astype(np.uint16)
    return ((lo + hi) >> 1).astype(np.uint8)


def temporal_median_py(columns: Sequence[Sequence[int]]) -> List[int]:
    """Pure-Python loop version for tiny cases: ``columns[c]`` = the T values of column c."""
    out = []
    for col in columns:
        s = sorted(int(v) for v in col)
        T = len(s)
        out.append((s[(T - 1) // 2] + s[T // 2]) >> 1)
    return out


def temporal_median_varlen(frames: np.ndarray, offsets: Sequence[int]) -> np.ndarray:
    """Median per video for videos concatenated along T: ``frames[offsets[v]:offsets[v+1]]``.

    Mirrors what ``bg_extract_multiple`` (extract_background.py:102-109) does one video at a
    time: one independent median per video.
    """
    offsets = [int(o) for o in offsets]
    outs = [temporal_median_int(frames[offsets[v]:offsets[v + 1]]) for v in range(len(offsets) - 1)]
    return np.stack(outs, axis=0) if outs else np.zeros((0,) + frames.shape[1:], np.uint8)


def bg_extraction_tmf_frames(decoded: np.ndarray, interval: int, max_frames: int) -> np.ndarray:
    """E1 on an already-decoded frame stack: selection (52-60) then median (73)."""
    idx = select_frame_indices(len(decoded), interval, max_frames)
    return temporal_median_np(decoded[idx])


# --------------------------------------------------------------------------- #
# sharding of the work list (extract_background.py:128-133)
# --------------------------------------------------------------------------- #
def reference_contiguous_splits(n_items: int, num_workers: int) -> List[range]:
    """The reference's ceil-sized contiguous slices, one per worker (some may be empty)."""
    import math
    per = math.ceil(n_items / num_workers) if num_workers > 0 else 0
    out, start = [], 0
    for _ in range(num_workers):
        out.append(range(min(start, n_items), min(start + per, n_items)))
        start += per
    return out
