"""Video-to-GPU staging: pinned double buffer between the host decoder and the median kernel.

The reference keeps a Python list of decoded frames per video and calls ``np.median`` on it
(cil_tools/extract_background.py:48-73).  Here decoded frames are written straight into a pinned
slab that holds several whole videos back to back (``[sum T, N]`` uint8 plus the offsets table);
when a slab is full it is copied to the device on its own stream, reduced with one
``temporal_median_varlen`` launch, and the ``[V, N]`` backgrounds are copied back -- while the
decoder is already filling the other slab.

``add_video`` may be called from many decoder threads at once: a thread reserves its rows under a lock and
copies its frames into the slab outside it (numpy releases the GIL for the copy), so packing scales with the
decoder threads instead of serialising on the caller.
"""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import ops as _ops  # noqa: F401  (registers torch.ops.bgdebias.*)


class _Slab:
    def __init__(self, nbytes: int, device: torch.device):
        with torch.cuda.device(device):
            self.host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.stream = torch.cuda.Stream(device)
            self.done = torch.cuda.Event()
        self.reset()

    def reset(self):
        self.used = 0
        self.N = None
        self.offsets: List[int] = [0]
        self.tags: list = []
        self.shapes: list = []
        self.out_host: Optional[torch.Tensor] = None
        self.pending = False
        self.filling = 0                   # reservations whose frames are still being copied in


class FrameStager:
    """``add_video(frames, tag)`` queues one decoded video (list/array of equal-shaped uint8 frames);
    ``on_result(tag, background_ndarray)`` is called for every video once its median is back on the
    host.  Call :meth:`flush` at the end."""

    def __init__(self, on_result: Callable, device="cuda", slab_mb: int = 512):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FrameStager needs a CUDA device: there is no CPU path")
        self.on_result = on_result
        self.slab_bytes = int(slab_mb) << 20
        self.slabs: List[Optional[_Slab]] = [None, None]
        self.cur = 0
        self._cv = threading.Condition()

    def add_video(self, frames: Sequence[np.ndarray], tag) -> None:
        """Thread-safe.  ``frames``: a list of equal-shaped uint8 frames or one ``[T, ...]`` uint8 array."""
        T = len(frames)
        if T == 0:
            # reference: np.median([]) -> nan -> cv2.imwrite raises (extract_background.py:73-74)
            raise ValueError(f"{tag}: no frames decoded")
        shape = tuple(frames[0].shape)
        stacked = isinstance(frames, np.ndarray)
        if stacked:
            if frames.dtype != np.uint8:
                raise ValueError(f"{tag}: frames must be uint8")
        else:
            for t, f in enumerate(frames):
                if tuple(f.shape) != shape or f.dtype != np.uint8:
                    raise ValueError(f"{tag}: frame {t} differs in shape or dtype")
        N = int(np.prod(shape))
        need = T * N
        with self._cv:
            slab = self._room_for(need, N)
            off = slab.used
            slab.N = N
            slab.used += need
            slab.offsets.append(slab.offsets[-1] + T)
            slab.tags.append(tag)
            slab.shapes.append(shape)
            slab.filling += 1
        try:
            view = slab.host[off:off + need].view(T, N).numpy()
            if stacked:
                view[:] = frames.reshape(T, N)
            else:
                for t, f in enumerate(frames):
                    view[t] = f.reshape(-1)
        finally:
            with self._cv:
                slab.filling -= 1
                self._cv.notify_all()

    def _room_for(self, need: int, N: int) -> _Slab:
        """The slab the next ``need`` bytes go to (lock held): submits the current one when it is full or holds
        frames of another size -- after every thread still copying into it has finished -- and moves to the other."""
        while True:
            i = self.cur
            slab = self.slabs[i]
            if slab is None:
                self.slabs[i] = slab = _Slab(max(self.slab_bytes, need), self.device)
                return slab
            if slab.pending:
                self._drain(i)
            if slab.used == 0 and slab.host.numel() < need:           # a video larger than the slab: grow it
                self.slabs[i] = slab = _Slab(need, self.device)
                return slab
            if slab.used + need <= slab.host.numel() and (slab.used == 0 or slab.N == N):
                return slab
            while slab.filling:
                self._cv.wait()
            if self.cur == i and self.slabs[i] is slab and slab.used and not slab.pending:    # nobody else did it meanwhile
                self._submit(i)
                self.cur ^= 1

    def _submit(self, i: int) -> None:
        slab = self.slabs[i]
        if slab is None or not slab.tags:
            return
        V, N = len(slab.tags), slab.N
        rows = slab.offsets[-1]
        # decoder threads land here too: the current device is per thread
        with torch.cuda.device(self.device), torch.cuda.stream(slab.stream):
            slab.dev[:slab.used].copy_(slab.host[:slab.used], non_blocking=True)
            out = torch.ops.bgdebias.temporal_median_varlen(slab.dev[:slab.used].view(rows, N),
                                                            torch.tensor(slab.offsets, dtype=torch.int64))
            slab.out_host = torch.empty((V, N), dtype=torch.uint8, pin_memory=True)
            slab.out_host.copy_(out, non_blocking=True)
            slab.done.record(slab.stream)
        slab.pending = True

    def _drain(self, i: int) -> None:
        slab = self.slabs[i]
        if slab is None or not slab.pending:
            return
        slab.done.synchronize()
        res = slab.out_host.numpy()
        for v, (tag, shape) in enumerate(zip(slab.tags, slab.shapes)):
            self.on_result(tag, res[v].reshape(shape).copy())
        slab.reset()

    def flush(self) -> None:
        """Call after every ``add_video`` has returned."""
        with self._cv:
            while any(sl is not None and sl.filling for sl in self.slabs):
                self._cv.wait()
            self._submit(self.cur)
            for i in (self.cur ^ 1, self.cur):
                self._drain(i)
