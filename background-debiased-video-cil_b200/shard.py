"""Per-GPU shard launcher: one process and one GPU per shard of the video list.

Mirrors the worker fan-out of ``cil_tools/extract_background.py:128-133,154-162`` (ceil-sized
contiguous slices, one ``multiprocessing.Process`` each) with two changes: the list is sorted first
(the reference iterates a ``set``, so its split is arbitrary), and processes are spawned, not
forked, because each one opens a CUDA context on its own device.
"""
from __future__ import annotations

import math
import multiprocessing as mp
from typing import Callable, List, Sequence


def contiguous_splits(items: Sequence, num_workers: int) -> List[list]:
    """The reference's split (extract_background.py:128-133): ``ceil(len/num_workers)`` items per
    worker, contiguous, later workers possibly empty."""
    if num_workers <= 0:
        raise ValueError("num_workers must be positive")
    per = math.ceil(len(items) / num_workers)
    return [list(items[i * per:(i + 1) * per]) for i in range(num_workers)]


def rank_slice(items: Sequence, rank: int, world: int) -> list:
    """Shard of ``rank`` among ``world`` equal-role processes (torchrun style)."""
    return contiguous_splits(items, world)[rank]


def device_for(rank: int, n_devices: int) -> int:
    if n_devices <= 0:
        raise RuntimeError("no CUDA device visible: bgdebias_b200 has no CPU fallback")
    return rank % n_devices


def _entry(fn, rank, shard, args, kwargs, errq):
    try:
        fn(rank, shard, *args, **kwargs)
    except BaseException as e:  # surface the failure in the parent (the reference silently joins)
        errq.put((rank, repr(e)))
        raise


def run_shards(fn: Callable, shards: Sequence[Sequence], *args, **kwargs) -> None:
    """Run ``fn(rank, shard, *args, **kwargs)`` in one spawned process per non-empty shard and wait.
    Raises ``RuntimeError`` listing the shards that failed (the reference joins crashed workers
    without noticing, extract_background.py:161-162)."""
    live = [(r, s) for r, s in enumerate(shards) if len(s)]
    if not live:
        return
    if len(live) == 1:
        fn(live[0][0], live[0][1], *args, **kwargs)
        return
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    procs = []
    for rank, shard in live:
        p = ctx.Process(target=_entry, args=(fn, rank, list(shard), args, kwargs, errq))
        p.start()
        procs.append((rank, p))
    failed = []
    for rank, p in procs:
        p.join()
        if p.exitcode != 0:
            failed.append(rank)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    if failed:
        raise RuntimeError(f"extraction shards {failed} failed: {msgs}")
