"""GPU-resident background pool for the BG-mix blend.

The reference keeps the pool as a list of JPEG paths (``BackgroundMixDataset.bg_files``,
libs/loader/comix_loader.py:84-103) and decodes + resizes one image per mixed sample inside a
DataLoader worker (:126-140).  Here the pool is decoded and resized ONCE, with the very ops the
reference uses (``torchvision.io.read_image`` and ``torchvision.transforms.Resize``), and kept on the
device as fp32 ``[P, 3, Hb, Wb]``; the blend kernel then crops, normalises and mixes.  Index ``i`` of
the pool is index ``i`` of ``bg_files`` -- the reference's ``bg_idx``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def resized_hw(h: int, w: int, size: int) -> tuple:
    """Output size of torchvision ``Resize(int)``: short edge -> ``size``, long edge truncated."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def resize_like_reference(img_chw: torch.Tensor, size: Optional[int]) -> torch.Tensor:
    """``Resize(size)`` exactly as bg_pipeline applies it (comix_loader.py:72): float CHW tensor,
    torchvision defaults.  ``size=None`` skips the resize."""
    img = img_chw.float()
    if size is None:
        return img
    from torchvision.transforms import Resize
    return Resize(size)(img)


def _map_in_order(fn, items, workers: Optional[int] = None) -> list:
    """``[fn(x) for x in items]`` on a small thread pool (JPEG decode and torch's resize release the GIL); a pool of
    13,320 UCF101 backgrounds takes ~50 s to decode + resize on one thread."""
    items = list(items)
    if workers is None:
        import os
        workers = max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    if workers <= 1 or len(items) < 8:
        return [fn(x) for x in items]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(workers) as ex:
        return list(ex.map(fn, items))


class BackgroundPool:
    """Backgrounds after ``Resize``, resident on one device.

    ``tensor``: fp32 (or uint8 when built with ``keep_uint8=True`` and no resize) ``[P, 3, Hb, Wb]``.
    ``names``:  one label per slot (file path or video name), same order as ``bg_files``.
    """

    def __init__(self, tensor: torch.Tensor, names: Sequence[str]):
        if tensor.dim() != 4 or tensor.shape[1] != 3:
            raise ValueError("pool tensor must be [P, 3, Hb, Wb]")
        if len(names) != tensor.shape[0]:
            raise ValueError("one name per pool slot")
        self.tensor = tensor
        self.names: List[str] = list(names)
        self.index_names: List[str] = self.names          # pool index i -> name (differs from `names` for a store view)
        self.slots: Optional[torch.Tensor] = None         # int32 [len(index_names)]: pool index -> row of `tensor`; None = identity

    def __len__(self) -> int:
        return len(self.index_names)

    @property
    def hw(self) -> tuple:
        return int(self.tensor.shape[2]), int(self.tensor.shape[3])

    def rows(self, bg_idx: torch.Tensor) -> torch.Tensor:
        """Rows of ``tensor`` for pool indices ``bg_idx`` (int32, on the pool's device)."""
        if self.slots is None:
            return bg_idx
        return self.slots[bg_idx.clamp(min=0, max=len(self.index_names) - 1).long()]

    @classmethod
    def from_images(cls, images: Sequence, names: Optional[Sequence[str]] = None, bg_resize: Optional[int] = 256,
                    device="cuda", keep_uint8: bool = False) -> "BackgroundPool":
        """``images``: uint8 RGB arrays/tensors ``[3, h, w]`` (what ``read_image(mode=RGB)`` returns).
        All images must resize to the same ``(Hb, Wb)`` -- true for a dataset of one resolution."""
        if len(images) == 0:
            raise ValueError("empty background pool")     # reference: torch.randint(0, ...) raises at draw time
        def prepare(im):
            t = torch.as_tensor(im)
            if t.dim() != 3 or t.shape[0] != 3:
                raise ValueError("background images must be [3, h, w]")
            return t.to(torch.uint8) if (keep_uint8 and bg_resize is None) else resize_like_reference(t, bg_resize)

        out = _map_in_order(prepare, images)
        shapes = {tuple(t.shape) for t in out}
        if len(shapes) != 1:
            raise ValueError(f"backgrounds resize to different shapes {sorted(shapes)}; build one pool per shape")
        pool = torch.stack(out).contiguous()
        names = list(names) if names is not None else [str(i) for i in range(len(out))]
        return cls(pool.to(device, non_blocking=True), names)

    @classmethod
    def from_files(cls, bg_files: Sequence[str], bg_resize: Optional[int] = 256, device="cuda") -> "BackgroundPool":
        """Decode with ``torchvision.io.read_image(mode=RGB)`` like ``_get_bg_image`` (comix_loader.py:130)."""
        from torchvision.io import ImageReadMode, read_image
        imgs = _map_in_order(lambda f: read_image(str(f), mode=ImageReadMode.RGB), bg_files)
        return cls.from_images(imgs, [str(f) for f in bg_files], bg_resize, device)

    @classmethod
    def from_index(cls, bg_dir, bg_resize: Optional[int] = 256, device="cuda") -> "BackgroundPool":
        """Rebuild the pool from the index the extraction CLI writes beside its JPEG folder
        (``extract_background.write_background_index``, ``<bg_dir>.bg_index.json``): same order on every rank,
        no video is read."""
        import json
        import pathlib
        bg_dir = pathlib.Path(bg_dir)
        index = bg_dir.parent / (bg_dir.name + ".bg_index.json")
        entries = json.loads(index.read_text())
        if not entries:
            raise ValueError(f"{index} lists no backgrounds")
        return cls.from_files([str(bg_dir / e["file"]) for e in entries], bg_resize, device)

    @staticmethod
    def all_gather_index(local_names: Sequence[str], group=None) -> list:
        """Index-only form of :meth:`all_gather` for pools too large to replicate (Sth-Sth-v2: ~68 GB of uint8
        backgrounds): every rank learns ``[(name, owner_rank, slot_on_owner), ...]`` in rank order -- the global
        pool index -- while the pixels stay sharded where they were extracted."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        names_per_rank: List[Optional[list]] = [None] * world
        dist.all_gather_object(names_per_rank, list(local_names), group=group)
        return [(n, r, i) for r, per in enumerate(names_per_rank) for i, n in enumerate(per)]

    # ---- one collective: assemble the pool from per-rank shards ----------
    @staticmethod
    def all_gather(local_names: Sequence[str], local_bgs: torch.Tensor, group=None):
        """All-gather per-rank backgrounds ``[P_r, ...]`` (uint8 or fp32, any device the backend
        supports) and their names.  Returns ``(names, tensor)`` in rank order, identical on every
        rank -- the index the single-process run would have produced for the rank-ordered list.
        Shards may differ in size: counts are exchanged first, pixels travel padded to the max."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        names_per_rank: List[Optional[list]] = [None] * world
        dist.all_gather_object(names_per_rank, list(local_names), group=group)
        counts = [len(n) for n in names_per_rank]
        if local_bgs.shape[0] != len(local_names):
            raise ValueError("one name per local background")
        max_n = max(counts) if counts else 0
        item_shape = tuple(local_bgs.shape[1:])
        padded = local_bgs.new_zeros((max_n,) + item_shape)
        padded[: local_bgs.shape[0]] = local_bgs
        gathered = local_bgs.new_empty((world * max_n,) + item_shape)
        dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
        gathered = gathered.view((world, max_n) + item_shape)
        parts = [gathered[r, : counts[r]] for r in range(world)]
        names = [n for per in names_per_rank for n in per]
        return names, torch.cat(parts, 0) if parts else gathered.view((0,) + item_shape)


class BackgroundStore:
    """Backgrounds resident on one device across pool changes.

    The reference's trainer rewrites ``dataset.bg_files`` between CIL tasks -- ``keep_all_backgrounds``
    (libs/cil/cil.py:193-195,690-694), ``cbf_full_bg`` (:146-160) and ``merge_bg_files`` (:390-393) are
    unions / extensions of path lists -- and every sample then decodes its background again.  Here each
    path is decoded + resized once, the pixels stay in one ``[capacity, 3, Hb, Wb]`` tensor, and a pool is
    an int32 slot table over it: set operations on the path lists never copy or re-decode pixels, and
    duplicates in ``bg_files`` (``extend`` creates them) share a slot and keep their draw probability.
    """

    def __init__(self, bg_resize: Optional[int] = 256, device="cuda", keep_uint8: bool = False):
        self.bg_resize = bg_resize
        self.device = torch.device(device)
        self.keep_uint8 = keep_uint8
        self.tensor: Optional[torch.Tensor] = None        # [capacity, 3, Hb, Wb]
        self.slot_of: dict = {}                           # name -> slot
        self.decoded = 0                                  # images decoded so far (what a rebuild would repeat)

    def __len__(self) -> int:
        return len(self.slot_of)

    @property
    def hw(self) -> tuple:
        if self.tensor is None:
            raise ValueError("empty background store")
        return int(self.tensor.shape[2]), int(self.tensor.shape[3])

    def ensure(self, names: Sequence[str], reader) -> None:
        """Make every name resident; ``reader(name) -> uint8 [3, h, w]`` is called for new names only."""
        new = [n for n in dict.fromkeys(names) if n not in self.slot_of]
        if not new:
            return
        def prepare(n):
            t = torch.as_tensor(reader(n))
            if t.dim() != 3 or t.shape[0] != 3:
                raise ValueError("background images must be [3, h, w]")
            return t.to(torch.uint8) if (self.keep_uint8 and self.bg_resize is None) else resize_like_reference(t, self.bg_resize)

        imgs = _map_in_order(prepare, new)
        shapes = {tuple(t.shape) for t in imgs}
        if self.tensor is not None:
            shapes.add(tuple(self.tensor.shape[1:]))
        if len(shapes) != 1:
            raise ValueError(f"backgrounds resize to different shapes {sorted(shapes)}; build one pool per shape")
        used = len(self.slot_of)
        need = used + len(imgs)
        if self.tensor is None or need > self.tensor.shape[0]:
            cap = max(need, 2 * (self.tensor.shape[0] if self.tensor is not None else 0))
            grown = torch.empty((cap,) + tuple(imgs[0].shape), dtype=imgs[0].dtype, device=self.device)
            if self.tensor is not None and used:
                grown[:used].copy_(self.tensor[:used])
            self.tensor = grown
        self.tensor[used:need].copy_(torch.stack(imgs), non_blocking=True)
        for i, n in enumerate(new):
            self.slot_of[n] = used + i
        self.decoded += len(new)

    def view(self, names: Sequence[str]) -> "BackgroundPool":
        """The pool whose index ``i`` is ``names[i]`` (the reference's ``bg_idx``): shares this store's
        pixels; ``slots`` maps pool index -> row of ``tensor``."""
        if len(names) == 0:
            raise ValueError("empty background pool")
        missing = [n for n in names if n not in self.slot_of]
        if missing:
            raise KeyError(f"{len(missing)} backgrounds are not resident (first: {missing[0]}); call ensure() first")
        pool = BackgroundPool(self.tensor[:len(self.slot_of)], list(self.slot_of))
        pool.index_names = list(names)
        pool.slots = torch.tensor([self.slot_of[n] for n in names], dtype=torch.int32, device=self.device)
        return pool
