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

    def __len__(self) -> int:
        return self.tensor.shape[0]

    @property
    def hw(self) -> tuple:
        return int(self.tensor.shape[2]), int(self.tensor.shape[3])

    @classmethod
    def from_images(cls, images: Sequence, names: Optional[Sequence[str]] = None, bg_resize: Optional[int] = 256,
                    device="cuda", keep_uint8: bool = False) -> "BackgroundPool":
        """``images``: uint8 RGB arrays/tensors ``[3, h, w]`` (what ``read_image(mode=RGB)`` returns).
        All images must resize to the same ``(Hb, Wb)`` -- true for a dataset of one resolution."""
        if len(images) == 0:
            raise ValueError("empty background pool")     # reference: torch.randint(0, ...) raises at draw time
        out = []
        for im in images:
            t = torch.as_tensor(im)
            if t.dim() != 3 or t.shape[0] != 3:
                raise ValueError("background images must be [3, h, w]")
            if keep_uint8 and bg_resize is None:
                out.append(t.to(torch.uint8))
            else:
                out.append(resize_like_reference(t, bg_resize))
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
        imgs = [read_image(str(f), mode=ImageReadMode.RGB) for f in bg_files]
        return cls.from_images(imgs, [str(f) for f in bg_files], bg_resize, device)

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
