"""GPU-resident background pool for the BG-mix blend.

The reference keeps the pool as a list of JPEG paths (``BackgroundMixDataset.bg_files``,
libs/loader/comix_loader.py:84-103) and decodes + resizes one image per mixed sample inside a
DataLoader worker (:126-140).  Here the pool is decoded and resized ONCE, with the very ops the
reference uses (``torchvision.io.read_image`` and ``torchvision.transforms.Resize``), and kept on the
device as fp32 ``[P, 3, Hb, Wb]``; the blend kernel then crops, normalises and mixes.  Index ``i`` of
the pool is index ``i`` of ``bg_files`` -- the reference's ``bg_idx``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

SLOT_DTYPE = np.dtype([("offset", "<i8"), ("h", "<i4"), ("w", "<i4"), ("Hb", "<i4"), ("Wb", "<i4"),
                       ("xtab", "<i4"), ("ytab", "<i4"), ("kx", "<i4"), ("ky", "<i4")])   # bgd_ragged_slot, 40 bytes


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


class AATables:
    """Weight tables of torchvision's antialiased bilinear ``Resize`` (ATen's separable CPU kernel), one per distinct
    ``(in_size, out_size)``, evaluated on the host by ``bgd_aa_resize_table`` and kept in one int32 device tensor."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._host: List[np.ndarray] = []
        self._where: Dict[Tuple[int, int], Tuple[int, int]] = {}
        self._words = 0
        self._dev: Optional[torch.Tensor] = None

    def lookup(self, in_size: int, out_size: int) -> Tuple[int, int]:
        """(word offset, taps) of the table for this axis; (-1, 0) when the size does not change (the pass is skipped)."""
        if in_size == out_size:
            return -1, 0
        key = (int(in_size), int(out_size))
        if key not in self._where:
            import ctypes
            from . import _cabi
            taps = ctypes.c_int32()
            _cabi.check(_cabi.lib().bgd_aa_resize_table(key[0], key[1], ctypes.byref(taps), None, 0))
            words = np.empty(key[1] * (2 + taps.value), dtype=np.int32)
            _cabi.check(_cabi.lib().bgd_aa_resize_table(key[0], key[1], ctypes.byref(taps), words.ctypes.data, words.size))
            self._where[key] = (self._words, int(taps.value))
            self._host.append(words)
            self._words += words.size
            self._dev = None
        return self._where[key]

    @property
    def tensor(self) -> torch.Tensor:
        if self._dev is None:
            host = np.concatenate(self._host) if self._host else np.zeros(1, np.int32)
            self._dev = torch.from_numpy(host).to(self.device)
        return self._dev


class RaggedPool:
    """Backgrounds of ANY sizes, uint8 at native size, in one device byte buffer plus a slot table -- the form
    ``bgdebias::bgmix_blend_ragged`` reads.  ``Resize(bg_resize)`` (comix_loader.py:72) happens inside the blend launch,
    so a pool costs ``3*h*w`` bytes per image (Sth-Sth-v2: ~68 GB instead of ~308 GB resized fp32) and images of
    different widths (HMDB51, Sth-Sth-v2) share one pool, each cropped at its own resized size as the reference does."""

    def __init__(self, bg_resize: Optional[int] = 256, device="cuda", tables: Optional[AATables] = None):
        self.bg_resize = bg_resize
        self.device = torch.device(device)
        self.tables = tables if tables is not None else AATables(self.device)
        self.data: Optional[torch.Tensor] = None          # uint8 [capacity] on the device
        self.used = 0                                     # bytes in use
        self.slots = np.zeros(0, dtype=SLOT_DTYPE)        # host copy of the slot table
        self._slots_dev: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return int(self.slots.shape[0])

    def make_slot(self, offset: int, h: int, w: int) -> np.void:
        Hb, Wb = (h, w) if self.bg_resize is None else resized_hw(h, w, self.bg_resize)
        xtab, kx = self.tables.lookup(w, Wb)
        ytab, ky = self.tables.lookup(h, Hb)
        return np.array([(offset, h, w, Hb, Wb, xtab, ytab, kx, ky)], dtype=SLOT_DTYPE)[0]

    def append(self, images: Sequence) -> None:
        """``images``: uint8 ``[3, h, w]`` arrays / tensors (what ``read_image(mode=RGB)`` returns), any sizes."""
        imgs = []
        for im in images:
            t = torch.as_tensor(im)
            if t.dim() != 3 or t.shape[0] != 3 or t.dtype != torch.uint8:
                raise ValueError("background images must be uint8 [3, h, w]")
            imgs.append(t.contiguous())
        if not imgs:
            return
        sizes = [int(t.numel()) for t in imgs]
        starts = [(o + 15) & ~15 for o in np.cumsum([0] + [(n + 15) & ~15 for n in sizes[:-1]]).tolist()]
        total = starts[-1] + sizes[-1]
        base = (self.used + 15) & ~15
        need = base + total
        if self.data is None or need > self.data.numel():
            cap = max(need, 2 * (self.data.numel() if self.data is not None else 0))
            grown = torch.empty(cap, dtype=torch.uint8, device=self.device)
            if self.data is not None and self.used:
                grown[:self.used].copy_(self.data[:self.used])
            self.data = grown
        staging = torch.empty(total, dtype=torch.uint8, pin_memory=self.device.type == "cuda")
        for t, o in zip(imgs, starts):
            staging[o:o + t.numel()] = t.reshape(-1)
        self.data[base:base + total].copy_(staging, non_blocking=True)
        new = np.zeros(len(imgs), dtype=SLOT_DTYPE)
        for i, (t, o) in enumerate(zip(imgs, starts)):
            new[i] = self.make_slot(base + o, int(t.shape[1]), int(t.shape[2]))
        self.slots = np.concatenate([self.slots, new])
        self.used = need
        self._slots_dev = None

    @classmethod
    def from_uniform_buffer(cls, data: torch.Tensor, n: int, h: int, w: int, bg_resize: Optional[int] = 256) -> "RaggedPool":
        """``n`` images of one size already resident back to back in ``data`` (uint8, ``n * 3 * h * w`` bytes, planar
        ``[3][h][w]`` each): adopt the buffer as a pool without copying -- e.g. the all-gathered backgrounds of a
        dataset of one resolution kept as uint8 (Sth-Sth-v2: 220,847 x 240 x 427 x 3 = 67.9 GB)."""
        if data.dtype != torch.uint8 or data.dim() != 1 or data.numel() < n * 3 * h * w:
            raise ValueError("data must be a flat uint8 buffer of at least n * 3 * h * w bytes")
        pool = cls(bg_resize, data.device)
        pool.data, pool.used = data, n * 3 * h * w
        one = pool.make_slot(0, h, w)
        slots = np.repeat(np.array([one], dtype=SLOT_DTYPE), n)
        slots["offset"] = np.arange(n, dtype=np.int64) * (3 * h * w)
        pool.slots = slots
        return pool

    @property
    def slots_tensor(self) -> torch.Tensor:
        """The slot table as a uint8 device tensor (``len(self) * 40`` bytes)."""
        if self._slots_dev is None:
            raw = np.ascontiguousarray(self.slots).view(np.uint8).reshape(-1)
            self._slots_dev = torch.from_numpy(raw.copy()).to(self.device)
        return self._slots_dev

    def hw(self, slot: int) -> Tuple[int, int]:
        """Size of image ``slot`` after Resize -- what RandomCrop draws its offsets from."""
        s = self.slots[slot]
        return int(s["Hb"]), int(s["Wb"])

    def uniform_hw(self) -> Optional[Tuple[int, int, int, int]]:
        """(h, w, Hb, Wb) if every image has the same size, else None."""
        if not len(self):
            return None
        s0 = self.slots[0]
        same = (self.slots["h"] == s0["h"]).all() and (self.slots["w"] == s0["w"]).all()
        return (int(s0["h"]), int(s0["w"]), int(s0["Hb"]), int(s0["Wb"])) if same else None

    def resized(self, slot: int) -> torch.Tensor:
        """fp32 ``[3, Hb, Wb]``: ``Resize(bg_resize)(read_image(...).float())`` of one image, computed on the device."""
        import ctypes
        from . import _cabi
        s = self.slots[slot]
        out = torch.empty((3, int(s["Hb"]), int(s["Wb"])), dtype=torch.float32, device=self.device)
        c_slot = _cabi.RaggedSlot(*[int(s[k]) for k in SLOT_DTYPE.names])
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().bgd_aa_resize_u8_f32(self.data.data_ptr(), ctypes.byref(c_slot), self.tables.tensor.data_ptr(),
                                                         out.data_ptr(), int(torch.cuda.current_stream(self.device).cuda_stream)))
        return out


class BackgroundPool:
    """Backgrounds after ``Resize``, resident on one device.

    ``tensor``: fp32 (or uint8 when built with ``keep_uint8=True`` and no resize) ``[P, 3, Hb, Wb]``, or ``None`` for a
    pool that only exists in ragged form (images of several sizes, or too large to keep resized).
    ``ragged``: the :class:`RaggedPool` behind a :class:`BackgroundStore` view.
    ``names``:  one label per slot (file path or video name), same order as ``bg_files``.
    """

    def __init__(self, tensor: Optional[torch.Tensor], names: Sequence[str], ragged: Optional["RaggedPool"] = None):
        if tensor is None and ragged is None:
            raise ValueError("a pool needs a dense tensor or a ragged store")
        if tensor is not None:
            if tensor.dim() != 4 or tensor.shape[1] != 3:
                raise ValueError("pool tensor must be [P, 3, Hb, Wb]")
            if len(names) != tensor.shape[0]:
                raise ValueError("one name per pool slot")
        self.tensor = tensor                              # dense form, or None: blend from `ragged`
        self.ragged = ragged                              # uint8 images of any sizes (a store view always has it)
        self._slots_host: Optional[List[int]] = None
        self.names: List[str] = list(names)
        self.index_names: List[str] = self.names          # pool index i -> name (differs from `names` for a store view)
        self.slots: Optional[torch.Tensor] = None         # int32 [len(index_names)]: pool index -> row of `tensor`; None = identity

    def __len__(self) -> int:
        return len(self.index_names)

    @property
    def hw(self) -> tuple:
        """(Hb, Wb) of a pool whose images all resize to one shape."""
        if self.tensor is not None:
            return int(self.tensor.shape[2]), int(self.tensor.shape[3])
        u = self.ragged.uniform_hw()
        if u is None:
            raise ValueError("backgrounds of several sizes: ask hw_of(bg_idx)")
        return u[2], u[3]

    def hw_of(self, bg_idx: int) -> tuple:
        """Size after Resize of the image pool index ``bg_idx`` draws -- what RandomCrop.get_params sees
        (comix_loader.py:73 applied to that image)."""
        if self.ragged is None:
            return self.hw
        slot = self._slots_host[bg_idx] if self._slots_host is not None else bg_idx
        return self.ragged.hw(slot)

    def rows(self, bg_idx: torch.Tensor) -> torch.Tensor:
        """Rows of ``tensor`` for pool indices ``bg_idx`` (int32, on the pool's device)."""
        if self.slots is None:
            return bg_idx
        return self.slots[bg_idx.clamp(min=0, max=len(self.index_names) - 1).long()]

    @classmethod
    def from_images(cls, images: Sequence, names: Optional[Sequence[str]] = None, bg_resize: Optional[int] = 256,
                    device="cuda", keep_uint8: bool = False, ragged: Optional[bool] = None) -> "BackgroundPool":
        """``images``: uint8 RGB arrays/tensors ``[3, h, w]`` (what ``read_image(mode=RGB)`` returns).

        Images of one size give a dense pool (resized with torchvision itself on the host); images of several sizes --
        or ``ragged=True`` -- give a ragged uint8 pool that is resized inside the blend launch."""
        if len(images) == 0:
            raise ValueError("empty background pool")     # reference: torch.randint(0, ...) raises at draw time
        tens = []
        for im in images:
            t = torch.as_tensor(im)
            if t.dim() != 3 or t.shape[0] != 3:
                raise ValueError("background images must be [3, h, w]")
            tens.append(t)
        names = list(names) if names is not None else [str(i) for i in range(len(tens))]
        if ragged is None:
            ragged = len({tuple(t.shape) for t in tens}) != 1
        if ragged:
            rp = RaggedPool(bg_resize, device)
            rp.append([t.to(torch.uint8) for t in tens])
            return cls(None, names, ragged=rp)
        prepare = lambda t: t.to(torch.uint8) if (keep_uint8 and bg_resize is None) else resize_like_reference(t, bg_resize)  # noqa: E731
        pool = torch.stack(_map_in_order(prepare, tens)).contiguous()
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
    def all_gather(local_names: Sequence[str], local_bgs: torch.Tensor, group=None, events=None):
        """All-gather per-rank backgrounds ``[P_r, ...]`` (uint8 or fp32, any device the backend
        supports) and their names.  Returns ``(names, tensor)`` in rank order, identical on every
        rank -- the index the single-process run would have produced for the rank-ordered list.

        The pixels are gathered IN PLACE: every rank's shard lands at its final position of the result, no padded
        staging copy and no concatenation.  Shards of equal size (and the contiguous ceil split of
        extract_background.py:128-133, where only the last shard is shorter) need nothing else; for any other
        mix of sizes one compacting copy follows.  ``events``: optional pair of CUDA events recorded around the pixel
        collective alone (the name exchange before it is host work)."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        names_per_rank: List[Optional[list]] = [None] * world
        dist.all_gather_object(names_per_rank, list(local_names), group=group)
        counts = [len(n) for n in names_per_rank]
        if local_bgs.shape[0] != len(local_names):
            raise ValueError("one name per local background")
        max_n = max(counts) if counts else 0
        item_shape = tuple(local_bgs.shape[1:])
        names = [n for per in names_per_rank for n in per]
        if max_n == 0:
            return names, local_bgs.new_empty((0,) + item_shape)
        send = local_bgs.contiguous()
        if send.shape[0] < max_n:                          # a short shard sends a padded copy (only its own, small, side)
            send = torch.cat([send, send.new_zeros((max_n - send.shape[0],) + item_shape)])
        gathered = local_bgs.new_empty((world * max_n,) + item_shape)
        if events is not None:
            events[0].record()
        dist.all_gather_into_tensor(gathered, send, group=group)
        if events is not None:
            events[1].record()
        total = sum(counts)
        if all(c == max_n for c in counts[:-1]):           # gaps only after the last shard: the result is a prefix view
            return names, gathered[:total]
        # any other mix of shard sizes (not produced by the extraction's split): one compacting copy
        return names, torch.cat([gathered[r * max_n:r * max_n + c] for r, c in enumerate(counts)], 0)


def all_gather_ragged(local_names: Sequence[str], local: "RaggedPool", group=None):
    """The same collective for backgrounds of mixed sizes: every rank contributes the byte buffer of its
    :class:`RaggedPool` (the decoded JPEGs it extracted) and the image sizes; the gathered ``[world, max_bytes]``
    buffer IS the pool of the result -- the slot table carries each image's offset, so nothing is moved or padded
    after the collective.  Returns ``(names, RaggedPool)`` in rank order, identical on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    meta = [(str(n), int(s["offset"]), int(s["h"]), int(s["w"])) for n, s in zip(local_names, local.slots)]
    if len(meta) != len(local):
        raise ValueError("one name per local background")
    metas: List[Optional[list]] = [None] * world
    dist.all_gather_object(metas, (int(local.used), meta), group=group)
    max_bytes = (max(u for u, _ in metas) + 15) & ~15
    out = RaggedPool(local.bg_resize, local.device)
    if max_bytes == 0:
        return [], out
    send = torch.zeros(max_bytes, dtype=torch.uint8, device=local.device)
    if local.used:
        send[:local.used].copy_(local.data[:local.used])
    out.data = torch.empty(world * max_bytes, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out.data, send, group=group)
    out.used = world * max_bytes
    names, slots = [], []
    for r, (_, per) in enumerate(metas):
        for n, off, h, w in per:
            names.append(n)
            slots.append(out.make_slot(r * max_bytes + off, h, w))
    out.slots = np.array(slots, dtype=SLOT_DTYPE) if slots else np.zeros(0, dtype=SLOT_DTYPE)
    return names, out


def gather_extracted_backgrounds(output_dir, image_suffix: str = ".jpg", bg_resize: Optional[int] = 256, device="cuda",
                                 group=None):
    """After a sharded extraction (``extract_background`` under torchrun: one rank = one GPU = one contiguous slice of
    the sorted video list, extract_background.py:128-133) every rank holds the JPEGs it wrote.  This is the path's one
    collective: each rank decodes ITS files with ``torchvision.io.read_image(mode=RGB)`` -- the JPEG round trip is part
    of what the reference mixes (libs/loader/comix_loader.py:130), so the raw medians are not gathered -- and the
    pixels + names are all-gathered over NCCL / NVLink into a pool resident on every GPU.  Returns
    ``(paths, RaggedPool)``: the same index and pixels as decoding the whole directory on one rank."""
    import pathlib
    import torch.distributed as dist
    from torchvision.io import ImageReadMode, read_image
    from .shard import contiguous_splits
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    files = sorted(str(f) for f in pathlib.Path(output_dir).glob("*" + image_suffix))
    mine = contiguous_splits(files, world)[rank] if files else []
    local = RaggedPool(bg_resize, device)
    local.append(_map_in_order(lambda f: read_image(f, mode=ImageReadMode.RGB), mine))
    return all_gather_ragged(mine, local, group)


class BackgroundStore:
    """Backgrounds resident on one device across pool changes.

    The reference's trainer rewrites ``dataset.bg_files`` between CIL tasks -- ``keep_all_backgrounds``
    (libs/cil/cil.py:193-195,690-694), ``cbf_full_bg`` (:146-160) and ``merge_bg_files`` (:390-393) are
    unions / extensions of path lists -- and every sample then decodes its background again.  Here each
    path is decoded once and stays on the device as uint8 at its native size (:class:`RaggedPool`: any mix of
    sizes, ``Resize`` evaluated inside the blend launch); a pool is an int32 slot table over the store, so set
    operations on the path lists never copy or re-decode pixels, and duplicates in ``bg_files`` (``extend``
    creates them) share a slot and keep their draw probability.

    When every image has the same size and the resized fp32 copy fits ``dense_budget_bytes`` (default: a quarter
    of the device's memory), :attr:`tensor` additionally caches ``Resize(bg_resize)`` of every image as fp32
    ``[n, 3, Hb, Wb]`` -- computed on the device by the same arithmetic -- and the blend reads that instead
    (one load per value instead of a 2x2..3x3 tap window).  Both forms give identical bits.
    """

    def __init__(self, bg_resize: Optional[int] = 256, device="cuda", keep_uint8: bool = False,
                 dense_budget_bytes: Optional[int] = None):
        self.bg_resize = bg_resize
        self.device = torch.device(device)
        self.keep_uint8 = keep_uint8                      # dense cache as uint8 (only without a resize)
        self.ragged = RaggedPool(bg_resize, self.device)
        self.slot_of: dict = {}                           # name -> slot
        self.decoded = 0                                  # images decoded so far (what a rebuild would repeat)
        self.dense_budget_bytes = dense_budget_bytes
        self._dense: Optional[torch.Tensor] = None        # [capacity, 3, Hb, Wb]
        self._dense_rows = 0

    def __len__(self) -> int:
        return len(self.slot_of)

    @property
    def hw(self) -> tuple:
        """(Hb, Wb) of a store whose images all have one size."""
        u = self.ragged.uniform_hw()
        if u is None:
            raise ValueError("empty background store" if not len(self) else
                             "backgrounds of several sizes: ask hw_of(slot)")
        return u[2], u[3]

    def hw_of(self, slot: int) -> tuple:
        return self.ragged.hw(slot)

    def ensure(self, names: Sequence[str], reader) -> None:
        """Make every name resident; ``reader(name) -> uint8 [3, h, w]`` is called for new names only."""
        new = [n for n in dict.fromkeys(names) if n not in self.slot_of]
        if not new:
            return
        def prepare(n):
            t = torch.as_tensor(reader(n))
            if t.dim() != 3 or t.shape[0] != 3:
                raise ValueError("background images must be [3, h, w]")
            return t.to(torch.uint8)

        imgs = _map_in_order(prepare, new)
        used = len(self.slot_of)
        self.ragged.append(imgs)
        for i, n in enumerate(new):
            self.slot_of[n] = used + i
        self.decoded += len(new)

    def adopt(self, names: Sequence[str], ragged: "RaggedPool") -> None:
        """Start from an already resident pool (``gather_extracted_backgrounds``): ``names[i]`` is image ``i`` of
        ``ragged``; nothing is decoded.  Only for an empty store."""
        if len(self.slot_of):
            raise ValueError("adopt() needs an empty store")
        if len(names) != len(ragged) or ragged.bg_resize != self.bg_resize:
            raise ValueError("one name per image, same bg_resize")
        self.ragged = ragged
        self.device = ragged.device
        self.slot_of = {}
        for i, n in enumerate(names):
            self.slot_of.setdefault(n, i)                # a repeated name keeps its first image

    def _budget(self) -> int:
        if self.dense_budget_bytes is not None:
            return int(self.dense_budget_bytes)
        if self.device.type == "cuda":
            return int(torch.cuda.get_device_properties(self.device).total_memory // 4)
        return 1 << 34

    @property
    def tensor(self) -> Optional[torch.Tensor]:
        """Dense cache ``[len(self), 3, Hb, Wb]`` (fp32 after Resize, or uint8 with ``keep_uint8`` and no resize) for a
        store of one image size within the budget, else ``None`` (the blend then reads the ragged store)."""
        n = len(self)
        u = self.ragged.uniform_hw()
        if u is None or n == 0:
            self._dense, self._dense_rows = None, 0
            return None
        h, w, Hb, Wb = u
        as_u8 = self.keep_uint8 and self.bg_resize is None
        if n * 3 * Hb * Wb * (1 if as_u8 else 4) > self._budget():
            self._dense, self._dense_rows = None, 0
            return None
        if self._dense is None or self._dense.shape[0] < n:
            cap = max(n, 2 * (self._dense.shape[0] if self._dense is not None else 0))
            grown = torch.empty((cap, 3, Hb, Wb), dtype=torch.uint8 if as_u8 else torch.float32, device=self.device)
            if self._dense is not None and self._dense_rows:
                grown[:self._dense_rows].copy_(self._dense[:self._dense_rows])
            self._dense = grown
        for slot in range(self._dense_rows, n):
            s = self.ragged.slots[slot]
            if (h, w) == (Hb, Wb) or self.device.type != "cuda":
                raw = self.ragged.data[int(s["offset"]):int(s["offset"]) + 3 * h * w].view(3, h, w)
                self._dense[slot].copy_(raw if as_u8 else resize_like_reference(raw, None if (h, w) == (Hb, Wb) else self.bg_resize))
            else:
                self._dense[slot].copy_(self.ragged.resized(slot))
        self._dense_rows = n
        return self._dense[:n]

    def view(self, names: Sequence[str]) -> "BackgroundPool":
        """The pool whose index ``i`` is ``names[i]`` (the reference's ``bg_idx``): shares this store's
        pixels; ``slots`` maps pool index -> store slot."""
        if len(names) == 0:
            raise ValueError("empty background pool")
        missing = [n for n in names if n not in self.slot_of]
        if missing:
            raise KeyError(f"{len(missing)} backgrounds are not resident (first: {missing[0]}); call ensure() first")
        pool = BackgroundPool(self.tensor, list(self.slot_of), ragged=self.ragged)
        pool.index_names = list(names)
        pool.slots = torch.tensor([self.slot_of[n] for n in names], dtype=torch.int32, device=self.device)
        pool._slots_host = [self.slot_of[n] for n in names]
        return pool
