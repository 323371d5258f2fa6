"""CPU oracle for the background ``Resize(bg_resize)`` of the BG-mix path.

TEST INFRASTRUCTURE ONLY (see oracle/median_oracle.py for the rule): imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs, never by the product.

The reference applies ``torchvision.transforms.Resize(bg_resize)`` to ``read_image(...).float()``
(libs/loader/comix_loader.py:72,130,140).  The arithmetic lives in third-party code that is not in the
reference tree: torchvision 0.26 ``F.resize`` -> ``torch.nn.functional.interpolate(mode="bilinear",
antialias=True, align_corners=False)`` -> ATen ``_upsample_bilinear2d_aa`` CPU kernel
(aten/src/ATen/native/cpu/UpSampleKernel.cpp, torch 2.11).  Its published algorithm, restated in numpy:

* per axis, output index ``i``: ``scale = f32(in)/out``, ``support = max(scale, 1)``, ``center = f32(scale*(i+0.5))``,
  taps ``[min, min+size)`` with ``min = max(int(center-support+0.5), 0)``,
  ``size = min(int(center+support+0.5), in) - min``, weights ``tri((j+min-center+0.5)/max(scale,1))`` normalised by
  their float32 sum (``HelperInterpBase::_compute_indices_min_size_weights_aa``);
* horizontal pass over the whole image, then vertical pass (``separable_upsample_generic_Nd_kernel_impl``); an axis
  whose size does not change is skipped;
* one output = ``v0*w0`` then ``+= v_j*w_j`` in tap order (``interpolate_aa_single_dim``).  In the installed build
  (GCC 13, AVX-512 dispatch) the first ``4*floor((size-1)/4)`` of those steps round product and sum separately and the
  remaining ``(size-1) % 4`` steps are fused multiply-adds.  That split is a property of the compiled library, found
  by experiment; it is what makes this restatement bit-exact here.

Parity status: PINNED -- against torchvision itself at test time (tests/test_oracle_aa_resize.py runs
``Resize`` on random images of many sizes and demands equal bits) and through the reference's own outputs in
``tests/golden/bgmix_ragged_reference.npz`` (oracle/gen_golden_ragged.py).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

f32 = np.float32


def resized_hw(h: int, w: int, size: int) -> Tuple[int, int]:
    """Output (h, w) of torchvision ``Resize(int)``: short edge -> size, long edge truncated."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def aa_tables(in_size: int, out_size: int):
    """``(min[out], size[out], weights[out, K])`` of one axis, float32 arithmetic as ATen does it."""
    scale = f32(f32(in_size) / f32(out_size))
    support = f32(np.float64(scale)) if scale >= 1.0 else f32(1.0)
    K = int(math.ceil(support)) * 2 + 1
    invscale = f32(1.0 / np.float64(scale)) if scale >= 1.0 else f32(1.0)
    mins = np.zeros(out_size, np.int64)
    sizes = np.zeros(out_size, np.int64)
    W = np.zeros((out_size, K), f32)
    for i in range(out_size):
        center = f32(np.float64(scale) * (i + 0.5))
        mn = max(int(np.float64(f32(center - support)) + 0.5), 0)
        sz = min(int(np.float64(f32(center + support)) + 0.5), in_size) - mn
        sz = min(max(sz, 0), K)
        total = f32(0.0)
        for j in range(sz):
            x = f32((np.float64(f32(f32(j + mn) - center)) + 0.5) * np.float64(invscale))
            ax = abs(x)
            w = f32(1.0 - np.float64(ax)) if ax < 1.0 else f32(0.0)
            W[i, j] = w
            total = f32(total + w)
        if total != 0.0:
            for j in range(sz):
                W[i, j] = f32(W[i, j] / total)
        mins[i], sizes[i] = mn, sz
    return mins, sizes, W


def _pass_last_axis(a: np.ndarray, out_size: int) -> np.ndarray:
    n = a.shape[-1]
    if n == out_size:
        return a                                            # ATen skips an axis that keeps its size
    mins, sizes, W = aa_tables(n, out_size)
    out = np.zeros(a.shape[:-1] + (out_size,), f32)
    for i in range(out_size):
        mn, sz = int(mins[i]), int(sizes[i])
        acc = (a[..., mn] * W[i, 0]).astype(f32)
        unfused = ((sz - 1) // 4) * 4
        for j in range(1, sz):
            v = a[..., mn + j]
            if j <= unfused:
                acc = (acc + (v * W[i, j]).astype(f32)).astype(f32)
            else:       # fused multiply-add: one rounding (product and sum are exact in the 64-bit mantissa of x87 long double)
                acc = (v.astype(np.longdouble) * np.longdouble(W[i, j]) + acc.astype(np.longdouble)).astype(f32)
        out[..., i] = acc
    return out


def aa_resize(img_chw: np.ndarray, size) -> np.ndarray:
    """``Resize(size)(torch.as_tensor(img).float())`` for a ``[C, h, w]`` image; ``size`` int (short edge) or (H, W)."""
    a = np.asarray(img_chw).astype(f32)
    h, w = a.shape[-2:]
    H, Wd = resized_hw(h, w, size) if isinstance(size, int) else size
    a = _pass_last_axis(a, Wd)                              # horizontal first
    a = np.swapaxes(_pass_last_axis(np.swapaxes(a, -1, -2), H), -1, -2)
    return np.ascontiguousarray(a)
