"""CPU oracle for the ActorCutMix blend.

TEST INFRASTRUCTURE ONLY (see median_oracle.py for the rules).

Restates the per-frame loop and the foreground ratio of ``ActorCutMixDataset.actor_cut_mix`` /
``_calc_foreground_ratio`` (libs/loader/actor_cut_mix_loader.py:135-163):

    actor_cut_mix = actor_img * actor_mask + scene_img * (1 - actor_mask)      :143-148
    foreground_ratio = sum_t human_mask[t][:, :, 0].sum() / (T * w * h)         :154-163

Images and masks are uint8 ``[H, W, 3]`` arrays (masks are 0/1, built by the box pipeline, libs/pipelines/
box.py:187-206), so every operation is numpy uint8 arithmetic (wrapping), restated here in wider integers.

Parity status: PINNED.  ``oracle/gen_golden_acm.py`` runs the reference method itself (stub pipelines feed it
seeded frames and masks) and commits inputs and outputs under ``tests/golden/acm_reference.npz``.
"""
from __future__ import annotations

import numpy as np


def cut_mix(actor: np.ndarray, mask: np.ndarray, scene: np.ndarray) -> np.ndarray:
    """uint8 ``actor * mask + scene * (1 - mask)`` with numpy's uint8 wrap-around, any shape."""
    a, m, s = (np.asarray(x, np.uint8).astype(np.uint32) for x in (actor, mask, scene))
    inv = (1 - m) & 0xFF
    return ((((a * m) & 0xFF) + ((s * inv) & 0xFF)) & 0xFF).astype(np.uint8)


def foreground_ratio(masks: np.ndarray) -> float:
    """``masks``: uint8 ``[T, H, W, 3]``; channel 0 summed over all frames / (T * H * W) (all channels are equal)."""
    masks = np.asarray(masks, np.uint8)
    T, H, W = masks.shape[:3]
    return float(masks[..., 0].astype(np.uint64).sum()) / float(T * W * H)
