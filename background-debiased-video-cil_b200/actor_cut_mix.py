"""The pixel loop of the reference's ActorCutMix augmentation on the GPU.

``ActorCutMixDataset.actor_cut_mix`` (libs/loader/actor_cut_mix_loader.py:135-152) runs two mmaction pipelines
(the actor clip with its ``human_mask`` from the box pipeline, and a scene clip drawn with
``random.randrange(len(video_infos))``), then mixes frame by frame and measures the foreground ratio.  The
pipelines and the draw stay on the host; :func:`actor_cut_mix` below replaces lines :143-152 -- the per-frame
``actor * mask + scene * (1 - mask)`` and ``_calc_foreground_ratio`` -- with one fused launch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops as _ops  # noqa: F401  (registers torch.ops.bgdebias.*)


def actor_cut_mix(result: dict, scene_video: dict, device="cuda") -> dict:
    """``result``: the actor pipeline's output (``imgs``: list of uint8 ``[H,W,3]`` frames, ``human_mask``: list of
    0/1 uint8 masks); ``scene_video``: the scene pipeline's output (``imgs``, ``label``).  Mutates and returns
    ``result`` like the reference: ``imgs`` replaced by the mixed frames, ``foreground_ratio`` and
    ``background_label`` set."""
    actor = torch.from_numpy(np.stack(result['imgs'])).to(device, non_blocking=True)
    mask = torch.from_numpy(np.stack(result['human_mask'])).to(device, non_blocking=True)
    scene = torch.from_numpy(np.stack(scene_video['imgs'][:len(result['imgs'])])).to(device, non_blocking=True)
    mixed, mask_sum = torch.ops.bgdebias.actor_cut_mix(actor, mask, scene)
    mixed = mixed.cpu().numpy()
    t, h, w = actor.shape[:3]
    for frame_idx in range(len(result['imgs'])):
        result['imgs'][frame_idx] = mixed[frame_idx]
    result['foreground_ratio'] = int(mask_sum.item()) / (t * w * h)
    result['background_label'] = scene_video['label']
    return result
