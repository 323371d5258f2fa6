"""CLI shim for ``cil_tools/type_b_and_c_bg.py`` of the reference: copy the extracted
backgrounds in which a person detector finds no person (COCO class 0) to ``--out_dir`` ("type B / C" pools,
type_b_and_c_bg.py:42-54).  Same flags (``-i/--image_dir``, ``-o/--out_dir``, ``--glob_pattern``, :13-20).

The detector is third-party (detectron2 ``mask_rcnn_R_50_FPN_3x`` with downloaded weights, :23-31) and cannot run
offline: it is a plug-in here.  ``--detector module:function`` names a factory returning ``predictor(img_bgr) -> iterable of
class ids``; the default builds the reference's detectron2 predictor and fails with a clear message when detectron2 is
absent.  Parity with the reference is unpinned for this file (no detector, no reference test); only the selection rule
(keep iff class 0 is not among the predictions, ``out_dir`` must not exist, :44,51-53) is restated and tested.
The reference dumps the raw detectron2 outputs with ``json.dump`` (:56-57), which raises for ``Instances`` objects; the
shim writes ``detection.json`` with the predicted class ids per file instead.
"""
from __future__ import annotations

import argparse
import importlib
import json
import pathlib
import shutil
from typing import Callable, Iterable, List

import cv2

PERSON = 0          # COCO class id the reference filters on (:52)


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Train a recognizer')
    parser.add_argument('-i', '--image_dir', required=True)
    parser.add_argument('-o', '--out_dir', required=True)
    parser.add_argument('--glob_pattern', default="*")
    parser.add_argument('--detector', default=None, help='module:function returning predictor(img_bgr) -> class ids')
    return parser.parse_args(argv)


def get_predictor() -> Callable:
    """The reference's detector (:23-31), wrapped to return class ids.  Needs detectron2 and its weights."""
    try:
        from detectron2 import model_zoo
        from detectron2.config import get_cfg
        from detectron2.engine import DefaultPredictor
    except ImportError as e:
        raise RuntimeError("type_b_and_c_bg needs a person detector: install detectron2 (the reference's choice) or pass "
                           "--detector module:function") from e
    cfg = get_cfg()
    cfg.merge_from_file(model_zoo.get_config_file("COCO-InstanceSegmentation/mask_rcnn_R_50_FPN_3x.yaml"))
    cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST = 0.3
    cfg.MODEL.WEIGHTS = model_zoo.get_checkpoint_url("COCO-InstanceSegmentation/mask_rcnn_R_50_FPN_3x.yaml")
    predictor = DefaultPredictor(cfg)
    return lambda img: [int(c) for c in predictor(img)['instances'].pred_classes]


def load_detector(spec: str) -> Callable:
    module, _, name = spec.partition(':')
    return getattr(importlib.import_module(module), name or 'get_predictor')()


def filter_backgrounds(image_files: Iterable[pathlib.Path], out_dir: pathlib.Path, predictor: Callable) -> List[dict]:
    """Copies every image without a person to ``out_dir``; returns one record per image (:47-54)."""
    records = []
    for im_file in image_files:
        img = cv2.imread(str(im_file))
        classes = [int(c) for c in predictor(img)]
        kept = PERSON not in classes
        if kept:
            shutil.copy(im_file, out_dir / im_file.name)
        records.append({"im_file": str(im_file), "pred_classes": classes, "copied": kept})
    return records


def main(argv=None):
    args = parse_args(argv)
    predictor = load_detector(args.detector) if args.detector else get_predictor()
    image_dir, out_dir = pathlib.Path(args.image_dir), pathlib.Path(args.out_dir)
    out_dir.mkdir(exist_ok=False, parents=True)                 # the reference refuses an existing directory (:44)
    records = filter_backgrounds(sorted(image_dir.glob(args.glob_pattern)), out_dir, predictor)
    with open("detection.json", "w") as f:
        json.dump(records, f)
    return records


if __name__ == '__main__':
    main()
