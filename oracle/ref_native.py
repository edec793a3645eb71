"""TEST INFRASTRUCTURE — loader for the reference's own native code compiled into oracle/_ref/ by build_ref.py.
`available()` is False where it was never built (tests then skip)."""
import glob
import importlib.util
import os
import sys

import numpy as np

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def _load(stem):
    if stem in _mods:
        return _mods[stem]
    files = glob.glob(os.path.join(REF, stem + "*.so"))
    if not files:
        return None
    if not hasattr(np, "float"):
        np.float = float            # bbox.pyx:12 uses the removed alias
    if stem == "nlv_ref_C":
        import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(stem, files[0])
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    _mods[stem] = m
    return m


def available():
    return all(glob.glob(os.path.join(REF, s + "*.so")) for s in ("draw_rectangles", "bbox", "nlv_ref_C"))


def draw_union_boxes(bp, ps=27):
    return _load("draw_rectangles").draw_union_boxes(np.ascontiguousarray(bp, dtype=np.float32), ps)


def bbox_overlaps(a, b):
    return _load("bbox").bbox_overlaps(a, b)


def roi_align_forward(x, rois, scale, ph, pw, sr):
    return _load("nlv_ref_C").roi_align_forward(x, rois, scale, ph, pw, sr)


def nms(dets, scores, thr):
    return _load("nlv_ref_C").nms(dets, scores, thr)
