"""The C restatement (oracle/cref.c) against the reference's OWN native code compiled into oracle/_ref/
(Cython draw_union_boxes / bbox_overlaps, C++ roi_align_forward / nms).  Skipped where oracle/_ref was not built."""
import numpy as np
import pytest
import torch

from oracle import cref, ref_native

pytestmark = pytest.mark.skipif(not ref_native.available(), reason="oracle/_ref not built (python oracle/build_ref.py)")


def test_draw_union_boxes_vs_reference_cython():
    rng = np.random.default_rng(5)
    b = rng.uniform(0, 400, (300, 8)).astype(np.float32)
    b[:, 2:4] += b[:, 0:2] + 1; b[:, 6:8] += b[:, 4:6] + 1
    assert np.array_equal(cref.draw_union_boxes(b, 27), ref_native.draw_union_boxes(b, 27))


def test_bbox_overlaps_vs_reference_cython():
    rng = np.random.default_rng(6)
    x = rng.uniform(0, 100, (60, 4)); x[:, 2:] += x[:, :2]
    y = rng.uniform(0, 100, (45, 4)); y[:, 2:] += y[:, :2]
    assert np.array_equal(cref.bbox_overlaps(x, y), ref_native.bbox_overlaps(x, y))


def test_roi_align_vs_reference_cpp():
    rng = np.random.default_rng(7)
    inp = rng.standard_normal((2, 6, 38, 67)).astype(np.float32)
    rois = np.column_stack((rng.integers(0, 2, 30), rng.uniform(-20, 500, 30), rng.uniform(-20, 300, 30), rng.uniform(400, 1200, 30),
                            rng.uniform(250, 700, 30))).astype(np.float32)
    for sr in (0, 2):
        want = ref_native.roi_align_forward(torch.from_numpy(inp), torch.from_numpy(rois), 1 / 16., 7, 7, sr).numpy()
        assert np.array_equal(cref.roi_align_forward(inp, rois, 1 / 16., 7, 7, sr), want)


def test_nms_vs_reference_cpp():
    rng = np.random.default_rng(8)
    d = rng.uniform(0, 300, (200, 4)).astype(np.float32); d[:, 2:] = d[:, :2] + rng.uniform(5, 120, (200, 2)).astype(np.float32)
    s = rng.uniform(0, 1, 200).astype(np.float32)
    for thr in (0.3, 0.6):
        want = ref_native.nms(torch.from_numpy(d), torch.from_numpy(s), thr).numpy()
        assert np.array_equal(cref.nms(d, s, thr, strict=False), want)
