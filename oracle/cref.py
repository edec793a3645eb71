"""TEST INFRASTRUCTURE — ctypes front-end of oracle/cref.c (built by `make -C oracle`)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def draw_union_boxes(box_pairs: np.ndarray, pooling_size: int = 27) -> np.ndarray:
    bp = np.ascontiguousarray(box_pairs, dtype=np.float32)
    n = bp.shape[0]
    out = np.empty((n, 2, pooling_size, pooling_size), dtype=np.float32)
    lib().oracle_draw_union_boxes(_p(bp, ctypes.c_float), n, int(pooling_size), _p(out, ctypes.c_float))
    return out


def bbox_overlaps(boxes: np.ndarray, query: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4)
    q = np.ascontiguousarray(query, dtype=np.float64).reshape(-1, 4)
    out = np.empty((b.shape[0], q.shape[0]), dtype=np.float64)
    lib().oracle_bbox_overlaps(_p(b, ctypes.c_double), b.shape[0], _p(q, ctypes.c_double), q.shape[0],
                               _p(out, ctypes.c_double))
    return out


def nms(dets: np.ndarray, scores: np.ndarray, thr: float, strict: bool = False) -> np.ndarray:
    """Kept original indices, ascending.  strict=False: CPU reference (>=); True: CUDA reference (>)."""
    d = np.ascontiguousarray(dets, dtype=np.float32).reshape(-1, 4)
    n = d.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    import torch
    order = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).sort(0, descending=True)[1].numpy()
    order = np.ascontiguousarray(order, dtype=np.int64)
    keep = np.empty(n, dtype=np.int64)
    fn = lib().oracle_nms
    fn.restype = ctypes.c_int
    m = fn(_p(d, ctypes.c_float), _p(order, ctypes.c_int64), n, ctypes.c_float(thr), int(strict),
           _p(keep, ctypes.c_int64))
    return keep[:m].copy()


def roi_align_forward(inp: np.ndarray, rois: np.ndarray, scale: float, ph: int, pw: int, sampling_ratio: int):
    x = np.ascontiguousarray(inp, dtype=np.float32)
    r = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    B, C, H, W = x.shape
    out = np.empty((r.shape[0], C, ph, pw), dtype=np.float32)
    lib().oracle_roi_align_fwd(_p(x, ctypes.c_float), B, C, H, W, _p(r, ctypes.c_float), r.shape[0],
                               ctypes.c_float(scale), ph, pw, sampling_ratio, _p(out, ctypes.c_float))
    return out
