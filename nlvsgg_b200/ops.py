"""Thin torch-tensor front-end over the C-ABI (no arithmetic here: pointers, sizes and the stream)."""
from __future__ import annotations

import ctypes

import torch

from . import _C
from ._C import NLV_BF16, NLV_F32, MAJOR_K, MAJOR_MN


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _dt(t) -> int:
    if t.dtype == torch.float32:
        return NLV_F32
    if t.dtype == torch.bfloat16:
        return NLV_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nlv_b200 kernels need CUDA tensors; there is no CPU fallback")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_major: int = MAJOR_K, b_major: int = MAJOR_K,
         bias: torch.Tensor | None = None, residual: torch.Tensor | None = None, relu: bool = False,
         force_simt: bool = False) -> torch.Tensor:
    """out[m,n] = act(sum_k A(m,k) B(n,k) + bias[n]) + residual[m,n].

    a: [m,k] (K-major) or [k,m] (MN-major); b: [n,k] or [k,n]; rows may be strided (stride(1) == 1)."""
    _need_cuda(a, b, out, bias, residual)
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    assert a.dtype == b.dtype
    m, k = (a.shape if a_major == MAJOR_K else (a.shape[1], a.shape[0]))
    n, kb = (b.shape if b_major == MAJOR_K else (b.shape[1], b.shape[0]))
    assert k == kb, f"gemm: k mismatch {k} vs {kb}"
    assert tuple(out.shape) == (m, n), f"gemm: out shape {tuple(out.shape)} != {(m, n)}"
    g = _C.GemmArgs()
    g.a, g.b, g.d = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.residual = residual.data_ptr() if residual is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n and bias.is_contiguous()
    g.m, g.n, g.k = m, n, k
    g.lda, g.ldb, g.ldd = a.stride(0), b.stride(0), out.stride(0)
    g.ldr = residual.stride(0) if residual is not None else 0
    if residual is not None:
        assert tuple(residual.shape) == (m, n) and residual.stride(1) == 1
    g.a_major, g.b_major = a_major, b_major
    g.ab_dtype = _dt(a) | (_C.FORCE_SIMT if force_simt else 0)
    g.d_dtype = _dt(out)
    g.r_dtype = _dt(residual) if residual is not None else NLV_F32
    g.relu = 1 if relu else 0
    _C.check(_C.lib().nlv_gemm(ctypes.byref(g), _stream()), "gemm")
    return out


def draw_union_boxes(box_pairs: torch.Tensor, pooling_size: int = 27, offset: float = 0.0) -> torch.Tensor:
    """lib/draw_rectangles draw_union_boxes on device: f32[r,8] -> f32[r,2,ps,ps] (+offset)."""
    _need_cuda(box_pairs)
    bp = box_pairs.contiguous().float()
    r = bp.shape[0]
    out = torch.empty(r, 2, pooling_size, pooling_size, device=bp.device, dtype=torch.float32)
    _C.check(_C.lib().nlv_draw_union_boxes(_ptr(bp), r, pooling_size, ctypes.c_float(offset), _ptr(out), _stream()),
             "draw_union_boxes")
    return out


def union_mask_pairs(boxes: torch.Tensor, pair_idx: torch.Tensor, pooling_size: int = 27, offset: float = -0.5):
    """spatial_masks of lib/sttran.py:279-281 fused with the pair gather."""
    _need_cuda(boxes, pair_idx)
    bx = boxes.contiguous().float()
    pi = pair_idx.contiguous().to(torch.int64)
    r = pi.shape[0]
    out = torch.empty(r, 2, pooling_size, pooling_size, device=bx.device, dtype=torch.float32)
    _C.check(_C.lib().nlv_union_mask_pairs(_ptr(bx), _ptr(pi), r, pooling_size, ctypes.c_float(offset), _ptr(out),
                                           _stream()), "union_mask_pairs")
    return out


def bbox_overlaps(boxes: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
    """lib/fpn/box_intersections_cpu bbox_overlaps on device (float64, +1 convention) -> f64[n,k]."""
    _need_cuda(boxes, query)
    b = boxes.contiguous().to(torch.float64).reshape(-1, 4)
    q = query.contiguous().to(torch.float64).reshape(-1, 4)
    out = torch.empty(b.shape[0], q.shape[0], device=b.device, dtype=torch.float64)
    _C.check(_C.lib().nlv_bbox_overlaps_f64(_ptr(b), b.shape[0], _ptr(q), q.shape[0], _ptr(out), _stream()),
             "bbox_overlaps")
    return out
