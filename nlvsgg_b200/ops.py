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
         force_simt: bool = False, gate: torch.Tensor | None = None, drop=None, gate_scale: float = 1.0) -> torch.Tensor:
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
    if gate is not None:
        assert tuple(gate.shape) == (m, n) and gate.stride(1) == 1
        g.gate, g.ldg, g.gate_dtype = gate.data_ptr(), gate.stride(0), _dt(gate)
    else:
        g.gate, g.ldg, g.gate_dtype = None, 0, 0
    g.gate_scale = gate_scale
    if drop is not None:
        g.drop = drop
    if PROFILE is None:
        _C.check(_C.lib().nlv_gemm(ctypes.byref(g), _stream()), "gemm")
        return out
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _C.check(_C.lib().nlv_gemm(ctypes.byref(g), _stream()), "gemm")
    e.record()
    tag = "nlv_gemm[%s %s%s m=%d n=%d k=%d]" % ("tc" if (a.dtype == torch.bfloat16 and not force_simt) else "simt",
                                                 "KM"[a_major == MAJOR_MN], "KM"[b_major == MAJOR_MN], m, n, k)
    PROFILE.setdefault(tag, []).append((s, e))
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


# ----------------------------------------------------------------------------------------------
# thin wrappers for the remaining entry points (one C call each)
# ----------------------------------------------------------------------------------------------
_F = ctypes.c_float
_LL = ctypes.c_longlong


PROFILE = None  # set to {} to collect (start, end) CUDA events per entry point; see profile_summary()


def _call(name, *args):
    if PROFILE is None:
        _C.check(getattr(_C.lib(), name)(*args, _stream()), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _C.check(getattr(_C.lib(), name)(*args, _stream()), name)
    e.record()
    PROFILE.setdefault(name, []).append((s, e))


def profile_summary():
    """{entry point: (calls, total ms)} from the events collected while PROFILE was a dict (synchronises)."""
    torch.cuda.synchronize()
    return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in (PROFILE or {}).items()}


def convert(src: torch.Tensor, dtype: torch.dtype, out: torch.Tensor | None = None) -> torch.Tensor:
    assert src.dim() == 2 and src.stride(1) == 1
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=dtype)
    _call("nlv_convert", _ptr(src), _dt(src), src.stride(0), _ptr(out), _dt(out), out.stride(0), _LL(src.shape[0]),
          src.shape[1])
    return out


def split3(src: torch.Tensor, block_dim: int, pattern: int) -> torch.Tensor:
    """fp32 [r,c] -> bf16 [r,3c] (block_dim=1) or [3r,c] (block_dim=0); row stride padded to a multiple of 8."""
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    r, c = src.shape
    if block_dim == 1:
        ld = (3 * c + 7) // 8 * 8
        out = torch.empty(r, ld, device=src.device, dtype=torch.bfloat16)[:, :3 * c]
    else:
        ld = (c + 7) // 8 * 8
        out = torch.empty(3 * r, ld, device=src.device, dtype=torch.bfloat16)[:, :c]
    _call("nlv_split3", _ptr(src), src.stride(0), _LL(r), c, _ptr(out), out.stride(0), block_dim, pattern)
    return out


def nchw_to_rows(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    r, c, h, w = x.shape
    x = x.contiguous()
    out = torch.empty(r * h * w, c, device=x.device, dtype=dtype)
    step = 32768
    for s in range(0, r, step):  # grid.y limit
        n = min(step, r - s)
        _call("nlv_nchw_to_rows", _ptr(x[s:s + n]), _dt(x), n, c, h * w, _ptr(out[s * h * w:(s + n) * h * w]), _dt(out))
    return out


def im2col_mask(masks: torch.Tensor, dtype: torch.dtype, ld: int = 104) -> torch.Tensor:
    r = masks.shape[0]
    out = torch.empty(r * 196, ld, device=masks.device, dtype=dtype)
    _call("nlv_im2col_mask", _ptr(masks.contiguous()), r, _ptr(out), _dt(out), ld)
    return out


def im2col_3x3(x: torch.Tensor, r: int, h: int, w: int, c: int, dtype: torch.dtype) -> torch.Tensor:
    out = torch.empty(r * h * w, c * 9, device=x.device, dtype=dtype)
    _call("nlv_im2col_3x3", _ptr(x), _dt(x), r, h, w, c, _ptr(out), _dt(out))
    return out


def col2im_3x3(dcol: torch.Tensor, r: int, h: int, w: int, c: int) -> torch.Tensor:
    out = torch.empty(r * h * w, c, device=dcol.device, dtype=torch.float32)
    _call("nlv_col2im_3x3", _ptr(dcol), _dt(dcol), r, h, w, c, _ptr(out))
    return out


def maxpool_fwd(x: torch.Tensor, r: int, c: int, dtype: torch.dtype):
    y = torch.empty(r * 49, c, device=x.device, dtype=dtype)
    arg = torch.empty(r * 49, c, device=x.device, dtype=torch.uint8)
    _call("nlv_maxpool_fwd", _ptr(x), _dt(x), r, c, _ptr(y), _dt(y), _ptr(arg))
    return y, arg


def maxpool_bwd(dy: torch.Tensor, arg: torch.Tensor, r: int, c: int, out_dtype=torch.float32) -> torch.Tensor:
    assert dy.dtype == torch.float32
    dx = torch.empty(r * 196, c, device=dy.device, dtype=out_dtype)
    _call("nlv_maxpool_bwd", _ptr(dy), _ptr(arg), r, c, _ptr(dx), _dt(dx))
    return dx


def gather_rows(src, idx, n_out, out_dtype=None, add=None, add_idx=None, out2_dtype=None, want_out=True):
    cols = src.shape[1]
    out = torch.empty(n_out, cols, device=src.device, dtype=out_dtype or src.dtype) if want_out else None
    out2 = torch.empty(n_out, cols, device=src.device, dtype=out2_dtype) if out2_dtype is not None else None
    _call("nlv_gather_rows", _ptr(src), _dt(src), src.stride(0), _ptr(idx), _ptr(add), _ptr(add_idx),
          add.stride(0) if add is not None else 0, _LL(n_out), cols,
          _ptr(out), _dt(out) if out is not None else 0, cols, _ptr(out2), _dt(out2) if out2 is not None else 0, cols)
    return out, out2


def gather_sum_rows(src, idx, fan, n_out, out=None, accumulate=False):
    cols = src.shape[1]
    if out is None:
        out = torch.empty(n_out, cols, device=src.device, dtype=torch.float32)
    _call("nlv_gather_sum_rows", _ptr(src), src.stride(0), _ptr(idx), fan, _LL(n_out), cols, _ptr(out), out.stride(0),
          1 if accumulate else 0)
    return out


def assemble_tokens(fo, pair_idx, labels, e1, e2, rel):
    _call("nlv_assemble_tokens", _ptr(fo), _ptr(pair_idx), _ptr(labels), _ptr(e1), _ptr(e2), _LL(pair_idx.shape[0]), _ptr(rel))


def assemble_tokens_bwd(drel, pair_idx, labels, dfo, de1, de2):
    _call("nlv_assemble_tokens_bwd", _ptr(drel), _ptr(pair_idx), _ptr(labels), _LL(pair_idx.shape[0]), _ptr(dfo),
          _ptr(de1), _ptr(de2))


def center_size(boxes: torch.Tensor) -> torch.Tensor:
    out = torch.empty(boxes.shape[0], 4, device=boxes.device, dtype=torch.float32)
    _call("nlv_center_size", _ptr(boxes), _LL(boxes.shape[0]), _ptr(out))
    return out


def colsum(x: torch.Tensor, row_class=None, n_class: int = 1, out=None) -> torch.Tensor:
    rows, cols = x.shape
    if out is None:
        out = torch.zeros(n_class, cols, device=x.device, dtype=torch.float32)
    _call("nlv_colsum", _ptr(x), _dt(x), x.stride(0), _LL(rows), cols, _ptr(row_class), n_class, _ptr(out))
    return out


def relu_mask(x: torch.Tensor, gate: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    rows, cols = x.shape
    y = torch.empty(rows, cols, device=x.device, dtype=out_dtype)
    _call("nlv_relu_mask", _ptr(x), _dt(x), x.stride(0), _ptr(gate), _dt(gate), gate.stride(0), _LL(rows), cols,
          _ptr(y), _dt(y), cols)
    return y


def add(a: torch.Tensor, b: torch.Tensor, out=None) -> torch.Tensor:
    if out is None:
        out = torch.empty_like(a)
    _call("nlv_add", _ptr(a), _ptr(b), _LL(a.numel()), _ptr(out))
    return out


def layernorm_fwd(x, w, b, eps=1e-5, y2_dtype=None, want_y=True):
    rows, cols = x.shape
    y = torch.empty_like(x) if want_y else None
    y2 = torch.empty(rows, cols, device=x.device, dtype=y2_dtype) if y2_dtype is not None else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("nlv_layernorm_fwd", _ptr(x), _LL(rows), cols, _ptr(w), _ptr(b), _F(eps), _ptr(y), _ptr(y2),
          _dt(y2) if y2 is not None else 0, _ptr(mean), _ptr(rstd))
    return y, y2, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, w, dx2_dtype=None, drop=None):
    rows, cols = x.shape
    dx = torch.empty_like(x)
    dx2 = torch.empty(rows, cols, device=x.device, dtype=dx2_dtype) if dx2_dtype is not None else None
    dw = torch.zeros(cols, device=x.device, dtype=torch.float32)
    db = torch.zeros(cols, device=x.device, dtype=torch.float32)
    _call("nlv_layernorm_bwd_drop", _ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(w), _LL(rows), cols, _ptr(dx), _ptr(dx2),
          _dt(dx2) if dx2 is not None else 0, _ptr(dw), _ptr(db), ctypes.byref(drop) if drop is not None else None)
    return dx, dx2, dw, db


def layernorm_bwd_fused(dy, x, mean, rstd, w, dx2_dtype=None, drop=None, want_dprev=True):
    """One-pass backward: -> (dx, dx2, dw, db, dprev) with dprev = column sums of the dx2 values (bias gradient of the Linear in
    front of the residual sum)."""
    rows, cols = x.shape
    dx = torch.empty_like(x)
    dx2 = torch.empty(rows, cols, device=x.device, dtype=dx2_dtype) if dx2_dtype is not None else None
    dw = torch.zeros(cols, device=x.device, dtype=torch.float32)
    db = torch.zeros(cols, device=x.device, dtype=torch.float32)
    dprev = torch.zeros(cols, device=x.device, dtype=torch.float32) if want_dprev else None
    _call("nlv_layernorm_bwd_fused", _ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(w), _LL(rows), cols, _ptr(dx), _ptr(dx2),
          _dt(dx2) if dx2 is not None else 0, _ptr(dw), _ptr(db), _ptr(dprev), ctypes.byref(drop) if drop is not None else None)
    return dx, dx2, dw, db, dprev


def dropout_apply(src, drop, out_dtype=None, out=None):
    rows, cols = src.shape
    if out is None:
        out = torch.empty(rows, cols, device=src.device, dtype=out_dtype or src.dtype)
    _call("nlv_dropout_apply", _ptr(src), _dt(src), src.stride(0), _ptr(out), _dt(out), out.stride(0), _LL(rows), cols, ctypes.byref(drop))
    return out


def dropout_mask(rows, cols, drop, device="cuda"):
    out = torch.empty(rows, cols, device=device, dtype=torch.uint8)
    _call("nlv_dropout_mask", _LL(rows), cols, ctypes.byref(drop), _ptr(out))
    return out


def dropout_mask_attn(rows, heads, nkeys, drop, device="cuda"):
    out = torch.empty(rows, heads, nkeys, device=device, dtype=torch.uint8)
    _call("nlv_dropout_mask_attn", _LL(rows), heads, nkeys, ctypes.byref(drop), _ptr(out))
    return out


def bn_stats(x, seg, nseg, c, momentum, running_mean, running_var):
    rows = x.shape[0]
    ws = torch.empty(nseg * 2 * c, device=x.device, dtype=torch.float64)
    mean = torch.empty(nseg, c, device=x.device, dtype=torch.float32)
    var = torch.empty(nseg, c, device=x.device, dtype=torch.float32)
    _call("nlv_bn_stats", _ptr(x), _dt(x), x.stride(0), _ptr(seg), nseg, _LL(rows), c, _F(momentum), _ptr(ws), _ptr(mean),
          _ptr(var), _ptr(running_mean), _ptr(running_var))
    return mean, var


def bn_apply(x, row_seg, mean, var, w, b, relu, out=None, out_dtype=None, out2_dtype=None, eps=1e-5, row_div=1):
    rows, c = x.shape
    if out is None:
        out = torch.empty(rows, c, device=x.device, dtype=out_dtype or torch.float32)
    out2 = torch.empty(rows, c, device=x.device, dtype=out2_dtype) if out2_dtype is not None else None
    _call("nlv_bn_apply", _ptr(x), _dt(x), x.stride(0), _ptr(row_seg), row_div, _ptr(mean), _ptr(var), _ptr(w), _ptr(b), _F(eps),
          1 if relu else 0, _LL(rows), c, _ptr(out), _dt(out), out.stride(0), _ptr(out2),
          _dt(out2) if out2 is not None else 0, c)
    return out, out2


def bn_bwd(dy, x, yout, seg, row_seg, nseg, mean, var, w, use_batch_stats, dx_dtype=torch.float32, eps=1e-5, gate_by_x=False, row_div=1):
    rows, c = x.shape
    ws = torch.empty(nseg * 2 * c, device=x.device, dtype=torch.float64)
    dx = torch.empty(rows, c, device=x.device, dtype=dx_dtype)
    dw = torch.zeros(c, device=x.device, dtype=torch.float32)
    db = torch.zeros(c, device=x.device, dtype=torch.float32)
    _call("nlv_bn_bwd", _ptr(dy), _dt(dy), dy.stride(0), _ptr(x), _dt(x), x.stride(0), _ptr(yout),
          _dt(yout) if yout is not None else 0, yout.stride(0) if yout is not None else 0, _ptr(seg), _ptr(row_seg), row_div, nseg,
          _ptr(mean), _ptr(var), _ptr(w), _F(eps), 1 if use_batch_stats else 0, 1 if gate_by_x else 0, _LL(rows), c, _ptr(ws), _ptr(dx),
          _dt(dx), c,
          _ptr(dw), _ptr(db))
    return dx, dw, db


def attn_fwd(q, k, v, hd, heads, work, n_work, out_dtype, want_lse=True, drop=None):
    rows = q.shape[0]
    o = torch.empty(rows, hd * heads, device=q.device, dtype=out_dtype)
    lse = torch.empty(rows * heads, device=q.device, dtype=torch.float32) if want_lse else None
    _call("nlv_attn_fwd_drop", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _dt(q), hd, heads,
          _F(1.0 / (hd ** 0.5)), _ptr(work), n_work, _ptr(o), o.stride(0), _dt(o), _ptr(lse), ctypes.byref(drop) if drop is not None else None)
    return o, lse


def attn_bwd(q, k, v, o, dout, lse, hd, heads, work, n_work, dq, dk, dv, drop=None):
    rows = q.shape[0]
    delta = torch.empty(rows * heads, device=q.device, dtype=torch.float32)
    assert dq.dtype == dk.dtype == dv.dtype
    _call("nlv_attn_bwd_drop", _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _dt(q), hd, heads,
          _F(1.0 / (hd ** 0.5)), _ptr(work), n_work, _ptr(o), o.stride(0), _dt(o), _ptr(dout), dout.stride(0), _dt(dout),
          _ptr(lse), _ptr(delta), _ptr(dq), dq.stride(0), _ptr(dk), dk.stride(0), _ptr(dv), dv.stride(0), _dt(dq),
          ctypes.byref(drop) if drop is not None else None)


def heads_activation(logits):
    r = logits.shape[0]
    att = torch.empty(r, 3, device=logits.device, dtype=torch.float32)
    spa = torch.empty(r, 6, device=logits.device, dtype=torch.float32)
    con = torch.empty(r, 17, device=logits.device, dtype=torch.float32)
    _call("nlv_heads_activation", _ptr(logits), _LL(r), _ptr(att), _ptr(spa), _ptr(con))
    return att, spa, con


def ce_loss(logits, c, labels, row_weight, loss, dlogits):
    _call("nlv_ce_loss", _ptr(logits), logits.stride(0), c, _ptr(labels), _ptr(row_weight), _LL(logits.shape[0]), _ptr(loss),
          _ptr(dlogits), dlogits.stride(0) if dlogits is not None else 0)


def bce_sigmoid_loss(logits, c, bits, row_weight, loss, dlogits):
    _call("nlv_bce_sigmoid_loss", _ptr(logits), logits.stride(0), c, _ptr(bits), _ptr(row_weight), _LL(logits.shape[0]),
          _ptr(loss), _ptr(dlogits), dlogits.stride(0) if dlogits is not None else 0)


def zero_(t: torch.Tensor) -> torch.Tensor:
    assert t.is_contiguous()
    _call("nlv_zero_bytes", _ptr(t), _LL(t.numel() * t.element_size()))
    return t


def heads_activation_bwd(datt, dspa, dcon, spa, con, r):
    d26 = torch.empty(r, 26, device=spa.device, dtype=torch.float32)
    c = lambda t: t.contiguous() if t is not None else None
    datt, dspa, dcon = c(datt), c(dspa), c(dcon)
    _call("nlv_heads_activation_bwd", _ptr(datt), _ptr(dspa), _ptr(dcon), _ptr(spa), _ptr(con), _LL(r), _ptr(d26))
    return d26


def sumsq(x, out):
    _call("nlv_sumsq", _ptr(x), _LL(x.numel()), _ptr(out))


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, wd, step, total_sq=None, max_norm=0.0, p_bf16=None):
    _call("nlv_adamw_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), _LL(p.numel()), _F(lr), _F(beta1), _F(beta2), _F(eps), _F(wd),
          int(step), _ptr(total_sq), _F(max_norm), _ptr(p_bf16))


# ---- spatial-mask branch, first stage without the im2col matrix (csrc/maskconv.cu; bf16 path) ----
def mask_conv1_fwd(masks, w, bias, pair_video, seg196, nv, momentum=0.01, running_mean=None, running_var=None, want_stats=True):
    """relu(conv7x7/2(masks) + bias) as bf16 [r*196,128] and, for training, its per-video BatchNorm statistics."""
    _need_cuda(masks, w, bias)
    r = masks.shape[0]
    out = torch.empty(r * 196, 128, device=masks.device, dtype=torch.bfloat16)
    ws = torch.empty(nv * 2 * 128, device=masks.device, dtype=torch.float64)
    mean = torch.empty(nv, 128, device=masks.device, dtype=torch.float32) if want_stats else None
    var = torch.empty(nv, 128, device=masks.device, dtype=torch.float32) if want_stats else None
    _call("nlv_mask_conv1_fwd", _ptr(masks.contiguous()), _ptr(w.contiguous()), _ptr(bias), _LL(r), _ptr(pair_video), _ptr(seg196), nv,
          _ptr(out), _F(momentum), _ptr(ws), _ptr(mean), _ptr(var), _ptr(running_mean), _ptr(running_var))
    return out, mean, var


def bn_apply_maxpool(x, pair_video, mean, var, w, b, r, eps=1e-5, want_xmax=False):
    y = torch.empty(r * 49, 128, device=x.device, dtype=torch.bfloat16)
    arg = torch.empty(r * 49, 128, device=x.device, dtype=torch.uint8)
    xmax = torch.empty(r * 49, 128, device=x.device, dtype=torch.bfloat16) if want_xmax else None
    _call("nlv_bn_apply_maxpool", _ptr(x), _ptr(pair_video), _ptr(mean), _ptr(var), _ptr(w), _ptr(b), _F(eps), _LL(r), _ptr(y), _ptr(arg),
          _ptr(xmax))
    return (y, arg, xmax) if want_xmax else (y, arg)


def pool_bn_bwd(dp, arg, x, xmax, pair_video, seg196, seg49, nv, mean, var, w, r, use_batch_stats=True, eps=1e-5):
    """dx (bf16, at the conv output), BatchNorm dw / db and the column sums of dx from the pooled gradient."""
    ws = torch.empty(nv * 2 * 128, device=x.device, dtype=torch.float64)
    dx = torch.empty(r * 196, 128, device=x.device, dtype=torch.bfloat16)
    dw, db, cs = (torch.zeros(128, device=x.device, dtype=torch.float32) for _ in range(3))
    _call("nlv_pool_bn_bwd", _ptr(dp), _ptr(arg), _ptr(x), _ptr(xmax), _ptr(pair_video), _ptr(seg196), _ptr(seg49), nv, _ptr(mean), _ptr(var),
          _ptr(w), _F(eps), 1 if use_batch_stats else 0, _LL(r), _ptr(ws), _ptr(dx), _ptr(dw), _ptr(db), _ptr(cs))
    return dx, dw, db, cs


def mask_conv1_dw(dy, masks):
    r = masks.shape[0]
    f = _C.lib().nlv_mask_conv1_dw_ws_floats
    f.restype = ctypes.c_longlong
    ws = torch.empty(int(f()), device=dy.device, dtype=torch.float32)
    dw = torch.empty(128, 98, device=dy.device, dtype=torch.float32)
    _call("nlv_mask_conv1_dw", _ptr(dy), _ptr(masks.contiguous()), _LL(r), _ptr(ws), _ptr(dw))
    return dw


def conv3x3_dgrad(dy, wt, r, c_out, c_in, out_dtype=torch.float32):
    """dx [r*49, c_in] of the 3x3 / pad 1 conv from dy bf16 [r*49, c_out] and wt bf16 [c_in, 9*c_out] (implicit GEMM)."""
    dx = torch.empty(r * 49, c_in, device=dy.device, dtype=out_dtype)
    _call("nlv_conv3x3_dgrad", _ptr(dy), _LL(r), c_out, _ptr(wt), c_in, _ptr(dx), _dt(dx))
    return dx
