"""Forward + backward of the STTran / DSG-DETR relation models as explicit sequences of C-ABI kernel
launches (no torch arithmetic on the data path; torch allocates buffers and owns the parameters).

Layout in HBM
  * boxes of all videos of a batch are concatenated ([N,*]); pairs likewise ([R,*], frame-sorted per video);
  * the spatial encoder runs on the R pair tokens, segments = frames;
  * the temporal decoder runs on a *window stream* of Mg = sum_w (n_j + n_{j+1}) rows: window j of a video is the
    contiguous row range of frames {j, j+1}, copied once; segments = windows (lib/transformer_wk.py:163-171);
  * DSG-DETR's temporal encoder runs on a class-sorted permutation of the R tokens, segments = object classes.
Precision modes
  * 'bf16'   : GEMM operands bf16 on tcgen05, fp32 accumulation, fp32 residual stream / norms / softmax
  * 'bf16x3' : fp32 tensors, each operand split hi/lo into 3 bf16 K-blocks -> fp32-faithful on the same tcgen05 kernel
  * 'fp32'   : exact-fp32 SIMT GEMM (debug / tiny problems)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from ._C import MAJOR_K as K_, MAJOR_MN as MN_

D_MODEL, N_HEAD, HEAD_DIM, D_FF = 1936, 8, 242, 2048
F32, BF16 = torch.float32, torch.bfloat16


# ================================================================================================
# kernel context: precision mode, operand caches, GEMM dispatch
# ================================================================================================
class Kernels:
    def __init__(self, precision: str = "bf16"):
        assert precision in ("bf16", "bf16x3", "fp32")
        self.precision = precision
        self.AD = BF16 if precision == "bf16" else F32   # dtype of GEMM-operand activations
        self._wcache: Dict[tuple, tuple] = {}
        self.mirror: Dict[str, torch.Tensor] = {}

    # ---- activations -------------------------------------------------------------------------
    def opnd(self, x: torch.Tensor) -> torch.Tensor:
        """GEMM-operand form of an activation matrix."""
        if self.precision == "bf16" and x.dtype != BF16:
            return ops.convert(x, BF16)
        return x

    # ---- weights -----------------------------------------------------------------------------
    def weight(self, key: str, srcs, make=None, f32: bool = False) -> torch.Tensor:
        """Cached operand form of one or several parameters (re-made when any of them is updated in place).
        `make(*tensors)` builds the matrix (concatenation / permutation / reshape); f32 keeps it in fp32."""
        m = self.mirror.get(key)
        if m is not None:      # operand copy kept current by the optimiser kernel (trainer.py)
            return m
        if torch.is_tensor(srcs):
            srcs = (srcs,)
        ver = tuple((t.data_ptr(), t._version) for t in srcs)
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        det = [t.detach() for t in srcs]
        w = make(*det) if make is not None else det[0]
        if w.dim() == 2 and self.precision == "bf16" and not f32:
            if w.shape[1] % 8 != 0:  # TMA needs 16-byte rows
                pad = torch.zeros(w.shape[0], (w.shape[1] + 7) // 8 * 8, device=w.device, dtype=F32)
                pad[:, :w.shape[1]] = w
                w = ops.convert(pad, BF16)[:, :w.shape[1]]
            else:
                w = ops.convert(w.contiguous(), BF16)
        else:
            w = w.contiguous()
        self._wcache[key] = (ver, w)
        return w

    # ---- GEMM --------------------------------------------------------------------------------
    def mm(self, a, b, *, a_major=K_, b_major=K_, out=None, out_dtype=None, bias=None, residual=None, relu=False,
           exact=False, gate=None):
        m = a.shape[0] if a_major == K_ else a.shape[1]
        n = b.shape[0] if b_major == K_ else b.shape[1]
        if out is None:
            od = out_dtype or F32
            if self.precision != "bf16":
                od = F32
            out = torch.empty(m, n, device=a.device, dtype=od)
        kdim = a.shape[1] if a_major == K_ else a.shape[0]
        if exact and self.precision == "bf16" and a.dtype == F32 and b.dtype == F32 and kdim >= 256:
            # fp32-grade product on the tensor cores (three bf16 terms per operand, error ~2^-17) instead of the SIMT kernel
            a3 = ops.split3(a, 1 if a_major == K_ else 0, 0)
            b3 = ops.split3(b, 1 if b_major == K_ else 0, 1)
            return ops.gemm(a3, b3, out, a_major=a_major, b_major=b_major, bias=bias, residual=residual, relu=relu, gate=gate)
        if exact or self.precision == "fp32":
            if a.dtype != F32:
                a = ops.convert(a, F32)
            if b.dtype != F32:
                b = ops.convert(b, F32)
            return ops.gemm(a, b, out, a_major=a_major, b_major=b_major, bias=bias, residual=residual, relu=relu, gate=gate)
        if self.precision == "bf16x3":
            a3 = ops.split3(a, 1 if a_major == K_ else 0, 0)
            b3 = ops.split3(b, 1 if b_major == K_ else 0, 1)
            return ops.gemm(a3, b3, out, a_major=a_major, b_major=b_major, bias=bias, residual=residual, relu=relu, gate=gate)
        a, b = self._tma_ready(a), self._tma_ready(b)
        return ops.gemm(a, b, out, a_major=a_major, b_major=b_major, bias=bias, residual=residual, relu=relu, gate=gate)

    def _tma_ready(self, x):
        if x.dtype != BF16:
            x = ops.convert(x, BF16)
        if x.stride(0) % 8 != 0 or x.data_ptr() % 16 != 0:
            ld = (x.shape[1] + 7) // 8 * 8
            buf = torch.zeros(x.shape[0], ld, device=x.device, dtype=BF16)
            ops.convert(x, BF16, out=buf[:, :x.shape[1]])
            x = buf[:, :x.shape[1]]
        return x


from .plan import Plan  # noqa: E402,F401  (host-side descriptors live in plan.py)


# ================================================================================================
# transformer layers
# ================================================================================================
def _lin_grads(k: Kernels, dy_op, x_op, dy_for_bias, grads, wname, bname):
    """dW = dY^T X (both operands MN-major), db = column sums.  When `grads` exposes flat-buffer views (trainer.GradSink)
    the GEMM and the column sum write straight into them; otherwise fresh tensors are stored in the mapping."""
    views = getattr(grads, "views", None)
    if views is not None and wname in views and views[wname].dim() == 2:
        k.mm(dy_op, x_op, a_major=MN_, b_major=MN_, out=views[wname])
        bv = views[bname].view(1, -1)
        bv.zero_()
        ops.colsum(dy_for_bias, out=bv)
        grads.seen.update((wname, bname))
        return
    grads[wname] = k.mm(dy_op, x_op, a_major=MN_, b_major=MN_)
    grads[bname] = ops.colsum(dy_for_bias).reshape(-1)


def encoder_fwd(k: Kernels, P: dict, pre: str, attn: str, x, xop, work, n_work, want_ctx: bool, out_op: bool):
    """Post-norm encoder layer (lib/transformer.py:20-30).  x: fp32 [M,d] residual stream, xop: its operand form."""
    w = lambda n: k.weight(pre + n, P[pre + n])
    qkv = k.mm(xop, w(f"{attn}.in_proj_weight"), bias=P[pre + f"{attn}.in_proj_bias"], out_dtype=k.AD)
    d = D_MODEL
    o, lse = ops.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], HEAD_DIM, N_HEAD, work, n_work, k.AD, want_ctx)
    y1 = k.mm(o, w(f"{attn}.out_proj.weight"), bias=P[pre + f"{attn}.out_proj.bias"], residual=x)
    x1, x1op, m1, r1 = ops.layernorm_fwd(y1, P[pre + "norm1.weight"], P[pre + "norm1.bias"],
                                         y2_dtype=BF16 if k.AD == BF16 else None)
    if x1op is None:
        x1op = x1
    h = k.mm(x1op, w("linear1.weight"), bias=P[pre + "linear1.bias"], relu=True, out_dtype=k.AD)
    y2 = k.mm(h, w("linear2.weight"), bias=P[pre + "linear2.bias"], residual=x1)
    x2, x2op, m2, r2 = ops.layernorm_fwd(y2, P[pre + "norm2.weight"], P[pre + "norm2.bias"],
                                         y2_dtype=BF16 if (k.AD == BF16 and out_op) else None)
    ctx = dict(xop=xop, qkv=qkv, o=o, lse=lse, y1=y1, m1=m1, r1=r1, x1op=x1op, h=h, y2=y2, m2=m2, r2=r2) if want_ctx else None
    return x2, (x2op if x2op is not None else x2), ctx


def encoder_bwd(k: Kernels, P: dict, pre: str, attn: str, c: dict, dx2, work, n_work, grads: dict, need_dx: bool = True):
    w = lambda n: k.weight(pre + n, P[pre + n])
    d = D_MODEL
    opd = BF16 if k.AD == BF16 else None
    dy2, dy2op, dw, db = ops.layernorm_bwd(dx2, c["y2"], c["m2"], c["r2"], P[pre + "norm2.weight"], dx2_dtype=opd)
    grads[pre + "norm2.weight"], grads[pre + "norm2.bias"] = dw, db
    if dy2op is None:
        dy2op = dy2
    _lin_grads(k, dy2op, c["h"], dy2, grads, pre + "linear2.weight", pre + "linear2.bias")
    dh = k.mm(dy2op, w("linear2.weight"), b_major=MN_, out_dtype=k.AD, gate=c["h"])       # ReLU backward fused
    _lin_grads(k, dh, c["x1op"], dh, grads, pre + "linear1.weight", pre + "linear1.bias")
    dx1 = k.mm(dh, w("linear1.weight"), b_major=MN_, residual=dy2)
    dy1, dy1op, dw, db = ops.layernorm_bwd(dx1, c["y1"], c["m1"], c["r1"], P[pre + "norm1.weight"], dx2_dtype=opd)
    grads[pre + "norm1.weight"], grads[pre + "norm1.bias"] = dw, db
    if dy1op is None:
        dy1op = dy1
    _lin_grads(k, dy1op, c["o"], dy1, grads, pre + f"{attn}.out_proj.weight", pre + f"{attn}.out_proj.bias")
    do = k.mm(dy1op, w(f"{attn}.out_proj.weight"), b_major=MN_, out_dtype=k.AD)
    qkv = c["qkv"]
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], c["o"], do, c["lse"], HEAD_DIM, N_HEAD, work, n_work,
                 dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
    _lin_grads(k, dqkv, c["xop"], dqkv, grads, pre + f"{attn}.in_proj_weight", pre + f"{attn}.in_proj_bias")
    if not need_dx:
        return None
    return k.mm(dqkv, w(f"{attn}.in_proj_weight"), b_major=MN_, residual=dy1)


def decoder_fwd(k: Kernels, P: dict, pre: str, x, xop, xpop, work, n_work, want_ctx: bool):
    """Temporal decoder layer (lib/transformer.py:49-58): q = k = x + pos, v = x; LayerNorm after the attention
    residual, plain residual after the FFN.  xpop = operand form of x + pos."""
    d = D_MODEL
    win = k.weight(pre + "multihead2.in_proj_weight", P[pre + "multihead2.in_proj_weight"])
    bin_ = P[pre + "multihead2.in_proj_bias"]
    w = lambda n: k.weight(pre + n, P[pre + n])
    M = x.shape[0]
    qkv = torch.empty(M, 3 * d, device=x.device, dtype=k.AD)
    k.mm(xpop, win[:2 * d], bias=bin_[:2 * d], out=qkv[:, :2 * d])
    k.mm(xop, win[2 * d:], bias=bin_[2 * d:], out=qkv[:, 2 * d:])
    o, lse = ops.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], HEAD_DIM, N_HEAD, work, n_work, k.AD, want_ctx)
    y = k.mm(o, w("multihead2.out_proj.weight"), bias=P[pre + "multihead2.out_proj.bias"], residual=x)
    t, top, m3, r3 = ops.layernorm_fwd(y, P[pre + "norm3.weight"], P[pre + "norm3.bias"],
                                       y2_dtype=BF16 if k.AD == BF16 else None)
    if top is None:
        top = t
    h = k.mm(top, w("linear1.weight"), bias=P[pre + "linear1.bias"], relu=True, out_dtype=k.AD)
    out = k.mm(h, w("linear2.weight"), bias=P[pre + "linear2.bias"], residual=t)
    ctx = dict(xop=xop, xpop=xpop, qkv=qkv, o=o, lse=lse, y=y, m3=m3, r3=r3, top=top, h=h) if want_ctx else None
    return out, ctx


def decoder_bwd(k: Kernels, P: dict, pre: str, c: dict, dout, slot, work, n_work, grads: dict):
    """Returns (dx, dpos[2,d])."""
    d = D_MODEL
    win = k.weight(pre + "multihead2.in_proj_weight", P[pre + "multihead2.in_proj_weight"])
    w = lambda n: k.weight(pre + n, P[pre + n])
    doutop = k.opnd(dout)
    _lin_grads(k, doutop, c["h"], dout, grads, pre + "linear2.weight", pre + "linear2.bias")
    dh = k.mm(doutop, w("linear2.weight"), b_major=MN_, out_dtype=k.AD, gate=c["h"])      # ReLU backward fused
    _lin_grads(k, dh, c["top"], dh, grads, pre + "linear1.weight", pre + "linear1.bias")
    dt = k.mm(dh, w("linear1.weight"), b_major=MN_, residual=dout)
    dy, dyop, dw, db = ops.layernorm_bwd(dt, c["y"], c["m3"], c["r3"], P[pre + "norm3.weight"],
                                         dx2_dtype=BF16 if k.AD == BF16 else None)
    grads[pre + "norm3.weight"], grads[pre + "norm3.bias"] = dw, db
    if dyop is None:
        dyop = dy
    _lin_grads(k, dyop, c["o"], dy, grads, pre + "multihead2.out_proj.weight", pre + "multihead2.out_proj.bias")
    do = k.mm(dyop, w("multihead2.out_proj.weight"), b_major=MN_, out_dtype=k.AD)
    qkv = c["qkv"]
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], c["o"], do, c["lse"], HEAD_DIM, N_HEAD, work, n_work,
                 dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
    views = getattr(grads, "views", None)
    wname = pre + "multihead2.in_proj_weight"
    dwin = views[wname] if views is not None else torch.empty(3 * d, d, device=dout.device, dtype=F32)
    k.mm(dqkv[:, :2 * d], c["xpop"], a_major=MN_, b_major=MN_, out=dwin[:2 * d])
    k.mm(dqkv[:, 2 * d:], c["xop"], a_major=MN_, b_major=MN_, out=dwin[2 * d:])
    if views is not None:
        grads.seen.add(wname)
    else:
        grads[wname] = dwin
    grads[pre + "multihead2.in_proj_bias"] = ops.colsum(dqkv).reshape(-1)
    dxp = k.mm(dqkv[:, :2 * d], win[:2 * d], b_major=MN_)                 # gradient w.r.t. (x + pos)
    dpos = ops.colsum(dxp, row_class=slot, n_class=2)
    dx = k.mm(dqkv[:, 2 * d:], win[2 * d:], b_major=MN_, residual=dxp, out=dxp)
    ops.add(dx, dy, out=dx)
    return dx, dpos


# ================================================================================================
# object classifier + pair tokens
# ================================================================================================
def _bn_fwd(x, seg, row_seg, nseg, P, pre, momentum, training, relu, out=None, out_dtype=None, update_running=True):
    """Returns (y, mean, var).  Training: per-video batch statistics, running stats updated in video order."""
    c = x.shape[1]
    if training:
        mean, var = ops.bn_stats(x, seg, nseg, c, momentum, P[pre + ".running_mean"] if update_running else None,
                                 P[pre + ".running_var"] if update_running else None)
        rs = row_seg
    else:
        mean, var, rs = P[pre + ".running_mean"].reshape(1, c), P[pre + ".running_var"].reshape(1, c), None
    y, _ = ops.bn_apply(x, rs, mean, var, P[pre + ".weight"], P[pre + ".bias"], relu, out=out, out_dtype=out_dtype)
    return y, mean, var


def object_classifier_fwd(k: Kernels, P: dict, plan: Plan, features, distribution, boxes, training: bool, want_ctx: bool):
    """sgdet / is_wks branch of lib/sttran.py:173-184.  Returns logits [N,37] (+ the [N,2376] operand, reused by
    the pair-token stage for its feature columns)."""
    pre = "object_classifier."
    N = features.shape[0]
    objfeat = torch.empty(N, 2376, device=features.device, dtype=k.AD)
    ops.convert(features, k.AD, out=objfeat[:, :2048])
    k.mm(distribution, P[pre + "obj_embed.weight"], b_major=MN_, out=objfeat[:, 2048:2248], exact=True)
    cs = ops.center_size(boxes)
    pos_bn, mean0, var0 = _bn_fwd(cs, plan.box_seg, plan.box_row, plan.nv, P, pre + "pos_embed.0", 0.01 / 10.0, training, False)
    k.mm(pos_bn, P[pre + "pos_embed.1.weight"], bias=P[pre + "pos_embed.1.bias"], relu=True, out=objfeat[:, 2248:], exact=True)
    h1 = k.mm(objfeat, k.weight(pre + "decoder_lin.0.weight", P[pre + "decoder_lin.0.weight"]), bias=P[pre + "decoder_lin.0.bias"])
    h2, mean1, var1 = _bn_fwd(h1, plan.box_seg, plan.box_row, plan.nv, P, pre + "decoder_lin.1", 0.1, training, True)
    logits = k.mm(h2, P[pre + "decoder_lin.3.weight"], bias=P[pre + "decoder_lin.3.bias"], exact=True)
    ctx = dict(objfeat=objfeat, cs=cs, pos_bn=pos_bn, mean0=mean0, var0=var0, h1=h1, h2=h2, mean1=mean1, var1=var1,
               distribution=distribution, training=training) if want_ctx else None
    return logits, objfeat, ctx


def object_classifier_bwd(k: Kernels, P: dict, plan: Plan, c: dict, dlogits, grads: dict):
    pre = "object_classifier."
    tr = c["training"]
    grads[pre + "decoder_lin.3.weight"] = k.mm(dlogits, c["h2"], a_major=MN_, b_major=MN_, exact=True)
    grads[pre + "decoder_lin.3.bias"] = ops.colsum(dlogits).reshape(-1)
    dh2 = k.mm(dlogits, P[pre + "decoder_lin.3.weight"], b_major=MN_, exact=True)
    dh1, dw, db = ops.bn_bwd(dh2, c["h1"], c["h2"], plan.box_seg, plan.box_row, plan.nv, c["mean1"], c["var1"],
                             P[pre + "decoder_lin.1.weight"], tr, dx_dtype=k.AD)
    grads[pre + "decoder_lin.1.weight"], grads[pre + "decoder_lin.1.bias"] = dw, db
    objfeat = c["objfeat"]
    _lin_grads(k, dh1, objfeat, dh1, grads, pre + "decoder_lin.0.weight", pre + "decoder_lin.0.bias")
    w0 = k.weight(pre + "decoder_lin.0.weight", P[pre + "decoder_lin.0.weight"])
    dtail = k.mm(dh1, w0[:, 2048:], b_major=MN_)                                    # [N, 200 + 128] fp32
    grads[pre + "obj_embed.weight"] = k.mm(c["distribution"], dtail[:, :200], a_major=MN_, b_major=MN_, exact=True)
    dpos = ops.relu_mask(dtail[:, 200:], objfeat[:, 2248:], F32)
    grads[pre + "pos_embed.1.weight"] = k.mm(dpos, c["pos_bn"], a_major=MN_, b_major=MN_, exact=True)
    grads[pre + "pos_embed.1.bias"] = ops.colsum(dpos).reshape(-1)
    dposbn = k.mm(dpos, P[pre + "pos_embed.1.weight"], b_major=MN_, exact=True)
    _, dw, db = ops.bn_bwd(dposbn, c["cs"], None, plan.box_seg, plan.box_row, plan.nv, c["mean0"], c["var0"],
                           P[pre + "pos_embed.0.weight"], tr)
    grads[pre + "pos_embed.0.weight"], grads[pre + "pos_embed.0.bias"] = dw, db


def _perm_c4(w):      # conv.4.weight [256,128,3,3] -> [256, (ky,kx,c)] to match the channels-innermost im2col
    return w.permute(0, 2, 3, 1).reshape(256, 1152)


def _perm_vr(w):      # vr_fc.weight [512, c*49 + hw] -> [512, hw*256 + c] (rows of the NHWC union tensor)
    return w.view(512, 256, 49).permute(0, 2, 1).reshape(512, 12544)


def _unperm_vr(g):
    return g.view(512, 49, 256).permute(0, 2, 1).reshape(512, 12544).contiguous()


def pair_tokens_fwd(k: Kernels, P: dict, plan: Plan, feat_op, union_feat, spatial_masks, pair_idx, pred_labels,
                    training: bool, want_ctx: bool):
    """1936-d relation tokens (lib/sttran.py:381-399).  feat_op: [N,2048] operand view of the box features."""
    R = pair_idx.shape[0]
    dev = union_feat.device
    w_so = k.weight("subjobj.weight", (P["subj_fc.weight"], P["obj_fc.weight"]), lambda a, b: torch.cat((a, b), 0))
    b_so = k.weight("subjobj.bias", (P["subj_fc.bias"], P["obj_fc.bias"]), lambda a, b: torch.cat((a, b)), f32=True)
    fo = k.mm(feat_op, w_so, bias=b_so)                                              # [N,1024] fp32
    uf_op = ops.nchw_to_rows(union_feat, k.AD)                                       # [R*49, 2048]
    col1 = ops.im2col_mask(spatial_masks, k.AD, 104)                                 # [R*196, 104]
    w_c0 = k.weight("conv.0.weight", P["conv.0.weight"], lambda w: torch.nn.functional.pad(w.reshape(128, 98), (0, 6)))
    c1 = k.mm(col1, w_c0, bias=P["conv.0.bias"], relu=True, out_dtype=k.AD)          # conv7x7 s2 + ReLU
    b1, mean2, var2 = _bn_fwd(c1, plan.seg196, plan.row196, plan.nv, P, "conv.2", 0.01, training, False, out_dtype=k.AD)
    p1, arg = ops.maxpool_fwd(b1, R, 128, k.AD)
    col2 = ops.im2col_3x3(p1, R, 7, 7, 128, k.AD)                                    # [R*49, 1152]
    w_c4 = k.weight("conv.4.weight.taps", P["conv.4.weight"], _perm_c4)
    c2 = k.mm(col2, w_c4, bias=P["conv.4.bias"], relu=True, out_dtype=k.AD)
    b2, mean6, var6 = _bn_fwd(c2, plan.seg49, plan.row49, plan.nv, P, "conv.6", 0.01, training, False, out_dtype=k.AD)
    w_u = k.weight("union_func1.weight", P["union_func1.weight"], lambda w: w.reshape(256, 2048))
    vr_in = k.mm(uf_op, w_u, bias=P["union_func1.bias"], residual=b2, out_dtype=k.AD)  # [R*49,256] = [R,12544] (hw,c)
    rel = torch.empty(R, D_MODEL, device=dev, dtype=F32)
    w_vr = k.weight("vr_fc.weight", P["vr_fc.weight"], _perm_vr)
    k.mm(vr_in.view(R, 12544), w_vr, bias=P["vr_fc.bias"], out=rel[:, 1024:1536])
    ops.assemble_tokens(fo, pair_idx, pred_labels, P["obj_embed.weight"], P["obj_embed2.weight"], rel)
    ctx = dict(feat_op=feat_op, uf_op=uf_op, col1=col1, c1=c1, mean2=mean2, var2=var2, arg=arg, col2=col2, c2=c2,
               mean6=mean6, var6=var6, vr_in=vr_in, pair_idx=pair_idx, labels=pred_labels, training=training,
               n_boxes=feat_op.shape[0]) if want_ctx else None
    return rel, ctx


def pair_tokens_bwd(k: Kernels, P: dict, plan: Plan, c: dict, drel, grads: dict):
    R = drel.shape[0]
    dev = drel.device
    tr = c["training"]
    dfo = torch.zeros(c["n_boxes"], 1024, device=dev, dtype=F32)
    de1 = torch.zeros(37, 200, device=dev, dtype=F32)
    de2 = torch.zeros(37, 200, device=dev, dtype=F32)
    ops.assemble_tokens_bwd(drel, c["pair_idx"], c["labels"], dfo, de1, de2)
    grads["obj_embed.weight"], grads["obj_embed2.weight"] = de1, de2
    dvr = k.opnd(drel[:, 1024:1536]) if k.AD == BF16 else drel[:, 1024:1536]
    vr_in2d = c["vr_in"].view(R, 12544)
    grads["vr_fc.weight"] = _unperm_vr(k.mm(dvr, vr_in2d, a_major=MN_, b_major=MN_))
    grads["vr_fc.bias"] = ops.colsum(drel[:, 1024:1536]).reshape(-1)
    w_vr = k.weight("vr_fc.weight", P["vr_fc.weight"], _perm_vr)
    dvr_in = k.mm(dvr, w_vr, b_major=MN_, out_dtype=k.AD).view(R * 49, 256)          # activation dtype: feeds a GEMM, a column sum and BN backward
    dvr_op = k.opnd(dvr_in)
    grads["union_func1.weight"] = k.mm(dvr_op, c["uf_op"], a_major=MN_, b_major=MN_).view(256, 2048, 1, 1)
    grads["union_func1.bias"] = ops.colsum(dvr_in).reshape(-1)
    dc2, dw, db = ops.bn_bwd(dvr_in, c["c2"], None, plan.seg49, plan.row49, plan.nv, c["mean6"], c["var6"],
                             P["conv.6.weight"], tr, dx_dtype=k.AD, gate_by_x=True)     # BN backward + ReLU backward fused
    grads["conv.6.weight"], grads["conv.6.bias"] = dw, db
    grads["conv.4.weight"] = k.mm(dc2, c["col2"], a_major=MN_, b_major=MN_).view(256, 3, 3, 128).permute(0, 3, 1, 2).contiguous()
    grads["conv.4.bias"] = ops.colsum(dc2).reshape(-1)
    w_c4 = k.weight("conv.4.weight.taps", P["conv.4.weight"], _perm_c4)
    dcol2 = k.mm(dc2, w_c4, b_major=MN_, out_dtype=k.AD)
    dp1 = ops.col2im_3x3(dcol2, R, 7, 7, 128)
    db1 = ops.maxpool_bwd(dp1, c["arg"], R, 128, k.AD)
    dc1, dw, db = ops.bn_bwd(db1, c["c1"], None, plan.seg196, plan.row196, plan.nv, c["mean2"], c["var2"],
                             P["conv.2.weight"], tr, dx_dtype=k.AD, gate_by_x=True)
    grads["conv.2.weight"], grads["conv.2.bias"] = dw, db
    grads["conv.0.weight"] = k.mm(dc1, c["col1"], a_major=MN_, b_major=MN_)[:, :98].reshape(128, 2, 7, 7).contiguous()
    grads["conv.0.bias"] = ops.colsum(dc1).reshape(-1)
    dwso = k.mm(k.opnd(dfo), c["feat_op"], a_major=MN_, b_major=MN_)
    dbso = ops.colsum(dfo).reshape(-1)
    grads["subj_fc.weight"], grads["obj_fc.weight"] = dwso[:512].contiguous(), dwso[512:].contiguous()
    grads["subj_fc.bias"], grads["obj_fc.bias"] = dbso[:512].contiguous(), dbso[512:].contiguous()


# ================================================================================================
# whole models
# ================================================================================================
GT = "glocal_transformer."


def sttran_transformer_fwd(k: Kernels, P: dict, plan: Plan, rel, want_ctx: bool):
    """transformer_wk.forward (mode='latter') on the concatenated batch."""
    ctx = {}
    x, xop = rel, k.opnd(rel)
    n_enc = sum(1 for n in P if n.startswith(GT + "local_attention.layers.") and n.endswith("norm1.weight"))
    n_dec = sum(1 for n in P if n.startswith(GT + "global_attention.layers.") and n.endswith("norm3.weight"))
    ctx["enc"] = []
    for i in range(n_enc):
        x, xop, c = encoder_fwd(k, P, f"{GT}local_attention.layers.{i}.", "self_attn", x, xop, plan.local_work,
                                plan.n_local_work, want_ctx, out_op=False)
        ctx["enc"].append(c)
    local_out = x
    if plan.Mg == 0:
        return local_out, (ctx if want_ctx else None)
    pe = P[GT + "position_embedding.weight"]
    # window stream: g = local_out[stream_src]; operand copies of g and g + pos[slot]
    g, gop = ops.gather_rows(local_out, plan.stream_src, plan.Mg, out_dtype=F32, out2_dtype=BF16 if k.AD == BF16 else None)
    _, gpop = ops.gather_rows(local_out, plan.stream_src, plan.Mg, add=pe, add_idx=plan.stream_slot, out2_dtype=k.AD, want_out=False)
    if gop is None:
        gop = g
    ctx["dec"] = []
    for i in range(n_dec):
        if i > 0:
            gop = k.opnd(g)
            _, gpop = ops.gather_rows(g, None, plan.Mg, add=pe, add_idx=plan.stream_slot, out2_dtype=k.AD, want_out=False)
        g, c = decoder_fwd(k, P, f"{GT}global_attention.layers.{i}.", g, gop, gpop, plan.glob_work, plan.n_glob_work, want_ctx)
        ctx["dec"].append(c)
    out, _ = ops.gather_rows(g, plan.out_src, plan.R, out_dtype=F32)
    if plan.has_passthrough:
        ops.gather_sum_rows(local_out, plan.passthrough, 1, plan.R, out=out, accumulate=True)
    return out, (ctx if want_ctx else None)


def sttran_transformer_bwd(k: Kernels, P: dict, plan: Plan, ctx: dict, dout, grads: dict):
    n_enc, n_dec = len(ctx["enc"]), len(ctx.get("dec", []))
    if plan.Mg == 0:
        dlocal = dout
    else:
        dg, _ = ops.gather_rows(dout, plan.out_inv, plan.Mg, out_dtype=F32)
        dpe = torch.zeros(2, D_MODEL, device=dout.device, dtype=F32)
        for i in reversed(range(n_dec)):
            dg, dpos = decoder_bwd(k, P, f"{GT}global_attention.layers.{i}.", ctx["dec"][i], dg, plan.stream_slot,
                                   plan.glob_work, plan.n_glob_work, grads)
            ops.add(dpe, dpos, out=dpe)
        grads[GT + "position_embedding.weight"] = dpe
        dlocal = ops.gather_sum_rows(dg, plan.inv, 2, plan.R)
        if plan.has_passthrough:
            ops.gather_sum_rows(dout, plan.passthrough, 1, plan.R, out=dlocal, accumulate=True)
    for i in reversed(range(n_enc)):
        dlocal = encoder_bwd(k, P, f"{GT}local_attention.layers.{i}.", "self_attn", ctx["enc"][i], dlocal,
                             plan.local_work, plan.n_local_work, grads)
    return dlocal


def _heads_w(k: Kernels, P: dict):
    names = ("a_rel_compress", "s_rel_compress", "c_rel_compress")
    w = k.weight("heads.weight", tuple(P[n + ".weight"] for n in names), lambda a, b, c: torch.cat((a, b, c), 0), f32=True)
    b = k.weight("heads.bias", tuple(P[n + ".bias"] for n in names), lambda a, b, c: torch.cat((a, b, c)), f32=True)
    return w, b


def heads_fwd(k: Kernels, P: dict, x):
    """lib/sttran.py:404-406 as one [R,1936] x [26,1936]^T product (exact fp32) -> logits [R,26]."""
    w, b = _heads_w(k, P)
    return k.mm(x, w, bias=b, exact=True)


def heads_bwd(k: Kernels, P: dict, x, dlogits, grads: dict):
    w, _ = _heads_w(k, P)
    dw = k.mm(dlogits, x, a_major=MN_, b_major=MN_, exact=True)
    db = ops.colsum(dlogits).reshape(-1)
    for name, a, b_ in (("a_rel_compress", 0, 3), ("s_rel_compress", 3, 9), ("c_rel_compress", 9, 26)):
        grads[name + ".weight"], grads[name + ".bias"] = dw[a:b_].contiguous(), db[a:b_].contiguous()
    return k.mm(dlogits, w, b_major=MN_, exact=True)
