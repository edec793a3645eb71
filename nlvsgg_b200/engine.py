"""Host side of the model sequencer (csrc/step.cu): parameter-slot tables, the ctypes descriptors and the workspace.

The whole forward / loss / backward of STTran or DSG-DETR is ONE C call each (`nlv_session_forward`,
`nlv_session_backward`): this module only names the parameters (slot table = the reference's state_dict names,
lib/sttran.py:316-372, lib/dsg_detr.py:466-512), points at the batch tensors and owns the workspace buffer the
C side bump-allocates from.  No arithmetic happens here.

Precision modes (`Kernels(precision)`, env NLV_PRECISION)
  * 'bf16'   : GEMM operands bf16 on tcgen05, fp32 accumulation, fp32 residual stream / norms / softmax
  * 'bf16x3' : fp32 tensors, each operand split hi/lo into 3 bf16 K-blocks -> fp32-faithful on the same tcgen05 kernel
  * 'fp32'   : exact-fp32 SIMT GEMM (debug / tiny problems)
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _C
from .plan import Plan  # noqa: F401  (host-side descriptors live in plan.py)

D_MODEL, N_HEAD, HEAD_DIM, D_FF = 1936, 8, 242, 2048
F32, BF16 = torch.float32, torch.bfloat16

_COMMON = [
    "object_classifier.obj_embed.weight",
    "object_classifier.pos_embed.0.weight", "object_classifier.pos_embed.0.bias",
    "object_classifier.pos_embed.0.running_mean", "object_classifier.pos_embed.0.running_var",
    "object_classifier.pos_embed.1.weight", "object_classifier.pos_embed.1.bias",
    "object_classifier.decoder_lin.0.weight", "object_classifier.decoder_lin.0.bias",
    "object_classifier.decoder_lin.1.weight", "object_classifier.decoder_lin.1.bias",
    "object_classifier.decoder_lin.1.running_mean", "object_classifier.decoder_lin.1.running_var",
    "object_classifier.decoder_lin.3.weight", "object_classifier.decoder_lin.3.bias",
    "union_func1.weight", "union_func1.bias",
    "conv.0.weight", "conv.0.bias",
    "conv.2.weight", "conv.2.bias", "conv.2.running_mean", "conv.2.running_var",
    "conv.4.weight", "conv.4.bias",
    "conv.6.weight", "conv.6.bias", "conv.6.running_mean", "conv.6.running_var",
    "subj_fc.weight", "subj_fc.bias", "obj_fc.weight", "obj_fc.bias",
    "vr_fc.weight", "vr_fc.bias",
    "obj_embed.weight", "obj_embed2.weight",
    "a_rel_compress.weight", "a_rel_compress.bias", "s_rel_compress.weight", "s_rel_compress.bias",
    "c_rel_compress.weight", "c_rel_compress.bias",
]
P_POS = len(_COMMON)          # NLV_P_POS
P_LAYER0 = P_POS + 1          # NLV_P_LAYER0
LAYER_STRIDE = 12             # NLV_P_LAYER_STRIDE


def _layer_names(prefix: str, attn: str, norm_a: str, norm_b: Optional[str]) -> List[Optional[str]]:
    n = [f"{prefix}{attn}.in_proj_weight", f"{prefix}{attn}.in_proj_bias", f"{prefix}{attn}.out_proj.weight",
         f"{prefix}{attn}.out_proj.bias", f"{prefix}linear1.weight", f"{prefix}linear1.bias", f"{prefix}linear2.weight",
         f"{prefix}linear2.bias", f"{prefix}{norm_a}.weight", f"{prefix}{norm_a}.bias"]
    return n + ([f"{prefix}{norm_b}.weight", f"{prefix}{norm_b}.bias"] if norm_b else [None, None])


def slot_names(arch: str, n_enc: int = 1, n_dec: int = 3, transformer_prefix: str = "glocal_transformer.") -> List[Optional[str]]:
    """state_dict name of every NLV_P_* slot (None = unused slot)."""
    names: List[Optional[str]] = list(_COMMON)
    if arch == "sttran":
        names.append(transformer_prefix + "position_embedding.weight")
        for i in range(n_enc):
            names += _layer_names(f"{transformer_prefix}local_attention.layers.{i}.", "self_attn", "norm1", "norm2")
        for i in range(n_dec):
            names += _layer_names(f"{transformer_prefix}global_attention.layers.{i}.", "multihead2", "norm3", None)
    else:
        names.append("positional_encoder.pe")
        names += _layer_names("local_transformer.layers.0.", "self_attn", "norm1", "norm2")
        for i in range(3):
            names += _layer_names(f"global_transformer.layers.{i}.", "self_attn", "norm1", "norm2")
    return names


def count_layers(P: Dict[str, torch.Tensor], prefix: str = "glocal_transformer."):
    n_enc = sum(1 for n in P if n.startswith(prefix + "local_attention.layers.") and n.endswith("norm1.weight"))
    n_dec = sum(1 for n in P if n.startswith(prefix + "global_attention.layers.") and n.endswith("norm3.weight"))
    return n_enc, n_dec


_NO_GRAD_SUFFIX = ("running_mean", "running_var", "num_batches_tracked", ".pe")


class Kernels:
    """Precision mode + the reusable C session / workspace of one model instance."""

    def __init__(self, precision: str = "bf16", dropout: float = 0.0):
        assert precision in ("bf16", "bf16x3", "fp32")
        self.precision = precision
        self.dropout = dropout
        self.additive_mask = False
        self.transformer_both = False      # temporal decoder output mode 'both' (standalone lib/transformer_wk.py module only)
        self.AD = BF16 if precision == "bf16" else F32
        self.mirror: Dict[str, torch.Tensor] = {}      # name -> bf16 operand copy kept current by the caller (trainer)
        self._ws: Optional[torch.Tensor] = None
        self.seed = 0x5EED

    def workspace(self, nbytes: int, device, fresh: bool) -> torch.Tensor:
        """fresh: a new buffer per call (its views are handed to the caller: drop-in modules); otherwise one grow-only buffer."""
        if fresh:
            return torch.empty(nbytes, dtype=torch.uint8, device=device)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != torch.device(device):
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.1) + (1 << 20), dtype=torch.uint8, device=device)
        return self._ws


class Session:
    """One forward (+ loss) (+ backward) of a model on a batch through the C sequencer."""

    def __init__(self):
        self.h = _C.lib().nlv_session_create()
        self.keep = []          # python objects that must outlive the C calls (tables, tensors)
        self.ws = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _C.lib().nlv_session_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _ptr(t) -> Optional[int]:
    return t.data_ptr() if t is not None else None


class ModelDesc:
    """nlv_model for a parameter dict.  The pointer tables are rebuilt on every call (parameters of a drop-in module are
    ordinary torch tensors that an optimiser may have replaced) unless `static` (trainer: views of one flat buffer)."""

    def __init__(self, k: Kernels, P: Dict[str, torch.Tensor], arch: str, mode: str, transformer_prefix: str = "glocal_transformer.",
                 static: bool = False):
        self.k, self.arch, self.mode = k, arch, mode
        if arch == "sttran":
            self.n_enc, self.n_dec = count_layers(P, transformer_prefix)
        else:
            self.n_enc, self.n_dec = 1, 3
        self.names = slot_names(arch, self.n_enc, self.n_dec, transformer_prefix)
        self.n_slots = len(self.names)
        self.static = static
        self._tables = None
        # gradient layout: every trainable slot, in slot order, 8-float aligned
        self.grad_names: List[str] = []
        self.grad_off: Dict[str, int] = {}
        off = 0
        for n in self.names:
            if n is None or n not in P or n.endswith(_NO_GRAD_SUFFIX):
                continue
            if mode == "predcls" and n.startswith("object_classifier."):
                continue       # never touched in predcls (lib/sttran.py:90-92): grad stays None, as in the reference
            self.grad_names.append(n)
            self.grad_off[n] = off
            off += (P[n].numel() + 7) // 8 * 8
        self.grad_elems = off

    def tables(self, P):
        if self._tables is not None:
            return self._tables
        PA = ctypes.c_void_p * self.n_slots
        pa, po = PA(), PA()
        for i, n in enumerate(self.names):
            t = P.get(n) if n is not None else None
            if t is not None:
                if not t.is_cuda:
                    raise RuntimeError("nlv_b200 kernels need CUDA tensors; there is no CPU fallback")
                if t.dtype != F32 or not t.is_contiguous():
                    raise TypeError(f"parameter {n}: expected a contiguous fp32 tensor")
                pa[i] = t.data_ptr()
            m = self.k.mirror.get(n) if n is not None else None
            if m is not None:
                po[i] = m.data_ptr()
        if self.static:
            self._tables = (pa, po)
        return pa, po

    def grad_offsets(self, base_names: Optional[Dict[str, int]] = None):
        LA = ctypes.c_longlong * self.n_slots
        go = LA()
        src = base_names if base_names is not None else self.grad_off
        for i, n in enumerate(self.names):
            go[i] = src.get(n, -1) if n is not None else -1
        return go

    def fill(self, P, training: bool, grad_base: Optional[torch.Tensor] = None, grad_offsets=None) -> "_C.Model":
        m = _C.Model()
        pa, po = self.tables(P)
        m.arch, m.mode, m.precision = _C.ARCH[self.arch], _C.MODE[self.mode], _C.PREC[self.k.precision]
        m.n_enc, m.n_dec, m.training, m.n_slots = self.n_enc, self.n_dec, 1 if training else 0, self.n_slots
        m.params = ctypes.cast(pa, ctypes.c_void_p)
        m.params_op = ctypes.cast(po, ctypes.c_void_p) if self.k.mirror else None
        if grad_base is not None:
            m.grad_base, m.grad_elems = grad_base.data_ptr(), grad_base.numel()
            m.grad_offset = ctypes.cast(grad_offsets, ctypes.c_void_p)
        m.dropout_p = float(self.k.dropout) if training else 0.0
        m.seed = self.k.seed
        m.additive_mask = 1 if self.k.additive_mask else 0
        m.transformer_both = 1 if self.k.transformer_both else 0
        pe = P.get("positional_encoder.pe")
        m.pe_rows = int(pe.shape[-2]) if pe is not None else 0
        m._keep = (pa, po, grad_offsets)
        return m


def plan_desc(plan: Plan, dsg: bool = False) -> "_C.BatchDesc":
    """The descriptor half of nlv_batch (sizes + the int32 index arrays of plan.py)."""
    b = _C.BatchDesc()
    b.nv, b.n_boxes, b.n_pairs, b.n_stream = plan.nv, plan.N, plan.R, plan.Mg
    for name in ("box_seg", "seg196", "seg49", "box_row", "pair_row", "local_work", "glob_work", "stream_src", "stream_slot",
                 "inv", "out_src", "out_inv", "passthrough", "both_w"):
        setattr(b, name, _ptr(getattr(plan, name, None)))
    b.n_local_work, b.n_glob_work, b.has_passthrough = plan.n_local_work, plan.n_glob_work, 1 if plan.has_passthrough else 0
    b.work_sorted, b.n_local_long, b.n_glob_long = 1, plan.n_local_long, plan.n_glob_long
    if dsg:
        b.cls_perm, b.cls_iperm, b.cls_pos, b.cls_work = (plan.cls_perm.data_ptr(), plan.cls_iperm.data_ptr(),
                                                          plan.cls_pos.data_ptr(), plan.cls_work.data_ptr())
        b.n_cls_work = plan.n_cls_work
        b.n_cls_long = plan.n_cls_long
    return b


def batch_desc(batch, plan: Plan, labels=None, dsg: bool = False) -> "_C.BatchDesc":
    """nlv_batch from a device-resident Batch (model.py) + Plan (plan.py) (+ Labels)."""
    b = plan_desc(plan, dsg)
    f = batch.features
    for t in (f, batch.boxes, batch.labels, batch.union_feat, batch.pair_idx):
        if not t.is_cuda:
            raise RuntimeError("nlv_b200 kernels need CUDA tensors; there is no CPU fallback")
    b.features, b.feat_dtype = f.data_ptr(), _C.NLV_BF16 if f.dtype == BF16 else _C.NLV_F32
    b.boxes, b.labels = batch.boxes.data_ptr(), batch.labels.data_ptr()
    b.distribution = _ptr(batch.distribution)
    u = batch.union_feat
    b.union_feat, b.union_dtype = u.data_ptr(), _C.NLV_BF16 if u.dtype == BF16 else _C.NLV_F32
    b.union_rows = int(getattr(batch, "union_rows", 0) or 0)
    b.union_bitmap, b.union_off = _ptr(getattr(batch, "union_bitmap", None)), _ptr(getattr(batch, "union_off", None))
    b.union_hx, b.union_base = _ptr(getattr(batch, "union_hx", None)), _ptr(getattr(batch, "union_base", None))
    ex = getattr(batch, "union_exc_pos", None)
    b.union_exc_pos, b.union_exc_val = _ptr(ex), _ptr(getattr(batch, "union_exc_val", None))
    b.n_union_exc = int(ex.numel()) if ex is not None else 0
    if b.union_rows == 3:
        b.union_dtype = _C.NLV_BF16       # 12-bit stored values decode to bf16
    b.dist_conf, b.dist_idx = _ptr(getattr(batch, "dist_conf", None)), _ptr(getattr(batch, "dist_idx", None))
    b.dist_other = _ptr(getattr(batch, "dist_other", None))
    b.spatial_masks = _ptr(batch.spatial_masks)
    b.pair_idx = batch.pair_idx.data_ptr()
    if labels is not None:
        b.lab_att, b.w_att, b.spa_bits, b.w_spa = _ptr(labels.att), _ptr(labels.w_att), _ptr(labels.spa_bits), _ptr(labels.w_spa)
        b.con_bits, b.w_con, b.w_obj = _ptr(labels.con_bits), _ptr(labels.w_con), _ptr(labels.w_obj)
    return b


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _view(ws: torch.Tensor, ptr: Optional[int], shape) -> Optional[torch.Tensor]:
    """fp32 tensor view of a region of the workspace the C side handed back."""
    if not ptr:
        return None
    off = ptr - ws.data_ptr()
    n = 1
    for s in shape:
        n *= int(s)
    return ws[off:off + 4 * n].view(F32).view(*shape)


def run_forward(k: Kernels, desc: ModelDesc, P, batch, plan: Plan, training: bool, want_ctx: bool, labels=None, with_loss: bool = False,
                activations: bool = False, with_backward: bool = False, fresh_ws: bool = False,
                grad_base: Optional[torch.Tensor] = None, grad_offsets=None, object_only: bool = False):
    """-> (outputs dict, Session).  The session holds everything backward needs (and keeps the workspace alive)."""
    lib = _C.lib()
    dev = batch.boxes.device
    sess = Session()
    flags = (_C.RUN_CTX if want_ctx else 0) | (_C.RUN_LOSS if with_loss else 0) | (_C.RUN_ACTIVATIONS if activations else 0) | \
        (_C.RUN_OBJECT_ONLY if object_only else 0)
    m = desc.fill(P, training, grad_base, grad_offsets)
    b = batch_desc(batch, plan, labels, dsg=(desc.arch == "dsg"))
    need = lib.nlv_session_plan(sess.h, ctypes.byref(m), ctypes.byref(b), flags | (_C.RUN_BACKWARD if (want_ctx and with_backward) else 0))
    if need < 0:
        _C.check(int(need), "session_plan")
    ws = k.workspace(int(need), dev, fresh_ws)
    out = _C.Outputs()
    _C.check(lib.nlv_session_forward(sess.h, ctypes.byref(m), ctypes.byref(b), ws.data_ptr(), ws.numel(), flags, ctypes.byref(out), _stream()),
             "session_forward")
    sess.ws, sess.keep = ws, [m, b, batch, plan, labels, P]
    N, R = plan.N, plan.R
    o = {"logits26": _view(ws, out.logits26, (R, 26))} if not object_only else {}
    if out.obj_logits:
        o["distribution"] = _view(ws, out.obj_logits, (N, 37))
    if activations:
        o["att"], o["spa"], o["con"] = _view(ws, out.att, (R, 3)), _view(ws, out.spa, (R, 6)), _view(ws, out.con, (R, 17))
    if with_loss:
        o["loss"] = _view(ws, out.loss, (1,))
        o["d26"], o["dobj"] = _view(ws, out.d26, (R, 26)), _view(ws, out.dobj, (N, 37))
    o["spatial_masks"] = batch.spatial_masks if batch.spatial_masks is not None else _view(ws, out.masks, (R, 2, 27, 27))
    o["rel_tokens"], o["rel_out"] = _view(ws, out.rel_tokens, (R, D_MODEL)), _view(ws, out.rel_out, (R, D_MODEL))
    return o, sess


def set_gradients(sess: Session, desc: ModelDesc, grad_base: torch.Tensor, grad_offsets=None):
    go = grad_offsets if grad_offsets is not None else desc.grad_offsets()
    sess.keep.append(go)
    _C.check(_C.lib().nlv_session_set_gradients(sess.h, grad_base.data_ptr(), grad_base.numel(), ctypes.cast(go, ctypes.c_void_p),
                                                desc.n_slots), "session_set_gradients")


def run_backward(sess: Session, d26: Optional[torch.Tensor] = None, dobj: Optional[torch.Tensor] = None):
    _C.check(_C.lib().nlv_session_backward(sess.h, _ptr(d26), _ptr(dobj), _stream()), "session_backward")


# ---- standalone spatio-temporal transformer (lib/transformer_wk.py forward(features, im_idx)) ---------------------------
def run_transformer_forward(k: Kernels, desc: ModelDesc, P, plan: Plan, x: torch.Tensor, training: bool, want_ctx: bool):
    lib = _C.lib()
    sess = Session()
    m = desc.fill(P, training)
    b = plan_desc(plan)
    flags = _C.RUN_CTX if want_ctx else 0
    need = lib.nlv_session_transformer_forward(sess.h, ctypes.byref(m), ctypes.byref(b), x.data_ptr(), None, 0,
                                               flags | (_C.RUN_BACKWARD if want_ctx else 0), None, _stream())
    if need < 0:
        _C.check(int(need), "session_transformer_forward(plan)")
    ws = k.workspace(int(need), x.device, True)
    out = ctypes.c_void_p()
    rc = lib.nlv_session_transformer_forward(sess.h, ctypes.byref(m), ctypes.byref(b), x.data_ptr(), ws.data_ptr(), ws.numel(), flags,
                                             ctypes.byref(out), _stream())
    _C.check(int(rc), "session_transformer_forward")
    sess.ws, sess.keep = ws, [m, b, plan, P, x]
    if out.value == x.data_ptr():      # zero layers: nothing to do
        return x, sess
    return _view(ws, out.value, (plan.R, D_MODEL)), sess


def run_transformer_backward(sess: Session, desc: ModelDesc, dout: torch.Tensor, P):
    """-> (dx [R,1936], {name: grad})"""
    flat = torch.empty(max(desc.grad_elems, 8), device=dout.device, dtype=F32)
    set_gradients(sess, desc, flat)
    dx = ctypes.c_void_p()
    _C.check(_C.lib().nlv_session_transformer_backward(sess.h, dout.data_ptr(), ctypes.byref(dx), _stream()), "session_transformer_backward")
    grads = {n: flat[desc.grad_off[n]:desc.grad_off[n] + P[n].numel()].view(P[n].shape) for n in desc.grad_names}
    if dx.value == dout.data_ptr():
        dxt = dout
    else:
        dxt = _view(sess.ws, dx.value, (dout.shape[0], D_MODEL))
    return dxt, grads
