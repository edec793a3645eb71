"""Whole-model forward / backward / fused training step on top of engine.py.

A *batch* is a list of per-video ``entry`` dicts (the reference contract, lib/assign_pseudo_label.py:1368-1382):
the videos are concatenated into one set of device tensors and one Plan; BatchNorm statistics stay per video
and the batch loss is the mean over videos of the reference's per-video loss (tools/train_STTran.py:169-189).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import engine as E
from . import ops

F32 = torch.float32


class Batch:
    pass


TENSOR_KEYS = ("features", "boxes", "labels", "scores", "distribution", "union_feat", "pair_idx", "spatial_masks")


def collate(entries: List[dict], mode: str, pin: bool = False, feat_dtype: torch.dtype = F32) -> Batch:
    """Concatenate per-video entries (wherever their tensors live) into one Batch of the same residency.
    Host-side metadata (frame ids, per-video counts, label lists) is extracted once here.
    feat_dtype: storage type of the two feature tensors (`features`, `union_feat`).  fp32 is the reference's entry
    contract; bf16 is the packed-feature-file format of SURVEY 8f-2 (half the host -> device bytes; in the bf16 compute
    mode the results are bit-identical, the same round-to-nearest just happens in the loader instead of on the device)."""
    b = Batch()
    b.n_boxes = [int(e["boxes"].shape[0]) for e in entries]
    b.frame_ids = [e["im_idx"].detach().cpu().numpy() for e in entries]
    b.n_pairs = [len(f) for f in b.frame_ids]
    b.gt_lists = [(e.get("attention_gt"), e.get("spatial_gt"), e.get("contacting_gt")) for e in entries]
    off = np.concatenate(([0], np.cumsum(b.n_boxes)))[:-1]

    def cat(key, dtype):
        ts = [e[key] for e in entries]
        t = ts[0] if len(ts) == 1 else torch.cat(ts, 0)
        t = t.to(dtype).contiguous() if t.dtype != dtype else t.contiguous()
        return t.pin_memory() if (pin and not t.is_cuda) else t

    b.features, b.boxes = cat("features", feat_dtype), cat("boxes", F32)
    b.labels, b.scores = cat("labels", torch.int64), cat("scores", F32)
    b.distribution = cat("distribution", F32) if mode != "predcls" else None
    b.union_feat = cat("union_feat", feat_dtype)
    pi = [e["pair_idx"].to(torch.int64) + int(o) for e, o in zip(entries, off)]
    b.pair_idx = (pi[0] if len(pi) == 1 else torch.cat(pi, 0)).contiguous()
    if pin and not b.pair_idx.is_cuda:
        b.pair_idx = b.pair_idx.pin_memory()
    b.spatial_masks = cat("spatial_masks", F32) if all("spatial_masks" in e for e in entries) else None
    return b


def repack(hb: Batch, feat_dtype: torch.dtype, pin: bool = True) -> Batch:
    """Host batch with the two feature tensors stored as `feat_dtype` (what a loader of packed feature files hands over)."""
    b = Batch()
    b.__dict__.update(hb.__dict__)
    for key in ("features", "union_feat"):
        t = getattr(hb, key).to(feat_dtype)
        setattr(b, key, t.pin_memory() if (pin and not t.is_cuda) else t)
    return b


def upload(hb: Batch, device, rasterise: bool = True, consumer_stream=None) -> Batch:
    """Device copy of a collated batch (async copies on the current stream); rasterises the spatial masks on
    device when the producer did not supply them (fused pair gather + draw_union_boxes - 0.5).
    consumer_stream: the stream that will read the tensors when the copies are issued on a side stream."""
    dev = torch.device(device)
    b = Batch()
    b.__dict__.update(hb.__dict__)
    for key in TENSOR_KEYS:
        t = getattr(hb, key)
        if t is not None:
            d = t.to(dev, non_blocking=True)
            if consumer_stream is not None and d is not t:
                d.record_stream(consumer_stream)
            setattr(b, key, d)
    if rasterise:
        ensure_masks(b)
    return b


def ensure_masks(b: Batch) -> Batch:
    if b.spatial_masks is None:
        b.spatial_masks = ops.union_mask_pairs(b.boxes, b.pair_idx, 27, -0.5)
    return b


def input_bytes(hb: Batch) -> int:
    return int(sum(getattr(hb, k).numel() * getattr(hb, k).element_size() for k in TENSOR_KEYS if getattr(hb, k) is not None))


def make_plan(b: Batch, device, mode: str, dsg: bool = False, with_labels: bool = False, consumer_stream=None) -> "E.Plan":
    """Descriptors (+ optionally the loss labels) -> one pinned async upload.  plan.labels is set when with_labels."""
    obj_class = subj_box = None
    if dsg:
        lab = b.labels.cpu().numpy()
        pi = b.pair_idx.cpu().numpy()
        obj_class, subj_box = lab[pi[:, 1]], pi[:, 0]
    extra = label_arrays(b) if with_labels else None
    plan = E.Plan(b.n_boxes, b.frame_ids, torch.device(device), obj_class=obj_class, subj_box=subj_box, dsg=dsg,
                  dsg_pos_by_rank=(mode == "sgdet"), extra=extra, consumer_stream=consumer_stream)
    if with_labels:
        L = Labels()
        L.att, L.w_att, L.spa_bits, L.w_spa = plan.lab_att, plan.lab_w_att, plan.lab_spa_bits, plan.lab_w_spa
        L.con_bits, L.w_con, L.w_obj = plan.lab_con_bits, plan.lab_w_con, plan.lab_w_obj
        plan.labels = L
    return plan


def make_batch(entries: List[dict], device, mode: str, dsg: bool = False):
    """collate + upload + plan; returns (Batch on device, Plan)."""
    hb = collate(entries, mode)
    plan = make_plan(hb, device, mode, dsg)
    return upload(hb, device), plan


# ================================================================================================
# STTran
# ================================================================================================
def sttran_forward(k: E.Kernels, P: Dict[str, torch.Tensor], batch: Batch, plan: E.Plan, mode: str, training: bool,
                   want_ctx: bool):
    """lib/sttran.py:375-411.  Returns (outputs dict, ctx)."""
    ctx = {}
    out = {}
    if mode == "predcls":
        feat_op = k.opnd(batch.features)
        ctx["oc"] = None
    else:
        logits, objfeat, ctx["oc"] = E.object_classifier_fwd(k, P, plan, batch.features, batch.distribution, batch.boxes,
                                                            training, want_ctx)
        out["distribution"] = logits
        feat_op = objfeat[:, :2048]
    rel, ctx["pt"] = E.pair_tokens_fwd(k, P, plan, feat_op, batch.union_feat, batch.spatial_masks, batch.pair_idx,
                                       batch.labels, training, want_ctx)
    glob, ctx["tr"] = E.sttran_transformer_fwd(k, P, plan, rel, want_ctx)
    logits26 = E.heads_fwd(k, P, glob)
    ctx["glob"], ctx["logits26"] = glob, logits26
    out["logits26"] = logits26
    return out, (ctx if want_ctx else None)


def sttran_backward(k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, ctx: dict, dlogits26, dobj_logits, grads=None):
    """`grads`: any mapping that accepts grads[name] = tensor (a dict, or the trainer's flat-buffer sink)."""
    grads = {} if grads is None else grads
    dglob = E.heads_bwd(k, P, ctx["glob"], dlogits26, grads)
    drel = E.sttran_transformer_bwd(k, P, plan, ctx["tr"], dglob, grads)
    E.pair_tokens_bwd(k, P, plan, ctx["pt"], drel, grads)
    if mode != "predcls" and dobj_logits is not None:
        E.object_classifier_bwd(k, P, plan, ctx["oc"], dobj_logits, grads)
    return grads


# ================================================================================================
# DSG-DETR (sgdet): lib/dsg_detr.py:514-572
# ================================================================================================
def dsg_forward(k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, training: bool, want_ctx: bool):
    ctx = {}
    out = {}
    if mode == "predcls":
        feat_op = k.opnd(batch.features)
        ctx["oc"] = None
    else:
        logits, objfeat, ctx["oc"] = E.object_classifier_fwd(k, P, plan, batch.features, batch.distribution, batch.boxes,
                                                            training, want_ctx)
        out["distribution"] = logits
        feat_op = objfeat[:, :2048]
    rel, ctx["pt"] = E.pair_tokens_fwd(k, P, plan, feat_op, batch.union_feat, batch.spatial_masks, batch.pair_idx,
                                       batch.labels, training, want_ctx)
    x, _, ctx["loc"] = E.encoder_fwd(k, P, "local_transformer.layers.0.", "self_attn", rel, k.opnd(rel), plan.local_work,
                                     plan.n_local_work, want_ctx, out_op=False)
    pe = P["positional_encoder.pe"].reshape(-1, E.D_MODEL)
    # class-sorted stream + sinusoidal encoding of the frame rank (dsg_detr.py:545-559)
    g, gop = ops.gather_rows(x, plan.cls_perm, plan.R, out_dtype=F32, add=pe, add_idx=plan.cls_pos,
                             out2_dtype=torch.bfloat16 if k.AD == torch.bfloat16 else None)
    if gop is None:
        gop = g
    ctx["glob"] = []
    for i in range(3):
        g, gop, c = E.encoder_fwd(k, P, f"global_transformer.layers.{i}.", "self_attn", g, gop, plan.cls_work,
                                  plan.n_cls_work, want_ctx, out_op=(i < 2))
        ctx["glob"].append(c)
    glob, _ = ops.gather_rows(g, plan.cls_iperm, plan.R, out_dtype=F32)
    logits26 = E.heads_fwd(k, P, glob)
    ctx["globout"], ctx["logits26"] = glob, logits26
    out["logits26"] = logits26
    return out, (ctx if want_ctx else None)


def dsg_backward(k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, ctx: dict, dlogits26, dobj_logits, grads=None):
    grads = {} if grads is None else grads
    dglob = E.heads_bwd(k, P, ctx["globout"], dlogits26, grads)
    dg, _ = ops.gather_rows(dglob, plan.cls_perm, plan.R, out_dtype=F32)
    for i in reversed(range(3)):
        dg = E.encoder_bwd(k, P, f"global_transformer.layers.{i}.", "self_attn", ctx["glob"][i], dg, plan.cls_work,
                           plan.n_cls_work, grads)
    dx, _ = ops.gather_rows(dg, plan.cls_iperm, plan.R, out_dtype=F32)   # the encoding is a constant buffer
    drel = E.encoder_bwd(k, P, "local_transformer.layers.0.", "self_attn", ctx["loc"], dx, plan.local_work,
                         plan.n_local_work, grads)
    E.pair_tokens_bwd(k, P, plan, ctx["pt"], drel, grads)
    if mode != "predcls" and dobj_logits is not None:
        E.object_classifier_bwd(k, P, plan, ctx["oc"], dobj_logits, grads)
    return grads


# ================================================================================================
# fused loss (tools/train_STTran.py:143-189 with bce_loss=True), batch = mean over videos
# ================================================================================================
class Labels:
    pass


def _bits_of(lists) -> np.ndarray:
    """uint32 multi-hot mask per row from a list of index lists (vectorised)."""
    n = len(lists)
    lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=n)
    out = np.zeros(n, dtype=np.uint32)
    if lens.sum() == 0:
        return out
    flat = np.fromiter((int(j) for x in lists for j in x), dtype=np.int64, count=int(lens.sum()))
    rows = np.repeat(np.arange(n), lens)
    np.bitwise_or.at(out, rows, (np.uint32(1) << flat.astype(np.uint32)))
    return out


def label_arrays(batch: Batch) -> dict:
    """Label tensors + per-row loss weights from the python label lists of the entries (train_STTran.py:143-167).
    Weight = 1 / (rows of that video entering the mean) / (classes, for BCE) / videos."""
    nv = len(batch.n_boxes)
    att, w_att, spa_bits, w_spa, con_bits, w_con, w_obj = [], [], [], [], [], [], []
    for (a_gt, s_gt, c_gt), nb in zip(batch.gt_lists, batch.n_boxes):
        n = len(a_gt)
        a = np.fromiter((int(x[0]) if len(x) else -1 for x in a_gt), dtype=np.int64, count=n)
        na = int((a >= 0).sum())
        att.append(a)
        w_att.append(np.where(a >= 0, 1.0 / (max(na, 1) * nv), 0.0).astype(np.float32))
        sb, cb = _bits_of(s_gt), _bits_of(c_gt)
        ns, nc = int((sb != 0).sum()), int((cb != 0).sum())
        spa_bits.append(sb); con_bits.append(cb)
        w_spa.append(np.where(sb != 0, 1.0 / (max(ns, 1) * 6 * nv), 0.0).astype(np.float32))
        w_con.append(np.where(cb != 0, 1.0 / (max(nc, 1) * 17 * nv), 0.0).astype(np.float32))
        w_obj.append(np.full(nb, 1.0 / (max(nb, 1) * nv), dtype=np.float32))
    c = np.concatenate
    return {"lab_att": c(att), "lab_w_att": c(w_att), "lab_spa_bits": c(spa_bits), "lab_w_spa": c(w_spa),
            "lab_con_bits": c(con_bits), "lab_w_con": c(w_con), "lab_w_obj": c(w_obj)}


def make_labels(batch: Batch, device, mode: str) -> Labels:
    from .plan import pack_upload
    d = pack_upload(label_arrays(batch), device)
    L = Labels()
    L._keep = d.pop("_pinned_keepalive")
    L.att, L.w_att, L.spa_bits, L.w_spa = d["lab_att"], d["lab_w_att"], d["lab_spa_bits"], d["lab_w_spa"]
    L.con_bits, L.w_con, L.w_obj = d["lab_con_bits"], d["lab_w_con"], d["lab_w_obj"]
    return L


def fused_loss(out: dict, batch: Batch, labels: Labels, mode: str, want_grad: bool = True):
    """Returns (loss scalar tensor on device, dlogits26, dobj_logits)."""
    logits26 = out["logits26"]
    dev = logits26.device
    loss = torch.zeros(1, device=dev, dtype=F32)
    d26 = torch.empty_like(logits26) if want_grad else None
    dobj = None
    if mode != "predcls":
        dobj = torch.empty_like(out["distribution"]) if want_grad else None
        ops.ce_loss(out["distribution"], 37, batch.labels, labels.w_obj, loss, dobj)
    ops.ce_loss(logits26[:, 0:3], 3, labels.att, labels.w_att, loss, d26[:, 0:3] if want_grad else None)
    ops.bce_sigmoid_loss(logits26[:, 3:9], 6, labels.spa_bits, labels.w_spa, loss, d26[:, 3:9] if want_grad else None)
    ops.bce_sigmoid_loss(logits26[:, 9:26], 17, labels.con_bits, labels.w_con, loss, d26[:, 9:26] if want_grad else None)
    return loss, d26, dobj
