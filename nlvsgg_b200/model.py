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
    """A collated batch of videos (host or device resident).  `union_rows`: 0 = union_feat is NCHW [R,2048,7,7] (the entry
    contract), 1 = channels-last rows [R*49,2048], 2 = zero-suppressed rows (values + union_bitmap + union_off), 3 = the same with
    12-bit values (union_feat = low bytes, union_hx = 4-bit high-byte codes, union_base = the rows' smallest high bytes) — the
    layouts of the packed feature files (featfile.py)."""
    union_rows = 0
    union_bitmap = union_off = union_hx = union_base = union_exc_pos = union_exc_val = dist_conf = dist_other = dist_idx = None
    distribution = spatial_masks = None
    lab_csr = None


TENSOR_KEYS = ("features", "boxes", "labels", "scores", "distribution", "union_feat", "pair_idx", "spatial_masks",
               "union_bitmap", "union_off", "union_hx", "union_base", "union_exc_pos", "union_exc_val", "dist_conf", "dist_other", "dist_idx")


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
    b.lab_csr = _label_csr(b.gt_lists) if all(g[0] is not None for g in b.gt_lists) else None
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
        t = getattr(hb, key, None)
        if t is not None:
            d = t.to(dev, non_blocking=True)
            if consumer_stream is not None and d is not t:
                d.record_stream(consumer_stream)
            setattr(b, key, d)
    if rasterise:
        ensure_masks(b)
    return b


def upload_from_side_stream(hb: Batch, device, copy_stream) -> Batch:
    """Device copy of a collated batch with the copies on `copy_stream` and the MEMORY from the current (consumer) stream's pool.

    Destination blocks allocated under the copy stream (and handed to the consumer with record_stream) come back to the caching
    allocator only once a cross-stream event has completed; when the next batch is allocated before that, the allocator calls
    cudaMalloc — a device-wide synchronisation — every step, and the pipelined loop runs 3x slower for the whole run (seen as a
    bimodal end-to-end time).  Blocks of the consumer's own pool are reusable in stream order the moment the previous batch is
    dropped; the copy stream only has to wait for the work already enqueued on the consumer stream before overwriting them."""
    dev = torch.device(device)
    main = torch.cuda.current_stream(dev)
    dst = {k: torch.empty_like(getattr(hb, k), device=dev) for k in TENSOR_KEYS if getattr(hb, k, None) is not None}
    b = Batch()
    b.__dict__.update(hb.__dict__)
    copy_stream.wait_stream(main)
    with torch.cuda.stream(copy_stream):
        for k, d in dst.items():
            d.copy_(getattr(hb, k), non_blocking=True)
            setattr(b, k, d)
    return b


def ensure_masks(b: Batch) -> Batch:
    if b.spatial_masks is None:
        b.spatial_masks = ops.union_mask_pairs(b.boxes, b.pair_idx, 27, -0.5)
    return b


def input_bytes(hb: Batch) -> int:
    return int(sum(getattr(hb, k).numel() * getattr(hb, k).element_size() for k in TENSOR_KEYS if getattr(hb, k, None) is not None))


def make_plan(b: Batch, device, mode: str, dsg: bool = False, with_labels: bool = False, consumer_stream=None,
              label_rng=None, stager=None) -> "E.Plan":
    """Descriptors (+ optionally the loss labels) -> one pinned async upload.  plan.labels is set when with_labels."""
    obj_class = subj_box = None
    if dsg:
        lab = b.labels.cpu().numpy()
        pi = b.pair_idx.cpu().numpy()
        obj_class, subj_box = lab[pi[:, 1]], pi[:, 0]
    extra = label_arrays(b, label_rng) if with_labels else None
    plan = E.Plan(b.n_boxes, b.frame_ids, torch.device(device), obj_class=obj_class, subj_box=subj_box, dsg=dsg,
                  dsg_pos_by_rank=(mode == "sgdet"), extra=extra, consumer_stream=consumer_stream, stager=stager)
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
    return upload(hb, device, rasterise=False), plan      # masks the producer did not supply are rasterised inside the forward call


# ================================================================================================
# whole-model forward / backward: one C call each (csrc/step.cu)
# ================================================================================================
def _desc(k: E.Kernels, P, arch: str, mode: str) -> E.ModelDesc:
    cache = k.__dict__.setdefault("_desc_cache", {})
    key = (arch, mode, len(P))
    d = cache.get(key)
    if d is None:
        d = cache[key] = E.ModelDesc(k, P, arch, mode)
    return d


def _forward(arch: str, k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, training: bool, want_ctx: bool, **kw):
    desc = _desc(k, P, arch, mode)
    kw.setdefault("fresh_ws", True)     # the outputs are views of the call's workspace: callers of this API own them
    out, sess = E.run_forward(k, desc, P, batch, plan, training, want_ctx, **kw)
    if batch.spatial_masks is None:
        batch.spatial_masks = out["spatial_masks"]      # rasterised inside the call (a3)
    return out, ((sess, desc) if want_ctx else None)


def _backward(k: E.Kernels, P, ctx, d26, dobj, grads=None):
    """Returns {name: gradient} — views of ONE flat fp32 buffer (zeroed by the call, so parameters the batch does not
    reach come back as exact zeros)."""
    sess, desc = ctx
    flat = torch.empty(desc.grad_elems, device=d26.device, dtype=F32)
    E.set_gradients(sess, desc, flat)
    E.run_backward(sess, d26.contiguous(), dobj.contiguous() if dobj is not None else None)
    grads = {} if grads is None else grads
    for n in desc.grad_names:
        o = desc.grad_off[n]
        grads[n] = flat[o:o + P[n].numel()].view(P[n].shape)
    return grads


def sttran_forward(k: E.Kernels, P: Dict[str, torch.Tensor], batch: Batch, plan: E.Plan, mode: str, training: bool,
                   want_ctx: bool, **kw):
    """lib/sttran.py:375-411.  Returns (outputs dict, ctx)."""
    return _forward("sttran", k, P, batch, plan, mode, training, want_ctx, **kw)


def sttran_backward(k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, ctx, dlogits26, dobj_logits, grads=None):
    return _backward(k, P, ctx, dlogits26, dobj_logits if mode != "predcls" else None, grads)


def dsg_forward(k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, training: bool, want_ctx: bool, **kw):
    """DSG-DETR (sgdet): lib/dsg_detr.py:514-572."""
    return _forward("dsg", k, P, batch, plan, mode, training, want_ctx, **kw)


def dsg_backward(k: E.Kernels, P, batch: Batch, plan: E.Plan, mode: str, ctx, dlogits26, dobj_logits, grads=None):
    return _backward(k, P, ctx, dlogits26, dobj_logits if mode != "predcls" else None, grads)


# ================================================================================================
# fused loss (tools/train_STTran.py:143-189 with bce_loss=True), batch = mean over videos
# ================================================================================================
class Labels:
    pass


def _flat_lists(lists):
    """(values int64[total], row int64[total], lens int64[n]) of a list of index lists."""
    n = len(lists)
    lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=n)
    total = int(lens.sum())
    vals = np.fromiter((j for x in lists for j in x), dtype=np.int64, count=total) if total else np.zeros(0, dtype=np.int64)
    return vals, np.repeat(np.arange(n), lens), lens


def _bits_of(lists) -> np.ndarray:
    """uint32 multi-hot mask per row from a list of index lists (vectorised)."""
    vals, rows, lens = _flat_lists(lists)
    out = np.zeros(len(lists), dtype=np.uint32)
    if len(vals):
        np.bitwise_or.at(out, rows, (np.uint32(1) << vals.astype(np.uint32)))
    return out


def _label_csr(gt_lists):
    """The python label lists of a batch flattened once (loader-side work: it walks every list):
    (attention values, attention list lengths, spatial bit masks, contacting bit masks)."""
    a_all = [x for g in gt_lists for x in g[0]]
    s_all = [x for g in gt_lists for x in g[1]]
    c_all = [x for g in gt_lists for x in g[2]]
    avals, _, alens = _flat_lists(a_all)
    return avals, alens, _bits_of(s_all), _bits_of(c_all)


def label_arrays(batch: Batch, rng: Optional[np.random.Generator] = None) -> dict:
    """Label tensors + per-row loss weights from the label lists of the entries (train_STTran.py:143-167).
    Weight = 1 / (rows of that video entering the mean) / (classes, for BCE) / videos.  A pair with several attention
    labels contributes ONE of them per step: the reference draws it with np.random.choice (:148-150) — pass `rng` for that;
    without it the first label is taken (deterministic; what the parity fixtures use).  One numpy pass over all videos; the
    python lists themselves were flattened at collate time (`batch.lab_csr`)."""
    nv = len(batch.n_boxes)
    n_pairs = np.asarray(batch.n_pairs, dtype=np.int64)
    csr = getattr(batch, "lab_csr", None)
    avals, alens, sb, cb = csr if csr is not None else _label_csr(batch.gt_lists)
    R = len(alens)
    vid = np.repeat(np.arange(nv), n_pairs)
    pick = np.cumsum(alens) - alens
    if rng is not None and len(avals):
        multi = alens >= 2
        pick = pick.copy()
        pick[multi] += (rng.random(int(multi.sum())) * alens[multi]).astype(np.int64)
    att = np.full(R, -1, dtype=np.int64)
    has = alens > 0
    att[has] = avals[pick[has]]

    def weights(mask, classes):
        cnt = np.bincount(vid[mask], minlength=nv).astype(np.float64)
        w = 1.0 / (np.maximum(cnt, 1.0) * classes * nv)
        return np.where(mask, w[vid], 0.0).astype(np.float32)

    nb = np.asarray(batch.n_boxes, dtype=np.int64)
    w_obj = np.repeat((1.0 / (np.maximum(nb, 1) * nv)).astype(np.float32), nb)
    return {"lab_att": att, "lab_w_att": weights(att >= 0, 1), "lab_spa_bits": sb, "lab_w_spa": weights(sb != 0, 6),
            "lab_con_bits": cb, "lab_w_con": weights(cb != 0, 17), "lab_w_obj": w_obj}


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
