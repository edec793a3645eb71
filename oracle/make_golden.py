"""TEST INFRASTRUCTURE — writes tests/golden/* by running the reference itself.

Build container only (needs /root/reference).  Inputs and weights are *not* stored (103 M
parameters): they are regenerated bit-identically from seeds by nlvsgg_b200.synth
(`synth_video`, `make_state_dict`, CPU generators).  Stored per case: the generator
arguments and the reference's outputs (and, for training cases, the loss, the updated BN
running statistics and gradient digests).

    python -m oracle.make_golden
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from nlvsgg_b200 import synth
from oracle import cref, ref_harness as H

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

STTRAN_CASES = [
    # name, mode, seed, frames, mean_boxes, empty_frame_prob, training
    ("sttran_predcls_c1", "predcls", 0, 20, 6, 0.0, False),   # BASELINE config C1
    ("sttran_sgdet_eval", "sgdet", 1, 20, 6, 0.0, False),
    ("sttran_sgdet_gaps", "sgdet", 2, 12, 5, 0.3, False),     # empty frames (transformer_wk.py:145-150)
    ("sttran_sgdet_1frame", "sgdet", 4, 1, 6, 0.0, False),    # single frame -> local output only (:187-188)
    ("sttran_sgdet_train", "sgdet", 3, 8, 5, 0.0, True),      # train-mode BN + loss + gradients
    ("sttran_sgdet_train_gaps", "sgdet", 5, 9, 5, 0.4, True),
]
DSG_CASES = [
    ("dsg_sgdet_eval", "sgdet", 6, 12, 6, False),
    ("dsg_sgdet_train", "sgdet", 7, 8, 5, True),
]
OUT_KEYS = ("attention_distribution", "spatial_distribution", "contacting_distribution", "distribution")


def _clone_entry(e):
    return {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in e.items()}


def _no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
        if isinstance(mod, torch.nn.TransformerEncoderLayer):
            mod.dropout.p = mod.dropout1.p = mod.dropout2.p = 0.0


def grad_digest(g: torch.Tensor) -> dict:
    """Small, order-independent-enough summary of a gradient tensor (full tensor if it is small)."""
    g = g.detach().double().flatten()
    d = {"sum": g.sum().item(), "abs_sum": g.abs().sum().item(), "sq_sum": (g * g).sum().item(),
         "head": g[:64].float().clone()}
    if g.numel() <= 4096:
        d["full"] = g.float().clone()
    return d


def reference_loss(ref, pred, entry):
    """tools/train_STTran.py:143-189 with bce_loss=True, executed with the reference's own ops."""
    import torch.nn as nn
    ce, bce = nn.CrossEntropyLoss(), nn.BCELoss()
    att_mask = torch.tensor([len(i) > 0 for i in pred["attention_gt"]])
    att_label = torch.tensor([int(i[0]) for i in pred["attention_gt"] if len(i) >= 1], dtype=torch.int64)
    R = len(pred["spatial_gt"])
    spa = torch.zeros(R, 6)
    con = torch.zeros(R, 17)
    for i in range(R):
        spa[i, pred["spatial_gt"][i]] = 1.0
        con[i, pred["contacting_gt"][i]] = 1.0
    losses = {"object_loss": ce(pred["distribution"], pred["labels"])}
    if att_mask.sum().item() > 0:
        losses["attention_relation_loss"] = ce(pred["attention_distribution"][att_mask], att_label)
    sm = (spa > 0).sum(-1) != 0
    cm = (con > 0).sum(-1) != 0
    if sm.sum().item() > 0:
        losses["spatial_relation_loss"] = bce(pred["spatial_distribution"][sm], spa[sm])
    if cm.sum().item() > 0:
        losses["contact_relation_loss"] = bce(pred["contacting_distribution"][cm], con[cm])
    return sum(losses.values())


def run_model_case(ref, builder, fwd_mode, name, mode, seed, frames, k, empty_p, training):
    m = builder(ref, mode)
    sd = synth.make_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    _no_dropout(m)
    m.train(training)
    entry, _ = synth.synth_video(seed, frames, k, mode, draw_fn=cref.draw_union_boxes, empty_frame_prob=empty_p)
    case = {"name": name, "model": fwd_mode, "mode": mode, "seed": seed, "frames": frames, "mean_boxes": k,
            "empty_frame_prob": empty_p, "training": training,
            "n_boxes": int(entry["boxes"].shape[0]), "n_pairs": int(entry["pair_idx"].shape[0])}
    e = _clone_entry(entry)
    if training:
        m(e)
        loss = reference_loss(ref, e, entry)
        loss.backward()
        case["loss"] = float(loss.item())
        case["grads"] = {n: grad_digest(p.grad) for n, p in m.named_parameters() if p.grad is not None}
        case["no_grad_params"] = [n for n, p in m.named_parameters() if p.grad is None]
        after = m.state_dict()
        case["running"] = {n: after[n].detach().clone() for n in after if "running_" in n}
    else:
        with torch.no_grad():
            m(e)
    case["outputs"] = {kk: e[kk].detach().clone() for kk in OUT_KEYS if kk in e and torch.is_tensor(e[kk])}
    torch.save(case, os.path.join(GOLDEN, name + ".pt"))
    print(f"wrote {name}: boxes={case['n_boxes']} pairs={case['n_pairs']}" + (f" loss={case['loss']:.6f}" if training else ""))


def native_cases(ref):
    rng = np.random.default_rng(11)
    b = rng.uniform(0, 400, (64, 8)).astype(np.float32)
    b[:, 2:4] = b[:, 0:2] + rng.uniform(1, 200, (64, 2)).astype(np.float32)
    b[:, 6:8] = b[:, 4:6] + rng.uniform(1, 200, (64, 2)).astype(np.float32)
    b[0] = [10, 10, 50, 50, 10, 10, 50, 50]          # identical boxes
    b[1] = [0, 0, 100, 100, 25, 25, 75, 75]          # nested
    b[2] = [0, 0, 10, 10, 200, 200, 210, 230]        # disjoint
    out = ref.draw_rectangles.draw_union_boxes(b, 27)
    np.savez_compressed(os.path.join(GOLDEN, "native_draw_union_boxes.npz"), box_pairs=b, out=out)
    x = rng.uniform(0, 100, (40, 4)); x[:, 2:] += x[:, :2]
    y = rng.uniform(0, 100, (30, 4)); y[:, 2:] += y[:, :2]
    x[0] = y[0]                                       # IoU exactly 1
    x[1] = [0, 0, 9, 9]; y[1] = [5, 0, 14, 9]         # +1 convention: 5x10 / (100+100-50) = 1/3
    x[2] = [0, 0, 1, 1]; y[2] = [50, 50, 60, 60]      # disjoint -> 0
    ov = ref.bbox.bbox_overlaps(x, y)
    np.savez_compressed(os.path.join(GOLDEN, "native_bbox_overlaps.npz"), boxes=x, query=y, out=ov)
    print("wrote native_draw_union_boxes, native_bbox_overlaps")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = H.load_reference()
    native_cases(ref)
    for (name, mode, seed, frames, k, ep, tr) in STTRAN_CASES:
        run_model_case(ref, H.build_reference_sttran, "sttran", name, mode, seed, frames, k, ep, tr)
    for (name, mode, seed, frames, k, tr) in DSG_CASES:
        run_model_case(ref, H.build_reference_dsg, "dsg_detr", name, mode, seed, frames, k, 0.0, tr)
    try:
        from oracle import make_golden_eval
        make_golden_eval.main(ref)
    except ImportError:
        pass


if __name__ == "__main__":
    main()
