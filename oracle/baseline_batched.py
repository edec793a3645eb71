"""TEST INFRASTRUCTURE — the reference's OWN formulation of the STTran step on the host CPU (bench.py's --impl reference arm
and cpu_baseline leg).  Not imported by the product.

oracle/model.py states the mathematics segment by segment (an independent check); this file states it the way the
reference executes it, so that its CPU time is the reference's CPU time: frames padded to the longest frame and run as
ONE batched nn.MultiheadAttention call with a key_padding_mask (lib/transformer_wk.py:136-157), 2-frame windows padded
to 2l and run as one batch through the three decoder layers (:159-198), then the 'latter' scatter (:209-215).
Pinned against oracle/model.py (same numbers to 1e-5) and timed against the reference itself, imported through
oracle/ref_harness.py, in tests/test_cpu_baseline.py (build container only).
"""
from __future__ import annotations

import time
from typing import Dict, List

import torch
import torch.nn.functional as F

from oracle import model as omodel
from oracle.baseline import adamw_update

NHEAD = 8


def _mha(q, k, v, sd, p, mask):
    d = q.shape[-1]
    out, _ = F.multi_head_attention_forward(q, k, v, d, NHEAD, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"], None, None, False, 0.0,
                                            sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"], training=False,
                                            key_padding_mask=mask, need_weights=False)
    return out


def _encoder_layer(x, mask, sd, p):                      # lib/transformer.py:20-30, [l, b, d]
    x = F.layer_norm(x + _mha(x, x, x, sd, p + ".self_attn", mask), (x.shape[-1],), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    h = F.linear(F.relu(F.linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return F.layer_norm(x + h, (x.shape[-1],), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])


def _decoder_layer(x, pos, mask, sd, p):                 # lib/transformer.py:49-58
    t = F.layer_norm(x + _mha(x + pos, x + pos, x, sd, p + ".multihead2", mask), (x.shape[-1],), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])
    h = F.linear(F.relu(F.linear(t, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return t + h


def glocal_transformer_padded(features, im_idx, sd, prefix="glocal_transformer"):
    """lib/transformer_wk.py:130-217 (mode 'latter') with the reference's padding; the pad / gather index arithmetic is
    vectorised (the reference's python loops would only make this arm slower)."""
    fid = im_idx.to(torch.int64)
    R, d = features.shape
    b = int(fid[-1]) + 1
    cnt = torch.bincount(fid, minlength=b)
    l = int(cnt.max())
    start = torch.cumsum(cnt, 0) - cnt
    pos_in_frame = torch.arange(R) - start[fid]
    keep = torch.nonzero(cnt > 0).flatten()                                  # frames without pairs are dropped (:145-150)
    col = torch.full((b,), -1, dtype=torch.int64)
    col[keep] = torch.arange(len(keep))
    x = features.new_zeros(l, len(keep), d)
    x[pos_in_frame, col[fid]] = features
    mask = torch.arange(l)[None, :] >= cnt[keep][:, None]                    # [b', l] True = padding (bool masking)
    n_enc = omodel._num_layers(sd, f"{prefix}.local_attention.layers")
    n_dec = omodel._num_layers(sd, f"{prefix}.global_attention.layers")
    for i in range(n_enc):
        x = _encoder_layer(x, mask, sd, f"{prefix}.local_attention.layers.{i}")
    local = x[pos_in_frame, col[fid]]                                        # [R, d]
    win = torch.nonzero((cnt[:-1] + cnt[1:]) > 0).flatten() if b > 1 else torch.zeros(0, dtype=torch.int64)
    if len(win) == 0:
        return local
    wcol = torch.full((max(b - 1, 1),), -1, dtype=torch.int64)
    wcol[win] = torch.arange(len(win))
    g = features.new_zeros(2 * l, len(win), d)
    pe = features.new_zeros(2 * l, len(win), d)
    pw = sd[f"{prefix}.position_embedding.weight"]
    # a token of frame f sits in window f (first half, position p) and window f-1 (second half, position cnt[f-1] + p)
    in_first = (fid < b - 1) & (wcol[fid.clamp(max=max(b - 2, 0))] >= 0)
    r1 = torch.nonzero(in_first).flatten()
    g[pos_in_frame[r1], wcol[fid[r1]]] = local[r1]
    pe[pos_in_frame[r1], wcol[fid[r1]]] = pw[0]
    in_second = (fid > 0) & (wcol[(fid - 1).clamp(min=0)] >= 0)
    r2 = torch.nonzero(in_second).flatten()
    p2 = cnt[fid[r2] - 1] + pos_in_frame[r2]
    g[p2, wcol[fid[r2] - 1]] = local[r2]
    pe[p2, wcol[fid[r2] - 1]] = pw[1]
    wlen = (cnt[:-1] + cnt[1:])[win]
    gmask = torch.arange(2 * l)[None, :] >= wlen[:, None]
    for i in range(n_dec):
        g = _decoder_layer(g, pe, gmask, sd, f"{prefix}.global_attention.layers.{i}")
    out = torch.zeros_like(features)
    first = int(win[0]) if len(win) else -1
    out[r2] = g[p2, wcol[fid[r2] - 1]]                                       # 'latter': frame j+1 from window j (:213-215)
    r0 = torch.nonzero(fid == 0).flatten()
    if len(r0) and wcol[0] >= 0:
        out[r0] = g[pos_in_frame[r0], wcol[0]]                               # frame 0 from window 0's first half (:210-211)
    del first
    return out


def sttran_forward_padded(sd, entry, mode="sgdet", training=False):
    out = omodel.object_classifier(entry, sd, mode, training)
    tok = omodel.pair_tokens(entry, out["pred_labels"], sd, training)
    out.update(omodel.relation_heads(glocal_transformer_padded(tok, entry["im_idx"], sd), sd))
    return out


def cpu_train_step(sd: Dict[str, torch.Tensor], entries: List[dict], mode: str, opt_state: dict) -> float:
    """One iteration of tools/train_STTran.py:129-195 per video (forward, losses, backward), gradients averaged over the
    sample, then clip_grad_norm_(5) + lib/AdamW.py."""
    params = {k: v for k, v in sd.items() if v.is_floating_point() and "running_" not in k and not k.endswith(".pe") and "encoder_tran" not in k}
    for p in params.values():
        p.requires_grad_(True)
    total = 0.0
    for e in entries:
        pred = sttran_forward_padded(sd, e, mode, training=True)
        loss = omodel.training_loss(pred, e, mode) / len(entries)
        loss.backward()
        total += float(loss.detach())
    adamw_update(params, opt_state)
    return total


def time_cpu_steps(sd, entries, mode, steps: int, warmup: int):
    """Returns (seconds per step, frames per step)."""
    frames = sum(int(e["im_idx"].max().item()) + 1 if e["im_idx"].numel() else 0 for e in entries)
    st = {}
    for _ in range(warmup):
        cpu_train_step(sd, entries, mode, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_train_step(sd, entries, mode, st)
    return (time.perf_counter() - t0) / max(steps, 1), frames
