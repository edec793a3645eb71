"""TEST INFRASTRUCTURE — CPU restatement of the reference relation models (torch fp32).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (nlvsgg_b200/) never does.

Functional restatement over a plain ``state_dict`` (reference parameter names), written
segment-by-segment (one frame / one 2-frame window / one class sequence at a time) instead of
the reference's pad-to-max + key_padding_mask formulation, so it is an independent statement of
the same mathematics:

* object classifier ........ lib/sttran.py:88-92,173-184  (predcls / sgdet-wks branches)
* pair token (1936-d) ...... lib/sttran.py:381-399        (= lib/dsg_detr.py:517-532)
* spatial encoder layer .... lib/transformer.py:5-30      (post-norm MHA + FFN)
* temporal decoder layer ... lib/transformer.py:33-58     (q=k=x+pos, v=x; norm3; FFN residual, no final norm)
* frame/window plumbing .... lib/transformer_wk.py:130-217 (mode='latter'; empty frames/windows dropped;
                             single-frame video returns the local output, :187-188).  Masking is *bool*
                             masking (lib/transformer.py:144), see SURVEY.md §7 "Hard parts".
* relation heads ........... lib/sttran.py:404-409
* DSG-DETR local/global .... lib/dsg_detr.py:536-564 with PositionalEncoding :25-48
* training loss ............ tools/train_STTran.py:143-189 (CE obj + CE attention + BCE spatial/contact)

Third-party arithmetic (torch.nn.functional linear / layer_norm / batch_norm / conv2d / softmax)
is called from the installed torch, the same library the reference calls.
Pinned by oracle/validate_oracle.py against the reference itself (run in the build container)
and by tests/golden/*.pt fixtures produced by oracle/make_golden.py.

Dropout (p=0.1) cannot be RNG-matched and is treated as identity (eval, or train with p=0).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

D_MODEL, N_HEAD, HEAD_DIM = 1936, 8, 242


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def center_size(boxes: torch.Tensor) -> torch.Tensor:
    """(cx, cy, w, h) with the +1 pixel convention — lib/fpn/box_utils.py:51-63."""
    wh = boxes[:, 2:] - boxes[:, :2] + 1.0
    return torch.cat((boxes[:, :2] + 0.5 * wh, wh), 1)


def batch_norm(x, sd, prefix, training, momentum, update_running=True):
    """BatchNorm1d/2d as torch does it (batch stats + biased var in train; running stats in eval)."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and update_running:
        return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, 1e-5)
    if training:
        return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, 1e-5)
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], False, momentum, 1e-5)


def mha_segment(xq, xk, xv, w_in, b_in, w_out, b_out, nhead=N_HEAD, n_pad=0):
    """nn.MultiheadAttention on one unpadded segment: xq,xk,xv [L,d] -> [L,d].
    n_pad > 0: the torch-1.10.1 reading of the int key_padding_mask of lib/transformer_wk.py:154 — the frame's n_pad all-zero
    padded rows stay in the softmax as keys, with the mask value 1 ADDED to their logits."""
    L, d = xq.shape
    hd = d // nhead
    if n_pad:
        xk = torch.cat((xk, xk.new_zeros(n_pad, d)), 0)
        xv = torch.cat((xv, xv.new_zeros(n_pad, d)), 0)
    Lk = xk.shape[0]
    q = F.linear(xq, w_in[:d], b_in[:d]).view(L, nhead, hd).transpose(0, 1)
    k = F.linear(xk, w_in[d:2 * d], b_in[d:2 * d]).view(Lk, nhead, hd).transpose(0, 1)
    v = F.linear(xv, w_in[2 * d:], b_in[2 * d:]).view(Lk, nhead, hd).transpose(0, 1)
    logits = (q * (1.0 / math.sqrt(hd))) @ k.transpose(1, 2)
    if n_pad:
        logits = logits + torch.cat((logits.new_zeros(L), logits.new_ones(n_pad)))
    att = torch.softmax(logits, dim=-1)
    o = (att @ v).transpose(0, 1).reshape(L, d)
    return F.linear(o, w_out, b_out)


def encoder_layer(x, sd, p, attn="self_attn", n_pad=0):
    """Post-norm encoder layer on one segment (lib/transformer.py:20-30 / nn.TransformerEncoderLayer)."""
    a = mha_segment(x, x, x, sd[f"{p}.{attn}.in_proj_weight"], sd[f"{p}.{attn}.in_proj_bias"],
                    sd[f"{p}.{attn}.out_proj.weight"], sd[f"{p}.{attn}.out_proj.bias"], n_pad=n_pad)
    x = F.layer_norm(x + a, (x.shape[1],), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"])
    h = F.linear(F.relu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                 sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
    return F.layer_norm(x + h, (x.shape[1],), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"])


def decoder_layer(x, pos, sd, p):
    """Temporal decoder layer on one window (lib/transformer.py:49-58)."""
    a = mha_segment(x + pos, x + pos, x, sd[f"{p}.multihead2.in_proj_weight"], sd[f"{p}.multihead2.in_proj_bias"],
                    sd[f"{p}.multihead2.out_proj.weight"], sd[f"{p}.multihead2.out_proj.bias"])
    t = F.layer_norm(x + a, (x.shape[1],), sd[f"{p}.norm3.weight"], sd[f"{p}.norm3.bias"])
    h = F.linear(F.relu(F.linear(t, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                 sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
    return t + h


def _num_layers(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def glocal_transformer(features, im_idx, sd, prefix="glocal_transformer", additive_mask=False, mode="latter"):
    """transformer_wk.forward, mode='latter' (lib/transformer_wk.py:130-217).  additive_mask: the spatial encoder's int
    key_padding_mask as torch 1.10.1 read it (see mha_segment; one encoder layer, as the reference configures it)."""
    fid = im_idx.to(torch.int64)
    b = int(fid[-1]) + 1
    rows = [torch.nonzero(fid == f).flatten() for f in range(b)]
    n_enc = _num_layers(sd, f"{prefix}.local_attention.layers")
    n_dec = _num_layers(sd, f"{prefix}.global_attention.layers")
    assert not additive_mask or n_enc == 1
    lmax = max(r.numel() for r in rows)
    local = torch.zeros_like(features)
    for f in range(b):
        if rows[f].numel() == 0:
            continue
        x = features[rows[f]]
        for i in range(n_enc):
            x = encoder_layer(x, sd, f"{prefix}.local_attention.layers.{i}", n_pad=(lmax - rows[f].numel()) if additive_mask else 0)
        local[rows[f]] = x
    windows = [j for j in range(b - 1) if rows[j].numel() + rows[j + 1].numel() > 0]
    if len(windows) == 0:
        return local
    pe = sd[f"{prefix}.position_embedding.weight"]
    out = torch.zeros_like(features)
    seen = torch.zeros(features.shape[0])          # mode 'both' (lib/transformer_wk.py:197-207): a frame inside the video is the
    for j in windows:                              # mean of its two windows' outputs, the first / last frame have one window
        n0, n1 = rows[j].numel(), rows[j + 1].numel()
        x = torch.cat((local[rows[j]], local[rows[j + 1]]), 0)
        pos = torch.cat((pe[0].expand(n0, -1), pe[1].expand(n1, -1)), 0)
        for i in range(n_dec):
            x = decoder_layer(x, pos, sd, f"{prefix}.global_attention.layers.{i}")
        if mode == "both":
            both = torch.cat((rows[j], rows[j + 1]))
            out = out.index_add(0, both, x)
            seen[both] += 1
            continue
        if j == 0 and n0:
            out[rows[0]] = x[:n0]
        if n1:
            out[rows[j + 1]] = x[n0:]
    if mode == "both":
        out = out / seen.clamp(min=1)[:, None]
    return out


def object_classifier(entry, sd, mode, training, p="object_classifier", update_running=True):
    out = {}
    if mode == "predcls":
        out["pred_labels"] = entry["labels"]
        return out
    # sgdet with is_wks (train and test): lib/sttran.py:173-184
    obj_embed = entry["distribution"] @ sd[f"{p}.obj_embed.weight"]
    cs = center_size(entry["boxes"][:, 1:])
    pos = batch_norm(cs, sd, f"{p}.pos_embed.0", training, 0.01 / 10.0, update_running)
    pos = F.relu(F.linear(pos, sd[f"{p}.pos_embed.1.weight"], sd[f"{p}.pos_embed.1.bias"]))
    x = torch.cat((entry["features"], obj_embed, pos), 1)
    x = F.linear(x, sd[f"{p}.decoder_lin.0.weight"], sd[f"{p}.decoder_lin.0.bias"])
    x = F.relu(batch_norm(x, sd, f"{p}.decoder_lin.1", training, 0.1, update_running))
    out["distribution"] = F.linear(x, sd[f"{p}.decoder_lin.3.weight"], sd[f"{p}.decoder_lin.3.bias"])
    out["pred_labels"] = entry["labels"]
    out["pred_scores"] = entry["scores"]
    return out


def pair_tokens(entry, pred_labels, sd, training, update_running=True):
    """1936-d relation token per (human, object) pair — lib/sttran.py:381-399."""
    pi = entry["pair_idx"]
    subj = F.linear(entry["features"][pi[:, 0]], sd["subj_fc.weight"], sd["subj_fc.bias"])
    obj = F.linear(entry["features"][pi[:, 1]], sd["obj_fc.weight"], sd["obj_fc.bias"])
    u = F.conv2d(entry["union_feat"], sd["union_func1.weight"], sd["union_func1.bias"])
    m = F.conv2d(entry["spatial_masks"], sd["conv.0.weight"], sd["conv.0.bias"], stride=2, padding=3)
    m = batch_norm(F.relu(m), sd, "conv.2", training, 0.01, update_running)
    m = F.max_pool2d(m, kernel_size=3, stride=2, padding=1)
    m = F.conv2d(m, sd["conv.4.weight"], sd["conv.4.bias"], stride=1, padding=1)
    m = batch_norm(F.relu(m), sd, "conv.6", training, 0.01, update_running)
    vr = F.linear((u + m).reshape(-1, 256 * 7 * 7), sd["vr_fc.weight"], sd["vr_fc.bias"])
    semb = sd["obj_embed.weight"][pred_labels[pi[:, 0]]]
    oemb = sd["obj_embed2.weight"][pred_labels[pi[:, 1]]]
    return torch.cat((subj, obj, vr, semb, oemb), 1)


def relation_heads(x, sd):
    return {
        "attention_distribution": F.linear(x, sd["a_rel_compress.weight"], sd["a_rel_compress.bias"]),
        "spatial_distribution": torch.sigmoid(F.linear(x, sd["s_rel_compress.weight"], sd["s_rel_compress.bias"])),
        "contacting_distribution": torch.sigmoid(F.linear(x, sd["c_rel_compress.weight"], sd["c_rel_compress.bias"])),
    }


# --------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------
def sgcls_test_branch(entry, logits, union_feature_fn, draw_fn):
    """lib/sttran.py:105-170 (mode 'sgcls', eval): object labels from the classifier head, the human of every frame, the
    duplicate-class clean-up of the frame's most frequent label, (human, object) pairs, union boxes, union features through
    `union_feature_fn(frame_id, boxes_xyxy[n,4]) -> [n,2048,7,7]` (the un-vendored VinVL extractor of :159) and the masks.
    Returns a dict with the keys the reference writes."""
    boxes = entry["boxes"]
    box_idx = boxes[:, 0].long()
    b = int(box_idx[-1] + 1)
    dist = torch.softmax(logits[:, 1:], dim=1)                                     # :107
    scores, labels = torch.max(dist[:, 1:], dim=1)                                 # :108
    labels = labels + 2
    gidx = torch.arange(boxes.shape[0])
    human = torch.zeros(b, dtype=torch.int64)
    for i in range(b):                                                             # :115-117
        human[i] = gidx[box_idx == i][torch.argmax(dist[box_idx == i, 0])]
    labels[human] = 1
    scores[human] = dist[human, 0]
    for i in range(b):                                                             # :123-135
        present = boxes[:, 0] == i
        dup = torch.mode(labels[present])[0]
        if torch.sum(labels[present] == dup) > 0:
            pos = labels[present] == dup
            for j in torch.argsort(dist[present][pos][:, dup - 1])[:-1]:
                ci = gidx[present][pos][j]
                dist[ci, dup - 1] = 0
                labels[ci] = torch.argmax(dist[ci]) + 1
                scores[ci] = torch.max(dist[ci])
    im_idx, pair = [], []
    for j in range(b):                                                             # :138-143
        for m in gidx[box_idx == j][labels[box_idx == j] != 1]:
            im_idx.append(j)
            pair.append([int(human[j]), int(m)])
    pair = torch.tensor(pair, dtype=torch.int64).reshape(-1, 2)
    im_idx = torch.tensor(im_idx, dtype=torch.float)
    union = torch.cat((im_idx[:, None], torch.min(boxes[:, 1:3][pair[:, 0]], boxes[:, 1:3][pair[:, 1]]),
                       torch.max(boxes[:, 3:5][pair[:, 0]], boxes[:, 3:5][pair[:, 1]])), 1)   # :150-151
    feats = [union_feature_fn(f, union[union[:, 0] == f][:, 1:]) for f in range(b) if (union[:, 0] == f).any()]   # :154-160
    rois = torch.cat((boxes[pair[:, 0], 1:], boxes[pair[:, 1], 1:]), 1).numpy()
    return {"distribution": dist, "pred_scores": scores, "pred_labels": labels, "pair_idx": pair, "im_idx": im_idx,
            "union_box": union, "union_feat": torch.cat(feats), "human_idx": human,
            "spatial_masks": torch.from_numpy(draw_fn(rois.astype("float32"), 27) - 0.5)}


def sttran_forward(sd: Dict[str, torch.Tensor], entry: dict, mode: str = "sgdet", training: bool = False,
                   update_running: bool = True, return_tokens: bool = False, additive_mask: bool = False) -> dict:
    """lib/sttran.py:375-411.  Returns a new dict with the keys the reference adds/overwrites."""
    out = object_classifier(entry, sd, mode, training, update_running=update_running)
    tok = pair_tokens(entry, out["pred_labels"], sd, training, update_running)
    g = glocal_transformer(tok, entry["im_idx"], sd, additive_mask=additive_mask)
    out.update(relation_heads(g, sd))
    if return_tokens:
        out["rel_features"], out["global_output"] = tok, g
    return out


def sinusoidal_pe(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding buffer — lib/dsg_detr.py:31-36."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def dsg_forward(sd: Dict[str, torch.Tensor], entry: dict, mode: str = "sgdet", training: bool = False,
                update_running: bool = True, return_tokens: bool = False) -> dict:
    """lib/dsg_detr.py:514-572 for predcls-without-tracks is not covered; sgdet (the tools' mode) is.

    local  : one post-norm encoder layer per frame group (frames keyed by the object box's frame id, :536-543)
    global : three encoder layers over per-object-class sequences of the whole video, tokens get
             pe[rank of the pair's subject box among the sequence's distinct subject boxes] (:545-559; sgdet only)
    """
    out = object_classifier(entry, sd, mode, training, update_running=update_running)
    tok = pair_tokens(entry, out["pred_labels"], sd, training, update_running)
    pi = entry["pair_idx"]
    frame_of = entry["boxes"][pi[:, 1], 0]
    local = torch.zeros_like(tok)
    for f in torch.unique(frame_of):
        r = torch.nonzero(frame_of == f).flatten()
        local[r] = encoder_layer(tok[r], sd, "local_transformer.layers.0")
    obj_class = out["pred_labels"][pi[:, 1]]
    pe = sd["positional_encoder.pe"][0] if "positional_encoder.pe" in sd else sinusoidal_pe(400, tok.shape[1])
    glob = torch.zeros_like(tok)
    for c in torch.unique(obj_class):
        r = torch.nonzero(obj_class == c).flatten()
        x = local[r]
        if mode == "sgdet":
            _, inv = torch.unique(pi[r, 0], sorted=True, return_inverse=True)
            # the reference lays ranks out as [0]*count0 + [1]*count1 ... in sequence order (:553-556);
            # pairs are frame-sorted so this equals the rank of each row's subject box
            counts = torch.bincount(inv)
            rank = torch.repeat_interleave(torch.arange(len(counts)), counts)
            x = x + pe[rank]
        else:
            x = x + pe[: x.shape[0]]
        for i in range(3):
            x = encoder_layer(x, sd, f"global_transformer.layers.{i}")
        glob[r] = x
    out.update(relation_heads(glob, sd))
    if return_tokens:
        out["rel_features"], out["global_output"] = tok, glob
    return out


# --------------------------------------------------------------------------------------
# loss (tools/train_STTran.py:143-189, bce_loss=True)
# --------------------------------------------------------------------------------------
def build_labels(entry):
    att = torch.tensor([a[0] for a in entry["attention_gt"] if len(a) > 0], dtype=torch.int64)
    att_mask = torch.tensor([len(a) > 0 for a in entry["attention_gt"]], dtype=torch.bool)
    R = len(entry["spatial_gt"])
    spa = torch.zeros(R, 6)
    con = torch.zeros(R, 17)
    for i in range(R):
        spa[i, entry["spatial_gt"][i]] = 1.0
        con[i, entry["contacting_gt"][i]] = 1.0
    return att, att_mask, spa, con


def training_loss(pred: dict, entry: dict, mode: str = "sgdet") -> torch.Tensor:
    """Sum of the (up to) four mean-reduced losses.  Multi-label attention picks the first label
    (the reference draws one at random, train_STTran.py:152-153; synthetic data has one)."""
    att, att_mask, spa, con = build_labels(entry)
    losses = []
    if mode != "predcls":
        losses.append(F.cross_entropy(pred["distribution"], entry["labels"]))
    else:
        pass
    if int(att_mask.sum()) > 0:
        losses.append(F.cross_entropy(pred["attention_distribution"][att_mask], att))
    sm = (spa > 0).sum(-1) != 0
    cm = (con > 0).sum(-1) != 0
    if int(sm.sum()) > 0:
        losses.append(F.binary_cross_entropy(pred["spatial_distribution"][sm], spa[sm]))
    if int(cm.sum()) > 0:
        losses.append(F.binary_cross_entropy(pred["contacting_distribution"][cm], con[cm]))
    return sum(losses)
