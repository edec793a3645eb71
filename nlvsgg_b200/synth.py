"""Deterministic synthetic Action-Genome-shaped inputs.

Produces the ``entry`` dict contract of the reference producers
(lib/assign_pseudo_label.py:1368-1382 for sgdet, lib/object_detector.py:126-139 for
predcls) and the ``gt_annotation`` list format of dataloader/wk_action_genome.py:281-292,
from a seed, on the CPU.  ``draw_fn`` is the union-mask rasteriser to use for
``spatial_masks`` (the CUDA op on the product path, the oracle's C restatement in tests).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

IMG_W, IMG_H = 480.0, 270.0
NUM_OBJ_CLASSES = 37  # background + 36 AG classes
AG_OBJECT_CLASSES = ("__background__ person bag bed blanket book box broom chair closetcabinet clothes "
                     "cupglassbottle dish door doorknob doorway floor food groceries laptop light medicine "
                     "mirror papernotebook phonecamera picture pillow refrigerator sandwich shelf shoe "
                     "sofacouch table television towel vacuum window").split()
AG_RELATIONS = ("lookingat notlookingat unsure above beneath infrontof behind onthesideof in carrying "
                "coveredby drinkingfrom eating haveitontheback holding leaningon lyingon notcontacting "
                "otherrelationship sittingon standingon touching twisting wearing wiping writingon").split()
AG_ATTENTION, AG_SPATIAL, AG_CONTACTING = AG_RELATIONS[0:3], AG_RELATIONS[3:9], AG_RELATIONS[9:]


def _create_dis(conf: float, idx: int) -> torch.Tensor:
    """36-way pseudo distribution: `conf` at idx, (1-conf)/35 elsewhere
    (lib/assign_pseudo_label.py:934-938)."""
    d = torch.full((36,), (1.0 - conf) / 35.0, dtype=torch.float32)
    d[idx] = conf
    return d


def synth_video(seed: int, frames: int = 20, mean_boxes: int = 6, mode: str = "sgdet",
                draw_fn: Optional[Callable[[np.ndarray, int], np.ndarray]] = None,
                empty_frame_prob: float = 0.0, fixed_boxes: Optional[int] = None,
                with_gt: bool = True, feat_dim: int = 2048, union_feat: bool = True, content_seed: Optional[int] = None):
    """One synthetic video.  Returns (entry, gt_annotation).

    entry tensors live on the CPU; callers move them.  ``spatial_masks`` is omitted when
    draw_fn is None (the product path rasterises on device from ``boxes``/``pair_idx``).
    content_seed: the two feature tensors are drawn from their own generator, so that videos of the same `seed` share their
    structure (boxes, pairs, labels: identical work) and differ in content — equal-size shards for data-parallel ranks.
    """
    g = torch.Generator().manual_seed(int(seed))
    gc = g if content_seed is None else torch.Generator().manual_seed(int(content_seed))

    def U(n=1):
        return torch.rand(n, generator=g)

    boxes, labels, scores, dist, pair_idx, im_idx = [], [], [], [], [], []
    gt = []
    att_gt, spa_gt, con_gt = [], [], []
    n_boxes = 0
    for f in range(frames):
        # an "empty" frame keeps its ground truth but contributes no detections / pairs
        # (weakly-supervised case handled by lib/transformer_wk.py:145-150,175-185)
        empty = empty_frame_prob > 0 and float(U()) < empty_frame_prob
        k = fixed_boxes if fixed_boxes is not None else int(torch.randint(max(2, mean_boxes - 2), mean_boxes + 3, (1,), generator=g))
        frame_first = n_boxes
        frame_gt = []
        for j in range(k):
            x1 = float(U()) * 0.6 * IMG_W
            y1 = float(U()) * 0.6 * IMG_H
            w = 20.0 + float(U()) * 0.35 * IMG_W
            h = 20.0 + float(U()) * 0.35 * IMG_H
            box = [float(f), x1, y1, min(x1 + w, IMG_W - 1), min(y1 + h, IMG_H - 1)]
            lab = 1 if j == 0 else int(torch.randint(2, 37, (1,), generator=g))
            sc = 1.0 if mode == "predcls" else 0.2 + 0.8 * float(U())
            if j > 0:
                a = [int(torch.randint(0, 3, (1,), generator=g))]
                ns = int(torch.randint(1, 3, (1,), generator=g))
                nc = int(torch.randint(1, 3, (1,), generator=g))
                s = torch.randperm(6, generator=g)[:ns].tolist()
                c = torch.randperm(17, generator=g)[:nc].tolist()
                jit = (torch.rand(4, generator=g) - 0.5) * 0.2 if mode != "predcls" else torch.zeros(4)
                bw, bh = box[3] - box[1], box[4] - box[2]
                gb = np.array([box[1] + float(jit[0]) * bw, box[2] + float(jit[1]) * bh,
                               box[3] + float(jit[2]) * bw, box[4] + float(jit[3]) * bh], dtype=np.float64)
                frame_gt.append({"bbox": gb, "class": lab,
                                 "attention_relationship": torch.tensor(a, dtype=torch.long),
                                 "spatial_relationship": torch.tensor(s, dtype=torch.long),
                                 "contacting_relationship": torch.tensor(c, dtype=torch.long)})
            else:
                frame_gt.append({"person_bbox": np.array([box[1:5]], dtype=np.float32),
                                 "frame": "synth.mp4/%06d.png" % f})
            if empty:
                continue
            boxes.append(box)
            labels.append(lab)
            scores.append(sc)
            dist.append(_create_dis(sc, lab - 1))
            if j > 0:
                pair_idx.append([frame_first, n_boxes])
                im_idx.append(f)
                att_gt.append(a); spa_gt.append(s); con_gt.append(c)
            n_boxes += 1
        gt.append(frame_gt)

    N, R = n_boxes, len(pair_idx)
    entry = {
        "boxes": torch.tensor(boxes, dtype=torch.float32).reshape(N, 5),
        "labels": torch.tensor(labels, dtype=torch.int64),
        "scores": torch.tensor(scores, dtype=torch.float32),
        "features": torch.relu(torch.randn(N, feat_dim, generator=gc)),
        "pair_idx": torch.tensor(pair_idx, dtype=torch.int64).reshape(R, 2),
        "im_idx": torch.tensor(im_idx, dtype=torch.float32 if mode == "predcls" else torch.int64),
        "attention_gt": att_gt, "spatial_gt": spa_gt, "contacting_gt": con_gt,
    }
    if mode != "predcls":
        entry["distribution"] = torch.stack(dist) if N else torch.zeros(0, 36)
    if union_feat:
        entry["union_feat"] = torch.relu(torch.randn(R, feat_dim, 7, 7, generator=gc))
    if draw_fn is not None:
        pr = pair_rois(entry)
        entry["spatial_masks"] = torch.from_numpy(draw_fn(pr, 27) - 0.5)
    return entry, (gt if with_gt else None)


def pair_rois(entry) -> np.ndarray:
    """f32[R,8] = (subject box, object box) rows fed to draw_union_boxes (lib/sttran.py:279-281)."""
    b, p = entry["boxes"], entry["pair_idx"]
    return torch.cat((b[p[:, 0], 1:], b[p[:, 1], 1:]), 1).numpy().astype(np.float32)


def make_state_dict(template: dict, seed: int = 0) -> dict:
    """Deterministic, init-order-independent weights for any model with `template`'s names/shapes.

    Keys are visited in sorted order with one CPU generator, so the reference model, the
    oracle and the CUDA model all load bit-identical tensors from the same (names, seed)."""
    g = torch.Generator().manual_seed(1000 + int(seed))
    out = {}
    for name in sorted(template.keys()):
        t = template[name]
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.int64)
        elif name.endswith("running_var"):
            out[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean"):
            out[name] = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".pe") or name.endswith("positional_encoder.pe"):
            out[name] = t.detach().clone()
        elif len(shape) == 1:
            is_norm_w = name.endswith("weight")  # 1-d weights are LayerNorm/BatchNorm scales
            out[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if is_norm_w else 0.05 * torch.randn(shape, generator=g)
        elif "embed" in name and len(shape) == 2:
            out[name] = torch.randn(shape, generator=g) if "position" not in name else torch.rand(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            out[name] = torch.randn(shape, generator=g) * (1.0 / np.sqrt(fan_in))
    return out


def synth_detections(seed: int, frames: int = 6, mean_boxes: int = 9, fmap_channels: int = 2048, fmap_hw=(17, 30)):
    """Raw detector output for the non-weakly-supervised sgdet TEST branch of lib/sttran.py:185-283 (before the per-class
    NMS): clusters of overlapping boxes that share an arg-max class, one strong person box per frame, class-distribution
    rows that sum to one, detector labels = arg-max + 1, and per-frame backbone maps for the union-box RoIAlign."""
    g = torch.Generator().manual_seed(int(seed))
    boxes, dists, labels = [], [], []
    for f in range(frames):
        k = int(torch.randint(max(3, mean_boxes - 3), mean_boxes + 4, (1,), generator=g))
        n_clusters = max(2, k // 2)
        centres = torch.rand(n_clusters, 2, generator=g) * torch.tensor([0.6 * IMG_W, 0.6 * IMG_H]) + 10.0
        sizes = 30.0 + torch.rand(n_clusters, 2, generator=g) * torch.tensor([0.3 * IMG_W, 0.3 * IMG_H])
        cls = torch.randint(1, 36, (n_clusters,), generator=g)          # column index in the 36-way distribution (0 = person)
        cls[0] = 0                                                       # the first cluster of every frame is the person
        for j in range(k):
            c = j % n_clusters
            jit = (torch.rand(4, generator=g) - 0.5) * 16.0              # +-8 px: IoU within a cluster straddles 0.6
            x1, y1 = float(centres[c, 0] + jit[0]), float(centres[c, 1] + jit[1])
            x2, y2 = x1 + float(sizes[c, 0] + jit[2]), y1 + float(sizes[c, 1] + jit[3])
            boxes.append([float(f), max(x1, 0.0), max(y1, 0.0), min(x2, IMG_W - 1.0), min(y2, IMG_H - 1.0)])
            logit = torch.randn(36, generator=g)
            logit[int(cls[c])] += 3.0 + 2.0 * float(torch.rand(1, generator=g))
            d = torch.softmax(logit, 0)
            dists.append(d)
            labels.append(int(torch.argmax(d)) + 1)
    N = len(boxes)
    return {
        "boxes": torch.tensor(boxes, dtype=torch.float32),
        "distribution": torch.stack(dists),
        "pred_labels": torch.tensor(labels, dtype=torch.int64),
        "features": torch.relu(torch.randn(N, 2048, generator=g)),
        "fmaps": torch.relu(torch.randn(frames, fmap_channels, fmap_hw[0], fmap_hw[1], generator=g)),
    }


def synth_pred(mode, seed, frames, k, empty_p, saturate):
    """A `pred` dict as lib/sttran.py would return it (CPU tensors) + the gt annotation list: random relation logits ->
    (logits | sigmoid | sigmoid); `saturate` produces exact ties between non-zero scores (saturated sigmoids, few distinct
    object scores) — the case the evaluator's canonical tie order exists for."""
    entry, gt = synth_video(seed, frames, k, mode, draw_fn=None, empty_frame_prob=empty_p, union_feat=False)
    g = torch.Generator().manual_seed(seed + 5000)
    R = entry["pair_idx"].shape[0]
    scale = 40.0 if saturate else 2.0
    pred = {k_: v for k_, v in entry.items() if torch.is_tensor(v)}
    pred["attention_distribution"] = torch.randn(R, 3, generator=g) * 2.0
    pred["spatial_distribution"] = torch.sigmoid(torch.randn(R, 6, generator=g) * scale)
    pred["contacting_distribution"] = torch.sigmoid(torch.randn(R, 17, generator=g) * scale)
    pred["pred_labels"] = entry["labels"].clone()
    pred["pred_scores"] = entry["scores"].clone()
    if saturate and mode != "predcls":
        pred["pred_scores"] = torch.round(pred["pred_scores"] * 4) / 4
    return pred, gt
