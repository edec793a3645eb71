"""Drop-in for lib/evaluation_recall.py:SceneGraphEvaluator (same constructor, methods and ``result_dict`` layout).

The per-frame python/numpy work of the reference (:402-465 — ~20 device->host syncs and an O(G * 3P) python loop
per frame) is replaced by one launch of the ``nlv_recall_match`` kernel per call (one CTA per frame, any number of
videos per launch); the host only packs the ground-truth python structures into flat integer arrays and turns the
kernel's integer match sets into the reference's per-frame floats, in the reference's order.  Tie order is the
canonical one: see csrc/eval.cu.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn

from .. import _C
from ..ops import _ptr, _stream

KS = (10, 20, 50)
_LIMITS = None


def _limits():
    global _LIMITS
    if _LIMITS is None:
        p, g, gb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _C.lib().nlv_recall_limits(ctypes.byref(p), ctypes.byref(g), ctypes.byref(gb))
        _LIMITS = (p.value, g.value, gb.value)
    return _LIMITS


class PackedGT:
    """Ground truth of a list of frames flattened into the kernel's arrays (host side, numpy)."""

    def __init__(self):
        self.rel, self.cls, self.box = [], [], []
        self.rel_off, self.box_off = [0], [0]

    def add_frame(self, frame_gt, idx_att, idx_spa, idx_con):
        """evaluation_recall.py:404-419: person = box 0 (class 1); attention/contacting triplets are
        (human, object, p), spatial ones (object, human, p)."""
        nb = len(frame_gt)
        boxes = np.zeros((nb, 4), dtype=np.float64)
        cls = np.zeros(nb, dtype=np.int32)
        cls[0] = 1
        boxes[0] = np.asarray(frame_gt[0]["person_bbox"], dtype=np.float64).reshape(-1)[:4]
        rel = []
        for m, obj in enumerate(frame_gt[1:]):
            boxes[m + 1] = np.asarray(obj["bbox"], dtype=np.float64)
            cls[m + 1] = int(obj["class"])
            a = np.asarray(obj["attention_relationship"]).reshape(-1)
            rel.append((0, m + 1, idx_att[int(a[0])]))
            for s in np.asarray(obj["spatial_relationship"]).reshape(-1).tolist():
                rel.append((m + 1, 0, idx_spa[int(s)]))
            for c in np.asarray(obj["contacting_relationship"]).reshape(-1).tolist():
                rel.append((0, m + 1, idx_con[int(c)]))
        self.rel.append(np.asarray(rel, dtype=np.int32).reshape(-1, 3))
        self.cls.append(cls)
        self.box.append(boxes.astype(np.float32))        # rounded to f32 exactly as :765 does
        self.rel_off.append(self.rel_off[-1] + len(rel))
        self.box_off.append(self.box_off[-1] + nb)


def recall_match(pair_off, gt: PackedGT, pair_sub, pair_obj, att, spa, con, obj_scores, pred_cls, pred_boxes):
    """Launch the kernel; returns u32[F,3,3,8] match sets (numpy)."""
    dev = att.device
    F = len(pair_off) - 1
    pmax, gmax, gbmax = _limits()
    po = np.asarray(pair_off, dtype=np.int32)
    ro, bo = np.asarray(gt.rel_off, dtype=np.int32), np.asarray(gt.box_off, dtype=np.int32)
    if F and (np.diff(po).max(initial=0) > pmax or np.diff(ro).max(initial=0) > gmax or np.diff(bo).max(initial=0) > gbmax):
        raise RuntimeError(f"recall_match: a frame exceeds the kernel limits (pairs<={pmax}, gt relations<={gmax}, gt boxes<={gbmax})")
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev, non_blocking=True)
    d_po, d_ro, d_bo = t(po, np.int32), t(ro, np.int32), t(bo, np.int32)
    d_rel = t(np.concatenate(gt.rel) if gt.rel else np.zeros((0, 3)), np.int32)
    d_cls = t(np.concatenate(gt.cls) if gt.cls else np.zeros(0), np.int32)
    d_box = t(np.concatenate(gt.box) if gt.box else np.zeros((0, 4)), np.float32)
    out = torch.empty(F, 3, 3, 8, device=dev, dtype=torch.int32)
    _C.check(_C.lib().nlv_recall_match(F, _ptr(d_po), _ptr(d_ro), _ptr(d_bo), _ptr(pair_sub), _ptr(pair_obj), _ptr(att),
                                       _ptr(spa), _ptr(con), _ptr(obj_scores), _ptr(pred_cls), _ptr(pred_boxes), _ptr(d_rel),
                                       _ptr(d_cls), _ptr(d_box), _ptr(out), _stream()), "recall_match")
    return out.cpu().numpy().view(np.uint32)


def _bits(words: np.ndarray) -> List[int]:
    """Indices of the set bits of a u32[8] mask, ascending."""
    idx = []
    for w in range(8):
        x = int(words[w])
        while x:
            b = x & -x
            idx.append(32 * w + b.bit_length() - 1)
            x ^= b
    return idx


class SceneGraphEvaluator:
    def __init__(self, mode, AG_object_classes, AG_all_predicates, AG_attention_predicates, AG_spatial_predicates,
                 AG_contacting_predicates, iou_threshold=0.5, constraint=False, semithreshold=None):
        self.result_dict = {}
        self.mode = mode
        self.subject_category = 1
        assert iou_threshold == 0.5, "the kernel implements the reference's fixed IoU threshold 0.5 (evaluation_recall.py:231)"
        self.iou_threshold, self.constraint, self.semithreshold = iou_threshold, constraint, semithreshold
        self.AG_object_classes, self.AG_all_predicates = AG_object_classes, list(AG_all_predicates)
        self.AG_attention_predicates, self.AG_spatial_predicates = list(AG_attention_predicates), list(AG_spatial_predicates)
        self.AG_contacting_predicates = list(AG_contacting_predicates)
        self.num_rel = len(self.AG_all_predicates)
        self._ia = [self.AG_all_predicates.index(p) for p in self.AG_attention_predicates]
        self._is = [self.AG_all_predicates.index(p) for p in self.AG_spatial_predicates]
        self._ic = [self.AG_all_predicates.index(p) for p in self.AG_contacting_predicates]
        assert self._ia == [0, 1, 2] and self._is == list(range(3, 9)) and self._ic == list(range(9, 26)), \
            "predicate blocks must be attention|spatial|contacting = [0:3|3:9|9:26] (dataloader/wk_action_genome.py:85-87)"

    # ---- containers (lib/evaluation_recall.py:375-380 and the per-class register_container methods) ----
    def register_container(self):
        m = self.mode
        for t in ("_recall", "_recall_nogc", "_semi_recall"):
            self.result_dict[m + t] = {k: [] for k in KS}
        for t in ("_mean_recall", "_ng_mean_recall"):
            self.result_dict[m + t] = {k: 0.0 for k in KS}
            self.result_dict[m + t + "_collect"] = {k: [[] for _ in range(self.num_rel)] for k in KS}
            self.result_dict[m + t + "_list"] = {k: [] for k in KS}

    # ---- evaluation ----
    def evaluate_scene_graph(self, gt, pred):
        """One video (the reference signature).  Mutates pred['attention_distribution'] (softmax), as :400 does."""
        self.evaluate_videos([(gt, pred)])

    def evaluate_videos(self, items: Sequence[tuple]):
        """Any number of (gt, pred) videos in ONE kernel launch; results are appended in input order."""
        packed = PackedGT()
        pair_off = [0]
        subs, objs, atts, spas, cons, oscs, clss, boxs = [], [], [], [], [], [], [], []
        box_base = 0
        dev = None
        for gt, pred in items:
            pred["attention_distribution"] = nn.functional.softmax(pred["attention_distribution"], dim=1)
            dev = pred["attention_distribution"].device
            if dev.type != "cuda":
                raise RuntimeError("SceneGraphEvaluator (nlvsgg_b200) needs CUDA tensors; there is no CPU fallback")
            im_idx = pred["im_idx"].detach().cpu().numpy().astype(np.int64)
            nf = len(gt)
            cnt = np.bincount(im_idx, minlength=nf) if len(im_idx) else np.zeros(nf, dtype=np.int64)
            assert len(cnt) == nf and (len(im_idx) == 0 or np.all(np.diff(im_idx) >= 0)), "im_idx must be sorted frame ids"
            for f in range(nf):
                packed.add_frame(gt[f], self._ia, self._is, self._ic)
                pair_off.append(pair_off[-1] + int(cnt[f]))
            pi = pred["pair_idx"].to(torch.int32) + box_base
            subs.append(pi[:, 0]); objs.append(pi[:, 1])
            atts.append(pred["attention_distribution"].float()); spas.append(pred["spatial_distribution"].float())
            cons.append(pred["contacting_distribution"].float())
            if self.mode == "predcls":
                clss.append(pred["labels"].to(torch.int32)); oscs.append(pred["scores"].float())
            else:
                clss.append(pred["pred_labels"].to(torch.int32)); oscs.append(pred["pred_scores"].float())
            boxs.append(pred["boxes"][:, 1:].float())
            box_base += int(pred["boxes"].shape[0])
        cat = lambda ts: (ts[0] if len(ts) == 1 else torch.cat(ts, 0)).contiguous()
        masks = recall_match(pair_off, packed, cat(subs), cat(objs), cat(atts), cat(spas), cat(cons), cat(oscs), cat(clss),
                             cat(boxs))
        self._book(masks, packed)

    def _book(self, masks: np.ndarray, packed: PackedGT):
        m = self.mode
        for f in range(masks.shape[0]):
            rel = packed.rel[f]
            G = rel.shape[0]
            for pi, key in enumerate(("_recall", "_recall_nogc", "_semi_recall")):
                for ki, k in enumerate(KS):
                    n = int(sum(bin(int(w)).count("1") for w in masks[f, pi, ki]))
                    self.result_dict[m + key][k].append(float(n) / float(G))
            for pi, key in ((0, "_mean_recall"), (1, "_ng_mean_recall")):     # :69-87 / :146-165
                for ki, k in enumerate(KS):
                    hit = [0] * self.num_rel
                    cnt = [0] * self.num_rel
                    for g in range(G):
                        cnt[int(rel[g, 2])] += 1
                        cnt[0] += 1
                    for g in _bits(masks[f, pi, ki]):
                        hit[int(rel[g, 2])] += 1
                        hit[0] += 1
                    for n in range(self.num_rel):
                        if cnt[n] > 0:
                            self.result_dict[m + key + "_collect"][k][n].append(float(hit[n] / cnt[n]))

    def calculate_mean_recall(self):
        for t in ("_mean_recall", "_ng_mean_recall"):      # :89-109 / :167-187
            for k in KS:
                s = 0
                for n in range(self.num_rel):
                    lst = self.result_dict[self.mode + t + "_collect"][k][n]
                    r = 0.0 if len(lst) == 0 else np.mean(lst)
                    self.result_dict[self.mode + t + "_list"][k].append(r)
                    s += r
                self.result_dict[self.mode + t][k] = s / float(self.num_rel)

    # ---- printing (string formats of :41-67, :130-144, :201-207, :249-255, :313-319) ----
    def print_stats(self, logger):
        m = self.mode
        logger.info("======================" + m + "============================")
        s = ""
        for key, title, tag in (("_recall", "Recall(Main).", "  R @ %d: %.4f; "), ("_recall_nogc", "No Graph Constraint Recall(Main).", "  R @ %d: %.4f; "),
                                ("_semi_recall", "Semi Recall.", "  R @ %d: %.4f; ")):
            s += "SGG eval: "
            for k, v in self.result_dict[m + key].items():
                s += tag % (k, np.mean(v))
            s += " for mode=%s, type=%s" % (m, title) + "\n"
        s += "SGG eval: "
        for k, v in self.result_dict[m + "_mean_recall"].items():
            s += " mR @ %d: %.4f; " % (k, float(v))
        s += " for mode=%s, type=Mean Recall." % m + "\n"
        for k in KS:
            s += "Per-class recall@%d: \n" % k
            for n, r in zip(self.AG_all_predicates, self.result_dict[m + "_mean_recall_list"][k]):
                s += "({}:{:.4f}) ".format(str(n), r)
            s += "\n"
        s += "\nSGG eval: "
        for k, v in self.result_dict[m + "_ng_mean_recall"].items():
            s += "ng-mR @ %d: %.4f; " % (k, float(v))
        s += " for mode=%s, type=No Graph Constraint Mean Recall." % m + "\n"
        s += "----------------------- Details ------------------------\n"
        for n, r in zip(self.AG_all_predicates, self.result_dict[m + "_ng_mean_recall_list"][50]):
            s += "({}:{:.4f}) ".format(str(n), r)
        s += "\n--------------------------------------------------------\n"
        logger.info(s)
