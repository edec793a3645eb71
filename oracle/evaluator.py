"""TEST INFRASTRUCTURE — numpy restatement of the reference Recall@K evaluator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Restates, per frame (lib/evaluation_recall.py):
  * ground-truth triplets .......... :402-425   (attention (human,obj,p); spatial (obj,human,p); contacting (human,obj,p))
  * prediction rows ................ :429-442   3P rows = pairs | reversed pairs | pairs, 26 block-diagonal scores
  * with-constraint ................ :209-236   argmax / max over the 26 columns
  * no-constraint .................. :321-353   top-100 of f32(obj_s*obj_s) * rel_scores over 3P x 26
  * semi-constraint ................ :257-302   attention argmax; spatial / contacting entries > 0.5
  * evaluate_recall ................ :630-695   sort by score product (desc), match, R@K = |U matches[:K]| / G
  * _compute_pred_matches .......... :731-773   class-equality x (IoU_sub >= .5 & IoU_obj >= .5), boxes rounded to f32, IoU in f64 (+1)
  * mean-recall collectors ......... :69-109, :146-187, calculate_mean_recall :89-109

Tie order.  The reference sorts with NumPy's default (unstable) argsort.  Canonical rule: the
no-constraint selection sorts descending with ties by ascending flat index (== np.argsort(-x, kind='stable'));
evaluate_recall sorts descending with ties by DESCENDING position (== x.argsort(kind='stable')[::-1]).
oracle/validate_oracle.py checks this restatement against the reference both as shipped and with those two
one-word stable-sort patches.

Output of `frame_matches`: for each of the three protocols and K in (10,20,50) the sorted tuple of matched GT
indices — integers only.  `Evaluator` mirrors SceneGraphEvaluator's result_dict bookkeeping on top of it.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from oracle import cref

KS = (10, 20, 50)


def build_frame_gt(frame_gt: list, att_names: Sequence[str], spa_names: Sequence[str], con_names: Sequence[str],
                   all_names: Sequence[str]):
    """-> gt_boxes f64[Gb,4], gt_classes i64[Gb], gt_rels i64[G,3] (sub, obj, predicate)."""
    gb = np.zeros((len(frame_gt), 4), dtype=np.float64)
    gc = np.zeros(len(frame_gt), dtype=np.int64)
    gc[0] = 1
    gb[0] = np.asarray(frame_gt[0]["person_bbox"], dtype=np.float64).reshape(-1)[:4]
    rels = []
    for m, obj in enumerate(frame_gt[1:]):
        gb[m + 1] = np.asarray(obj["bbox"], dtype=np.float64)
        gc[m + 1] = int(obj["class"])
        a = np.asarray(obj["attention_relationship"]).reshape(-1)
        rels.append((0, m + 1, all_names.index(att_names[int(a[0]) if a.size == 1 else int(a)])))
        for s in np.asarray(obj["spatial_relationship"]).reshape(-1).tolist():
            rels.append((m + 1, 0, all_names.index(spa_names[int(s)])))
        for c in np.asarray(obj["contacting_relationship"]).reshape(-1).tolist():
            rels.append((0, m + 1, all_names.index(con_names[int(c)])))
    return gb, gc, np.asarray(rels, dtype=np.int64).reshape(-1, 3)


def _match_sets(gt_rels, gt_boxes, gt_classes, cand_sub, cand_obj, cand_pred, cand_score, pred_boxes, pred_classes,
                obj_scores):
    """evaluate_recall + _compute_pred_matches: returns, in rank order, the list of matched-GT index lists."""
    n = len(cand_pred)
    if n == 0:
        return []
    trip_score = (obj_scores[cand_sub].astype(np.float64) * obj_scores[cand_obj].astype(np.float64)) * cand_score
    order = np.argsort(trip_score, kind="stable")[::-1]
    gt_trip = np.column_stack((gt_classes[gt_rels[:, 0]], gt_rels[:, 2], gt_classes[gt_rels[:, 1]]))
    gtb32 = gt_boxes.astype(np.float32).astype(np.float64)
    pb32 = pred_boxes.astype(np.float32).astype(np.float64)
    out = []
    for i in order:
        s, o, p = cand_sub[i], cand_obj[i], cand_pred[i]
        hits = []
        for g in range(gt_rels.shape[0]):
            if gt_trip[g, 0] == pred_classes[s] and gt_trip[g, 1] == p and gt_trip[g, 2] == pred_classes[o]:
                iou_s = cref.bbox_overlaps(gtb32[gt_rels[g, 0]][None], pb32[s][None])[0, 0]
                iou_o = cref.bbox_overlaps(gtb32[gt_rels[g, 1]][None], pb32[o][None])[0, 0]
                if iou_s >= 0.5 and iou_o >= 0.5:
                    hits.append(g)
        out.append(hits)
    return out


def frame_matches(gt_boxes, gt_classes, gt_rels, pairs, att, spa, con, pred_boxes, pred_classes, obj_scores):
    """pairs i64[P,2]; att f32[P,3] (already softmaxed), spa f32[P,6], con f32[P,17];
    pred_boxes f32[N,4]; pred_classes i64[N]; obj_scores f32[N].
    Returns {protocol: [match lists in rank order]} for 'with', 'nogc', 'semi'."""
    P = pairs.shape[0]
    rows_sub = np.concatenate((pairs[:, 0], pairs[:, 1], pairs[:, 0])).astype(np.int64)
    rows_obj = np.concatenate((pairs[:, 1], pairs[:, 0], pairs[:, 1])).astype(np.int64)
    scores = np.zeros((3 * P, 26), dtype=np.float64)
    scores[:P, 0:3] = att
    scores[P:2 * P, 3:9] = spa
    scores[2 * P:, 9:26] = con
    res = {}
    common = (pred_boxes, pred_classes, obj_scores)
    # with constraint
    res["with"] = _match_sets(gt_rels, gt_boxes, gt_classes, rows_sub, rows_obj, scores.argmax(1), scores.max(1), *common) \
        if P else []
    # no constraint: top-100 of f32(obj*obj) * rel
    if P:
        per_rel = (obj_scores[rows_sub] * obj_scores[rows_obj]).astype(np.float32)       # float32 product, as numpy computes it
        overall = per_rel[:, None].astype(np.float64) * scores
        flat = np.argsort(-overall.ravel(), kind="stable")[:100]
        r, c = np.unravel_index(flat, overall.shape)
        res["nogc"] = _match_sets(gt_rels, gt_boxes, gt_classes, rows_sub[r], rows_obj[r], c, scores[r, c], *common)
    else:
        res["nogc"] = []
    # semi constraint
    cs, co, cp, cv = [], [], [], []
    for i in range(3 * P):
        row = scores[i]
        if row[0] + row[1] > 0:
            cs.append(rows_sub[i]); co.append(rows_obj[i]); cp.append(int(row.argmax())); cv.append(row.max())
        elif row[3] + row[4] > 0 or row[9] + row[10] > 0:
            for k in np.where(row > 0.5)[0]:
                cs.append(rows_sub[i]); co.append(rows_obj[i]); cp.append(int(k)); cv.append(row[k])
    res["semi"] = _match_sets(gt_rels, gt_boxes, gt_classes, np.asarray(cs, dtype=np.int64), np.asarray(co, dtype=np.int64),
                              np.asarray(cp, dtype=np.int64), np.asarray(cv, dtype=np.float64), *common) if cs else []
    return res


def matched_at_k(match_lists: List[List[int]], k: int):
    s = set()
    for m in match_lists[:k]:
        s.update(m)
    return tuple(sorted(s))


class Evaluator:
    """Mirror of SceneGraphEvaluator (lib/evaluation_recall.py:355-467): same result_dict keys and list contents."""

    def __init__(self, mode, AG_object_classes, AG_all_predicates, AG_attention_predicates, AG_spatial_predicates,
                 AG_contacting_predicates, iou_threshold=0.5, constraint=False, semithreshold=None):
        self.mode = mode
        self.all, self.att, self.spa, self.con = (list(AG_all_predicates), list(AG_attention_predicates),
                                                  list(AG_spatial_predicates), list(AG_contacting_predicates))
        self.num_rel = len(self.all)
        self.result_dict: Dict[str, dict] = {}

    def register_container(self):
        m = self.mode
        for t in ("_recall", "_recall_nogc", "_semi_recall"):
            self.result_dict[m + t] = {k: [] for k in KS}
        for t in ("_mean_recall", "_ng_mean_recall"):
            self.result_dict[m + t] = {k: 0.0 for k in KS}
            self.result_dict[m + t + "_collect"] = {k: [[] for _ in range(self.num_rel)] for k in KS}
            self.result_dict[m + t + "_list"] = {k: [] for k in KS}

    def _collect_mean(self, key, match_lists, gt_rels):
        for k in KS:
            match = matched_at_k(match_lists, k)
            hit = [0] * self.num_rel
            cnt = [0] * self.num_rel
            for g in range(gt_rels.shape[0]):
                cnt[int(gt_rels[g, 2])] += 1
                cnt[0] += 1
            for g in match:
                hit[int(gt_rels[g, 2])] += 1
                hit[0] += 1
            for n in range(self.num_rel):
                if cnt[n] > 0:
                    self.result_dict[self.mode + key + "_collect"][k][n].append(float(hit[n] / cnt[n]))

    def evaluate_scene_graph(self, gt, pred):
        """pred tensors may be torch (any device) or numpy; attention_distribution must already be softmaxed
        by the caller (the reference does it at :400 with torch, which this restatement does not re-implement)."""
        to_np = lambda x: x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
        im_idx = to_np(pred["im_idx"])
        pair_idx = to_np(pred["pair_idx"]).astype(np.int64)
        att, spa, con = to_np(pred["attention_distribution"]), to_np(pred["spatial_distribution"]), to_np(pred["contacting_distribution"])
        boxes = to_np(pred["boxes"])[:, 1:].astype(np.float32)
        if self.mode == "predcls":
            classes, oscores = to_np(pred["labels"]).astype(np.int64), to_np(pred["scores"]).astype(np.float32)
        else:
            classes, oscores = to_np(pred["pred_labels"]).astype(np.int64), to_np(pred["pred_scores"]).astype(np.float32)
        for idx, frame_gt in enumerate(gt):
            gb, gc, gr = build_frame_gt(frame_gt, self.att, self.spa, self.con, self.all)
            sel = im_idx == idx
            res = frame_matches(gb, gc, gr, pair_idx[sel], att[sel], spa[sel], con[sel], boxes, classes, oscores)
            G = gr.shape[0]
            for key, proto in (("_recall", "with"), ("_recall_nogc", "nogc"), ("_semi_recall", "semi")):
                for k in KS:
                    self.result_dict[self.mode + key][k].append(float(len(matched_at_k(res[proto], k))) / float(G))
            self._collect_mean("_mean_recall", res["with"], gr)
            self._collect_mean("_ng_mean_recall", res["nogc"], gr)

    def calculate_mean_recall(self):
        for t in ("_mean_recall", "_ng_mean_recall"):
            for k in KS:
                s = 0
                for n in range(self.num_rel):
                    lst = self.result_dict[self.mode + t + "_collect"][k][n]
                    r = 0.0 if len(lst) == 0 else np.mean(lst)
                    self.result_dict[self.mode + t + "_list"][k].append(r)
                    s += r
                self.result_dict[self.mode + t][k] = s / float(self.num_rel)
