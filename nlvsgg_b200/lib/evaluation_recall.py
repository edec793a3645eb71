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
    """Ground truth of a list of frames flattened into the kernel's arrays (host side, numpy).

    `pack_video` walks the python annotation structure of ONE video once (evaluation_recall.py:404-419: person = box 0,
    class 1; attention / contacting triplets are (human, object, p), spatial ones (object, human, p)); ground truth is static
    across epochs, so evaluators cache the result per annotation object and a batch is a concatenation of cached arrays."""

    def __init__(self):
        self.rel, self.cls, self.box = [], [], []       # per frame (kept for callers that index frames)
        self.rel_off, self.box_off = [0], [0]

    @staticmethod
    def pack_video(gt, idx_att, idx_spa, idx_con):
        """-> (rel i32[G,3], cls i32[Gb], box f32[Gb,4], rels per frame i64[F], boxes per frame i64[F])"""
        rel, cls, box, nrel, nbox = [], [], [], [], []
        for frame_gt in gt:
            nb = len(frame_gt)
            cls.append(1)
            box.append(np.asarray(frame_gt[0]["person_bbox"], dtype=np.float64).reshape(-1)[:4])
            g0 = len(rel)
            for m, obj in enumerate(frame_gt[1:]):
                box.append(np.asarray(obj["bbox"], dtype=np.float64))
                cls.append(int(obj["class"]))
                a = np.asarray(obj["attention_relationship"]).reshape(-1)
                rel.append((0, m + 1, idx_att[int(a[0])]))
                for sp in np.asarray(obj["spatial_relationship"]).reshape(-1).tolist():
                    rel.append((m + 1, 0, idx_spa[int(sp)]))
                for c in np.asarray(obj["contacting_relationship"]).reshape(-1).tolist():
                    rel.append((0, m + 1, idx_con[int(c)]))
            nrel.append(len(rel) - g0)
            nbox.append(nb)
        return (np.asarray(rel, dtype=np.int32).reshape(-1, 3), np.asarray(cls, dtype=np.int32),
                np.asarray(box, dtype=np.float64).reshape(-1, 4).astype(np.float32),     # rounded to f32 exactly as :765 does
                np.asarray(nrel, dtype=np.int64), np.asarray(nbox, dtype=np.int64))

    def add_frame(self, frame_gt, idx_att, idx_spa, idx_con):
        rel, cls, box, nrel, nbox = self.pack_video([frame_gt], idx_att, idx_spa, idx_con)
        self.rel.append(rel); self.cls.append(cls); self.box.append(box)
        self.rel_off.append(self.rel_off[-1] + int(nrel[0]))
        self.box_off.append(self.box_off[-1] + int(nbox[0]))


def recall_match(pair_off, gt, pair_sub, pair_obj, att, spa, con, obj_scores, pred_cls, pred_boxes, stats=None):
    """Launch the kernel; returns u32[F,3,3,8] match sets (numpy).  gt: PackedGT (per-frame lists) or a dict of flat arrays
    {rel, cls, box, rel_off, box_off}.  All host arrays travel in ONE pinned upload."""
    from ..plan import pack_upload
    dev = att.device
    F = len(pair_off) - 1
    pmax, gmax, gbmax = _limits()
    po = np.asarray(pair_off, dtype=np.int32)
    if isinstance(gt, PackedGT):
        gt = {"rel_off": gt.rel_off, "box_off": gt.box_off,
              "rel": np.concatenate(gt.rel) if gt.rel else np.zeros((0, 3), np.int32),
              "cls": np.concatenate(gt.cls) if gt.cls else np.zeros(0, np.int32),
              "box": np.concatenate(gt.box) if gt.box else np.zeros((0, 4), np.float32)}
    ro, bo = np.asarray(gt["rel_off"], dtype=np.int32), np.asarray(gt["box_off"], dtype=np.int32)
    if F and (np.diff(po).max(initial=0) > pmax or np.diff(ro).max(initial=0) > gmax or np.diff(bo).max(initial=0) > gbmax):
        raise RuntimeError(f"recall_match: a frame exceeds the kernel limits (pairs<={pmax}, gt relations<={gmax}, gt boxes<={gbmax})")
    arrays = {"po": po, "ro": ro, "bo": bo, "rel": np.ascontiguousarray(gt["rel"], dtype=np.int32).reshape(-1, 3),
              "cls": np.ascontiguousarray(gt["cls"], dtype=np.int32), "box": np.ascontiguousarray(gt["box"], dtype=np.float32).reshape(-1, 4)}
    d = pack_upload(arrays, dev)
    out = torch.empty(F, 3, 3, 8, device=dev, dtype=torch.int32)
    ev0 = ev1 = None
    if stats is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _C.check(_C.lib().nlv_recall_match(F, _ptr(d["po"]), _ptr(d["ro"]), _ptr(d["bo"]), _ptr(pair_sub), _ptr(pair_obj), _ptr(att),
                                       _ptr(spa), _ptr(con), _ptr(obj_scores), _ptr(pred_cls), _ptr(pred_boxes), _ptr(d["rel"]),
                                       _ptr(d["cls"]), _ptr(d["box"]), _ptr(out), _stream()), "recall_match")
    if stats is not None:
        ev1.record()
    res = out.cpu().numpy().view(np.uint32)
    if stats is not None:
        stats["kernel_ms"] = ev0.elapsed_time(ev1)
        stats["h2d_bytes"] = int(sum(a.nbytes for a in arrays.values()))
        stats["d2h_bytes"] = int(res.nbytes)
        # algorithmic input of the kernel: scores 3P*26*4... = per pair 26 floats + 2 indices, per box 4+1+1 words, GT arrays
        P, N = int(pair_sub.shape[0]), int(pred_cls.shape[0])
        stats["algorithmic_bytes"] = P * (26 * 4 + 8) + N * 24 + stats["h2d_bytes"] + stats["d2h_bytes"]
    return res


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


class _FloatList(list):
    """A python list of floats (what the reference's mean-recall collectors are) that can be FED with numpy chunks: the chunks
    are turned into list items only when somebody looks at the list; the mean the evaluator itself needs is taken from the
    arrays (creating ~9 M python floats per evaluation of the test split cost more than the kernel and all other booking)."""

    def __init__(self, *a):
        super().__init__(*a)
        self._chunks = []

    def add_array(self, arr):
        self._chunks.append(np.asarray(arr, dtype=np.float64))

    def _flush(self):
        if self._chunks:
            ch, self._chunks = self._chunks, []
            super().extend(np.concatenate(ch).tolist())

    def mean_or_zero(self):
        """np.mean over the items in order (0.0 when empty), without materialising them."""
        parts = ([np.asarray([x for x in list.__iter__(self)], dtype=np.float64)] if list.__len__(self) else []) + self._chunks
        if not parts:
            return 0.0
        return np.mean(parts[0] if len(parts) == 1 else np.concatenate(parts))

    def __len__(self):
        return list.__len__(self) + sum(len(c) for c in self._chunks)

    def __iter__(self):
        self._flush(); return list.__iter__(self)

    def __getitem__(self, i):
        self._flush(); return list.__getitem__(self, i)

    def __eq__(self, other):
        self._flush()
        if isinstance(other, _FloatList):
            other._flush()
        return list.__eq__(self, other)

    __hash__ = None

    def __repr__(self):
        self._flush(); return list.__repr__(self)

    def append(self, x):
        self._flush(); list.append(self, x)

    def extend(self, xs):
        self._flush(); list.extend(self, xs)


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
            self.result_dict[m + t + "_collect"] = {k: [_FloatList() for _ in range(self.num_rel)] for k in KS}
            self.result_dict[m + t + "_list"] = {k: [] for k in KS}

    # ---- evaluation ----
    def evaluate_scene_graph(self, gt, pred):
        """One video (the reference signature).  Mutates pred['attention_distribution'] (softmax), as :400 does."""
        self.evaluate_videos([(gt, pred)])

    def _packed(self, gt):
        """Flattened ground truth of one video, cached per annotation object (static across epochs)."""
        cache = self.__dict__.setdefault("_gt_cache", {})
        hit = cache.get(id(gt))
        if hit is not None and hit[0] is gt:
            return hit[1]
        packed = PackedGT.pack_video(gt, self._ia, self._is, self._ic)
        if len(cache) > 65536:
            cache.clear()
        cache[id(gt)] = (gt, packed)
        return packed

    def evaluate_videos(self, items: Sequence[tuple]):
        """Any number of (gt, pred) videos in ONE kernel launch; results are appended in input order.  One device -> host
        read for the frame ids of all videos, one pinned upload of all ground truth, one read-back of the match sets."""
        if len(items) == 0:
            return
        preds = [p for _, p in items]
        dev = preds[0]["attention_distribution"].device
        if dev.type != "cuda":
            raise RuntimeError("SceneGraphEvaluator (nlvsgg_b200) needs CUDA tensors; there is no CPU fallback")
        cat = lambda ts: (ts[0] if len(ts) == 1 else torch.cat(ts, 0)).contiguous()
        n_pairs = [int(p["pair_idx"].shape[0]) for p in preds]
        n_boxes = [int(p["boxes"].shape[0]) for p in preds]
        att = nn.functional.softmax(cat([p["attention_distribution"].float() for p in preds]), dim=1)
        o = 0
        for p, n in zip(preds, n_pairs):          # :400 overwrites the caller's tensor with its softmax
            p["attention_distribution"] = att[o:o + n]
            o += n
        im_all = cat([p["im_idx"].reshape(-1) for p in preds]).to(torch.int64).cpu().numpy()
        packs = [self._packed(gt) for gt, _ in items]
        nf = np.asarray([len(pk[3]) for pk in packs], dtype=np.int64)
        # pairs per (video, frame): frame ids are sorted inside a video
        voff = np.concatenate(([0], np.cumsum(n_pairs)))
        foff = np.concatenate(([0], np.cumsum(nf)))
        vid = np.repeat(np.arange(len(items)), n_pairs)
        if len(im_all):
            assert np.all(im_all < nf[vid]) and np.all((np.diff(im_all) >= 0) | (np.diff(vid) > 0)), "im_idx must be sorted frame ids"
        cnt = np.bincount(foff[vid] + im_all, minlength=int(foff[-1])) if len(im_all) else np.zeros(int(foff[-1]), dtype=np.int64)
        pair_off = np.concatenate(([0], np.cumsum(cnt)))
        boff = np.concatenate(([0], np.cumsum(n_boxes)))[:-1]
        pi = cat([p["pair_idx"] for p in preds]).to(torch.int32) + torch.from_numpy(np.repeat(boff, n_pairs).astype(np.int32)).to(dev)[:, None]
        key_c, key_s = ("labels", "scores") if self.mode == "predcls" else ("pred_labels", "pred_scores")
        gtd = {"rel": np.concatenate([pk[0] for pk in packs]), "cls": np.concatenate([pk[1] for pk in packs]),
               "box": np.concatenate([pk[2] for pk in packs]),
               "rel_off": np.concatenate(([0], np.cumsum(np.concatenate([pk[3] for pk in packs])))),
               "box_off": np.concatenate(([0], np.cumsum(np.concatenate([pk[4] for pk in packs]))))}
        stats = {}
        masks = recall_match(pair_off, gtd, pi[:, 0].contiguous(), pi[:, 1].contiguous(), att, cat([p["spatial_distribution"].float() for p in preds]),
                             cat([p["contacting_distribution"].float() for p in preds]), cat([p[key_s].float() for p in preds]),
                             cat([p[key_c] for p in preds]).to(torch.int32), cat([p["boxes"] for p in preds])[:, 1:].float().contiguous(), stats)
        self.last_kernel_ms, self.last_h2d_bytes = stats["kernel_ms"], stats["h2d_bytes"]
        self.last_d2h_bytes, self.last_algorithmic_bytes = stats["d2h_bytes"], stats["algorithmic_bytes"]
        self._book(masks, gtd)

    def _book(self, masks: np.ndarray, gtd):
        """Integer match sets -> the reference's per-frame floats, in the reference's order (numpy over all frames at once)."""
        m = self.mode
        if isinstance(gtd, PackedGT):
            gtd = {"rel": np.concatenate(gtd.rel) if gtd.rel else np.zeros((0, 3), np.int32), "rel_off": np.asarray(gtd.rel_off)}
        F = masks.shape[0]
        rel_off = np.asarray(gtd["rel_off"], dtype=np.int64)
        G = np.diff(rel_off)                                              # GT relations per frame
        pred_of = gtd["rel"][:, 2].astype(np.int64)                       # predicate of every GT relation
        frame_of = np.repeat(np.arange(F), G)
        local = np.arange(len(pred_of)) - rel_off[frame_of]               # index of the relation inside its frame
        words = masks.reshape(F, 3, 3, 8)
        pop = np.bitwise_count(words).sum(-1).astype(np.float64)          # |matched set| per (frame, protocol, K): popcount of the 256-bit set
        Gf = G.astype(np.float64)
        for pi, key in enumerate(("_recall", "_recall_nogc", "_semi_recall")):
            for ki, k in enumerate(KS):
                self.result_dict[m + key][k].extend((pop[:, pi, ki] / Gf).tolist())
        # mean-recall collectors (:69-87 / :146-165): per frame and predicate n, hits / count for predicates present in the
        # frame; index 0 additionally counts every relation (the reference's `[0] += 1` alongside `[predicate] += 1`)
        # Sparse over the (frame, predicate) buckets that exist (frame-major, so a predicate's buckets are in frame order, the
        # order the reference appends in); predicate 0 is the special one: it is present in every frame with relations.
        flat = frame_of * self.num_rel + pred_of                          # (frame, predicate) bucket of every GT relation
        order = np.argsort(flat, kind="stable")
        fs = flat[order]
        first = np.ones(len(fs), dtype=bool)
        first[1:] = fs[1:] != fs[:-1]
        starts = np.nonzero(first)[0]                                     # one entry per existing bucket
        b_frame, b_pred = fs[starts] // self.num_rel, fs[starts] % self.num_rel
        b_cnt = np.diff(np.append(starts, len(fs))).astype(np.float64)
        sel_of = [np.nonzero(b_pred == n)[0] for n in range(1, self.num_rel)]       # buckets of predicate n >= 1, in frame order
        has_rel = G > 0
        cnt0 = G.astype(np.float64)                                       # predicate 0: every relation + the relations of predicate 0
        z = b_pred == 0
        cnt0[b_frame[z]] += b_cnt[z]
        wsel, bsel = local >> 5, (local & 31).astype(np.uint32)           # word / bit of a relation inside its frame's match set
        for pi, key in ((0, "_mean_recall"), (1, "_ng_mean_recall")):
            for ki, k in enumerate(KS):
                matched = ((words[frame_of, pi, ki, wsel] >> bsel) & np.uint32(1)).astype(np.float64)
                b_hit = np.add.reduceat(matched[order], starts) if len(starts) else np.zeros(0)
                coll = self.result_dict[m + key + "_collect"][k]
                hit0 = np.bincount(frame_of, weights=matched, minlength=F)
                hit0[b_frame[z]] += b_hit[z]
                coll[0].add_array((hit0 / np.maximum(cnt0, 1.0))[has_rel])
                ratio = b_hit / b_cnt
                for n in range(1, self.num_rel):
                    if len(sel_of[n - 1]):
                        coll[n].add_array(ratio[sel_of[n - 1]])

    def calculate_mean_recall(self):
        for t in ("_mean_recall", "_ng_mean_recall"):      # :89-109 / :167-187
            for k in KS:
                s = 0
                for n in range(self.num_rel):
                    lst = self.result_dict[self.mode + t + "_collect"][k][n]
                    r = lst.mean_or_zero() if isinstance(lst, _FloatList) else (0.0 if len(lst) == 0 else np.mean(lst))
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
