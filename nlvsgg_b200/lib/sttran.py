"""Drop-in for lib/sttran.py (reference lines cited inline).  Same constructor, same ``forward(entry)`` dict
contract, same state_dict names; the torch submodules below only *hold* the parameters — every FLOP of
forward and backward runs in the nlv_b200 kernels (nlvsgg_b200/engine.py)."""
import os

import torch
import torch.nn as nn

from .. import autograd as A
from .. import engine as E
from .. import ops
from .roi_layers import ROIAlign, nms
from .transformer_wk import transformer_wk
from .word_vectors import obj_edge_vectors


class ObjectClassifier(nn.Module):
    """Parameter container of lib/sttran.py:20-51 (predcls pass-through :90-92, sgdet/wks head :173-184) and the
    detection-filtering / pairing logic of the non-wks sgdet TEST branch (:185-283)."""

    def __init__(self, mode="sgdet", obj_classes=None, is_wks=True):
        super().__init__()
        self.classes, self.mode, self.is_wks = obj_classes, mode, is_wks
        embed_vecs = obj_edge_vectors(obj_classes[1:], wv_type="glove.6B", wv_dir="data", wv_dim=200)
        self.obj_embed = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=0.01 / 10.0), nn.Linear(4, 128), nn.ReLU(inplace=True),
                                       nn.Dropout(0.1))
        self.obj_dim = 2048
        self.decoder_lin = nn.Sequential(nn.Linear(self.obj_dim + 200 + 128, 1024), nn.BatchNorm1d(1024), nn.ReLU(),
                                         nn.Linear(1024, len(self.classes)))
        self.RCNN_roi_align = ROIAlign((7, 7), 1.0 / 16.0, 0)          # lib/sttran.py:37
        self.nms_strict = True     # reference CUDA rule (IoU > 0.6 suppresses, nms.cu); False = its CPU rule (>=, nms_cpu.cpp)

    # ---- non-weakly-supervised sgdet, eval mode: lib/sttran.py:185-283 ------------------------------------------
    @staticmethod
    def _clean_class(boxes, dist, feats, labels, class_idx):
        """lib/sttran.py:53-81 without the per-frame loop: every frame keeps its detections and gets, right after them, a second
        copy of those labelled `class_idx` with that class's score zeroed and the label re-decided."""
        sel = labels == class_idx
        nd = dist[sel].clone()
        nd[:, class_idx - 1] = 0
        nl = nd.argmax(1) + 1 if nd.shape[0] else labels[:0]
        frame = torch.cat((boxes[:, 0], boxes[sel, 0]))
        key = frame * 2 + torch.cat((torch.zeros_like(boxes[:, 0]), torch.ones_like(boxes[sel, 0])))
        order = torch.sort(key, stable=True)[1]
        return (torch.cat((boxes, boxes[sel]))[order], torch.cat((dist, nd))[order], torch.cat((feats, feats[sel]))[order],
                torch.cat((labels, nl))[order])

    @torch.no_grad()
    def sgdet_test_branch(self, entry):
        """Detector output -> (NMS-filtered boxes, labels, human per frame, pairs, union features, spatial masks); mutates and
        returns `entry` with the keys the reference writes.  Index logic is torch on the device; NMS, RoIAlign and the mask
        rasteriser are the nlv_b200 kernels.  One host sync (group sizes) replaces the reference's per-frame / per-class syncs."""
        boxes, dist = entry["boxes"].float(), entry["distribution"].float()
        feats, labels = entry["features"], entry["pred_labels"].long()
        dev = boxes.device
        b = int(boxes[-1, 0].item()) + 1                                                        # :190
        for c in (5, 8, 17):                                                                    # :196-198
            boxes, dist, feats, labels = self._clean_class(boxes, dist, feats, labels, c)
        # per (frame, arg-max class) groups, score-descending inside a group (:204-232)
        am = dist.argmax(1)
        cls_score = dist.gather(1, am[:, None])[:, 0]
        o1 = torch.sort(cls_score, descending=True, stable=True)[1]
        group = (boxes[:, 0].long() * (len(self.classes) - 1) + am)[o1]
        o2 = torch.sort(group, stable=True)[1]
        order = o1[o2]
        group = group[o2]
        boxes, dist, feats, cls_score = boxes[order], dist[order], feats[order], cls_score[order]
        sizes = torch.unique_consecutive(group, return_counts=True)[1].tolist()              # the one host sync
        keep, start = [], 0
        for n in sizes:
            k = nms(boxes[start:start + n, 1:], cls_score[start:start + n], 0.6, strict=self.nms_strict)
            keep.append(k + start)
            start += n
        keep = torch.cat(keep)
        boxes, dist, feats = boxes[keep], dist[keep], feats[keep]
        box_idx = boxes[:, 0].long()
        n_box = boxes.shape[0]
        pred_scores, pred_labels = torch.max(dist[:, 1:], dim=1)                                # :240-241
        pred_labels = pred_labels + 2
        # human of a frame = its detection with the highest person score, first one on ties (:244-252); frames without
        # detections keep index 0, exactly as the reference's zero-initialised HUMAN_IDX does
        hs = torch.sort(dist[:, 0], descending=True, stable=True)[1]
        hs = hs[torch.sort(box_idx[hs], stable=True)[1]]
        first = torch.ones(n_box, dtype=torch.bool, device=dev)
        first[1:] = box_idx[hs][1:] != box_idx[hs][:-1]
        human = torch.zeros(b, dtype=torch.int64, device=dev)
        human[box_idx[hs][first]] = hs[first]
        pred_labels[human] = 1                                                                   # :254-255
        pred_scores[human] = dist[human, 0]
        objs = torch.nonzero(pred_labels != 1)[:, 0]                                            # :257-263 (boxes are frame-sorted)
        pair = torch.stack((human[box_idx[objs]], objs), 1)
        im_idx = box_idx[objs].float()
        union_boxes = torch.cat((im_idx[:, None], torch.min(boxes[pair[:, 0], 1:3], boxes[pair[:, 1], 1:3]),
                                 torch.max(boxes[pair[:, 0], 3:5], boxes[pair[:, 1], 3:5])), 1)  # :270-273
        entry["boxes"], entry["distribution"], entry["features"] = boxes, dist, feats
        entry["pred_scores"], entry["pred_labels"] = pred_scores, pred_labels
        entry["pair_idx"], entry["im_idx"], entry["human_idx"] = pair, im_idx, human[:, None]
        entry["union_feat"] = self.RCNN_roi_align(entry["fmaps"], union_boxes)                   # :275
        entry["union_box"] = union_boxes
        entry["spatial_masks"] = ops.union_mask_pairs(boxes, pair, 27, -0.5)                     # :279-281
        return entry


    # ---- sgcls, eval mode: lib/sttran.py:105-170 --------------------------------------------------------------------
    def union_features(self, entry, frame_id, boxes_xyxy):
        """Union-box features of one frame (lib/sttran.py:154-160).  The reference calls VinVL through
        lib/extract_bbox_features.py (un-vendored scene_graph_benchmark); assign `union_feature_extractor` (callable
        (entry, frame_id, boxes_xyxy[n,4]) -> [n,2048,7,7]) to plug any extractor in."""
        fn = getattr(self, "union_feature_extractor", None)
        if fn is not None:
            return fn(entry, frame_id, boxes_xyxy)
        try:
            from lib.extract_bbox_features import extract_feature_given_bbox_base_feat_torch as vinvl
        except Exception as e:
            raise RuntimeError("sgcls evaluation re-extracts union features with the VinVL detector (lib/sttran.py:159, "
                               "lib/extract_bbox_features.py); it is not importable here — set "
                               "model.object_classifier.union_feature_extractor") from e
        return vinvl(entry["faset_rcnn_model"], entry["transforms"], entry["cv2_imgs"][frame_id], boxes_xyxy, entry["fmaps"][frame_id], False)

    @torch.no_grad()
    def sgcls_test_branch(self, entry, logits):
        """Classifier logits [N,37] -> labels, the human of every frame, the duplicate clean-up of the frame's most frequent
        label, (human, object) pairs, union boxes / features and spatial masks; mutates and returns `entry` with the keys
        the reference writes.  The per-frame / per-box python loops of :115-143 are index arithmetic on the device."""
        boxes = entry["boxes"].float()
        dev = boxes.device
        box_idx = boxes[:, 0].long()
        n = boxes.shape[0]
        b = int(box_idx[-1].item()) + 1
        dist = torch.softmax(logits[:, 1:].float(), dim=1)                                      # :107
        scores, labels = torch.max(dist[:, 1:], dim=1)                                          # :108-109
        labels = labels + 2
        # human of a frame = its box with the highest person probability, first one on ties (:115-117)
        hs = torch.sort(dist[:, 0], descending=True, stable=True)[1]
        hs = hs[torch.sort(box_idx[hs], stable=True)[1]]
        first = torch.ones(n, dtype=torch.bool, device=dev)
        first[1:] = box_idx[hs][1:] != box_idx[hs][:-1]
        human = torch.zeros(b, dtype=torch.int64, device=dev)
        human[box_idx[hs][first]] = hs[first]
        labels[human] = 1                                                                        # :119-120
        scores[human] = dist[human, 0]
        # most frequent label of every frame (smallest one on ties, as torch.mode); all its boxes but the most confident
        # one lose that class and are re-labelled (:123-135)
        ncls = dist.shape[1] + 1
        hist = torch.zeros(b, ncls + 1, dtype=torch.int64, device=dev)
        hist.index_put_((box_idx, labels), torch.ones_like(labels), accumulate=True)
        dup = hist.argmax(1)[box_idx]                                                            # per box: its frame's duplicate class
        cand = labels == dup
        conf = dist.gather(1, (dup - 1).clamp(min=0)[:, None])[:, 0]
        # the highest-confidence candidate of each frame survives (ascending argsort, last one kept)
        key = torch.where(cand, conf, torch.full_like(conf, -1.0))
        o1 = torch.sort(key, stable=True)[1]
        o2 = o1[torch.sort(box_idx[o1], stable=True)[1]]
        last = torch.ones(n, dtype=torch.bool, device=dev)
        last[:-1] = box_idx[o2][:-1] != box_idx[o2][1:]
        keep = torch.zeros(n, dtype=torch.bool, device=dev)
        keep[o2[last]] = True
        change = cand & ~keep
        rows = torch.nonzero(change)[:, 0]
        dist[rows, dup[rows] - 1] = 0
        new_score, new_label = torch.max(dist[rows], dim=1)
        labels[rows], scores[rows] = new_label + 1, new_score
        # pairs (:138-146), union boxes (:150-151)
        objs = torch.nonzero(labels != 1)[:, 0]
        pair = torch.stack((human[box_idx[objs]], objs), 1)
        im_idx = box_idx[objs].float()
        union_boxes = torch.cat((im_idx[:, None], torch.min(boxes[pair[:, 0], 1:3], boxes[pair[:, 1], 1:3]),
                                 torch.max(boxes[pair[:, 0], 3:5], boxes[pair[:, 1], 3:5])), 1)
        frames_with_pairs = torch.unique(box_idx[objs]).tolist()
        feats = [self.union_features(entry, f, union_boxes[union_boxes[:, 0] == f][:, 1:]) for f in frames_with_pairs]   # :154-160
        entry["distribution"], entry["pred_scores"], entry["pred_labels"] = dist, scores, labels
        entry["pair_idx"], entry["im_idx"], entry["human_idx"] = pair, im_idx, human[:, None]
        entry["union_feat"] = torch.cat(feats).to(dev)
        entry["union_box"] = union_boxes
        entry["spatial_masks"] = ops.union_mask_pairs(boxes, pair, 27, -0.5)                     # :163-165
        return entry


class STTran(nn.Module):
    def __init__(self, mode="sgdet", attention_class_num=None, spatial_class_num=None, contact_class_num=None,
                 obj_classes=None, enc_layer_num=None, dec_layer_num=None, transformer_mode=None, is_wks=True,
                 feat_dim=2048, motifs_path=None, conf=None, precision=None):
        super().__init__()
        assert mode in ("sgdet", "sgcls", "predcls")          # lib/sttran.py:328
        self.conf, self.obj_classes, self.mode, self.is_wks = conf, obj_classes, mode, is_wks
        self.attention_class_num, self.spatial_class_num, self.contact_class_num = \
            attention_class_num, spatial_class_num, contact_class_num
        self.transformer_mode, self.motifs_path = transformer_mode, motifs_path
        assert (attention_class_num, spatial_class_num, contact_class_num) == (3, 6, 17) and feat_dim == 2048
        self.object_classifier = ObjectClassifier(mode=mode, obj_classes=obj_classes, is_wks=is_wks)
        self.union_func1 = nn.Conv2d(feat_dim, 256, 1, 1)
        self.conv = nn.Sequential(
            nn.Conv2d(2, 256 // 2, kernel_size=7, stride=2, padding=3, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256 // 2, momentum=0.01), nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(256 // 2, 256, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256, momentum=0.01))
        self.subj_fc, self.obj_fc, self.vr_fc = nn.Linear(2048, 512), nn.Linear(2048, 512), nn.Linear(256 * 7 * 7, 512)
        embed_vecs = obj_edge_vectors(obj_classes, wv_type="glove.6B", wv_dir="data", wv_dim=200)
        self.obj_embed = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.obj_embed2 = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed2.weight.data = embed_vecs.clone()
        self.glocal_transformer = transformer_wk(enc_layer_num=enc_layer_num, dec_layer_num=dec_layer_num, embed_dim=1936,
                                                 nhead=8, dim_feedforward=2048, dropout=0.1, mode="latter")
        self.a_rel_compress = nn.Linear(1936, attention_class_num)
        self.s_rel_compress = nn.Linear(1936, spatial_class_num)
        self.c_rel_compress = nn.Linear(1936, contact_class_num)
        self.kernels = E.Kernels(precision or os.environ.get("NLV_PRECISION", "bf16"), dropout=0.1)
        # NLV_ADDITIVE_INT_MASK=1 (or model.kernels.additive_mask = True): evaluate with the torch-1.10.1 reading of the int
        # key_padding_mask of lib/transformer_wk.py:154 (checkpoints trained by the original environment); inference only
        self.kernels.additive_mask = os.environ.get("NLV_ADDITIVE_INT_MASK", "0") == "1"

    def forward(self, entry):
        """lib/sttran.py:375-411: mutates and returns ``entry``."""
        if self.mode == "sgcls" and not self.training:
            # lib/sttran.py:105-170: the classifier head runs first (one sequencer call that stops after it), the pairs are
            # built from its labels, then the relation path runs like predcls on the inferred labels
            logits = A.run_object_head(self, self.kernels, entry)
            entry = self.object_classifier.sgcls_test_branch(entry, logits)
            view = dict(entry)
            view["labels"], view["scores"] = entry["pred_labels"], entry["pred_scores"]
            _, att, spa, con, _ = A.run_module(self, self.kernels, [view], "predcls", "sttran")
            entry["attention_distribution"], entry["spatial_distribution"], entry["contacting_distribution"] = att, spa, con
            return entry
        if self.mode == "sgdet" and not self.is_wks and not self.training:
            # :185-283 — detections are filtered and paired here; the object classifier head is NOT applied (the detector's
            # distribution is kept), so the relation path runs exactly like predcls on the inferred labels
            entry = self.object_classifier.sgdet_test_branch(entry)
            view = dict(entry)
            view["labels"], view["scores"] = entry["pred_labels"], entry["pred_scores"]
            _, att, spa, con, _ = A.run_module(self, self.kernels, [view], "predcls", "sttran")
            entry["attention_distribution"], entry["spatial_distribution"], entry["contacting_distribution"] = att, spa, con
            return entry
        obj, att, spa, con, batch = A.run_module(self, self.kernels, [entry], self.mode, "sttran")
        entry["pred_labels"] = entry["labels"]                      # :91 / :183
        if self.mode != "predcls":
            entry["distribution"] = obj                            # :182  logits [N,37]
            entry["pred_scores"] = entry["scores"]                  # :184
        if "spatial_masks" not in entry:
            entry["spatial_masks"] = batch.spatial_masks
        entry["attention_distribution"] = att                       # :404 (logits)
        entry["spatial_distribution"] = spa                         # :408
        entry["contacting_distribution"] = con                      # :409
        return entry
