"""Drop-in for lib/track.py:get_sequence (:127-262): writes ``entry["indices"]`` (list of LongTensor index groups).

predcls / sgdet: group boxes by (ground-truth / arg-max) class (:128-152; singletons collected in indices[0] for sgdet).
sgcls: per-frame Hungarian tracking — detections are matched to live tracks with the fused cost kernel
(lib/matcher.py), accepted when either cosine distance is below 0.5 (:198), tracks die after a gap of 50 frame
numbers (:55).  The per-frame loop is inherently sequential (tracks evolve frame by frame) and stays on the host;
each frame costs one kernel launch + one host LSAP."""
import numpy as np
import torch
import torch.nn.functional as F

from .matcher import HungarianMatcher  # noqa: F401


class Tracker(object):
    def __init__(self, box, index, cluster):
        self.box, self.index, self.cluster, self.updated = box, index, cluster, False

    def update(self, box, index):
        if self.updated:
            return True
        self.updated = True
        if box is None:
            return index - self.index < 50
        self.box, self.index = box, index
        return True


def _groups_by_value(values: torch.Tensor):
    v = values.detach().cpu().numpy()
    return [np.nonzero(v == u)[0] for u in np.unique(v)]


def get_sequence(entry, gt_annotation, matcher, shape, task="sgcls"):
    dev = entry["boxes"].device
    if task == "predcls":
        entry["indices"] = [torch.from_numpy(g).to(dev) for g in _groups_by_value(entry["labels"])]
        return
    if task == "sgdet":
        groups = _groups_by_value(torch.argmax(entry["distribution"], 1))
        single = [g for g in groups if len(g) == 1]
        indices = [torch.from_numpy(np.concatenate(single)).to(dev) if single else torch.tensor([])]
        indices += [torch.from_numpy(g).to(dev) for g in groups if len(g) != 1]
        entry["indices"] = indices
        return
    assert task == "sgcls", "%s is not defined" % task
    w, h = shape
    key_frames = [annotation[0]["frame"] for annotation in gt_annotation]
    boxes, feats = entry["boxes"], entry["features"]
    dists = F.one_hot(entry["distribution"].argmax(1), entry["distribution"].shape[1]).float()
    frame_id = boxes[:, 0].detach().cpu().numpy()
    counts = np.cumsum([0] + np.unique(frame_id, return_counts=True)[1].tolist())
    boxes_cpu = boxes[:, 1:].detach().cpu()
    Z = torch.tensor([[w, h, w, h]], dtype=torch.float32)
    cluster, cluster_feature, cluster_dist, tracks = [], [], [], []

    def outside(p):
        return bool((p[0] + p[2] > h) or (p[1] + p[3] > w) or (p[0] < 0) or (p[1] < 0))

    for index, img in enumerate(key_frames):
        current_key = int(img.split("/")[1].split(".")[0])
        for t in tracks:
            t.updated = False
        rows = np.nonzero(frame_id == index)[0]
        pred = boxes_cpu[rows].clone()
        pred[:, 2:] = pred[:, 2:] - pred[:, :2]                     # xyxy -> xywh (matcher.py:15-19)
        norm_pred = pred / Z
        row_ind = []
        if len(tracks) > 0 and len(rows) > 0:
            norm_boxes = torch.stack([t.box for t in tracks]) / Z
            rows_t = torch.from_numpy(rows).to(dev)
            trk_feat = torch.cat([cluster_feature[t.cluster].mean(0, keepdim=True) for t in tracks])
            trk_dist = torch.cat([cluster_dist[t.cluster].mean(0, keepdim=True) for t in tracks])
            row_ind, col_ind, cost1, cost2 = matcher({"boxes": norm_pred, "features": feats[rows_t], "dists": dists[rows_t]},
                                                     {"boxes": norm_boxes, "features": trk_feat, "dists": trk_dist})
            for t_, (r, c) in enumerate(zip(row_ind, col_ind)):
                one = slice(int(rows[r]), int(rows[r]) + 1)
                if (cost1[t_] < 0.5) or (cost2[t_] < 0.5):
                    cluster[tracks[c].cluster].append(counts[index] + r)
                    if outside(pred[r]):
                        continue
                    cluster_feature[tracks[c].cluster] = torch.cat([cluster_feature[tracks[c].cluster], feats[one]])
                    cluster_dist[tracks[c].cluster] = torch.cat([cluster_dist[tracks[c].cluster], dists[one]])
                    tracks[c].update(pred[r], current_key)
                else:
                    cluster.append([counts[index] + r])
                    if outside(pred[r]):
                        cluster_feature.append([]); cluster_dist.append([])
                        continue
                    cluster_feature.append(feats[one]); cluster_dist.append(dists[one])
                    tracks.append(Tracker(pred[r], current_key, len(cluster) - 1))
        if len(row_ind) < len(pred):
            for j in range(len(pred)):
                if j not in row_ind:
                    cluster.append([counts[index] + j])
                    if outside(pred[j]):
                        cluster_feature.append([]); cluster_dist.append([])
                        continue
                    one = slice(int(rows[j]), int(rows[j]) + 1)
                    cluster_feature.append(feats[one]); cluster_dist.append(dists[one])
                    tracks.append(Tracker(pred[j], current_key, len(cluster) - 1))
        tracks = [t for t in tracks if t.updated or t.update(None, current_key)]
    entry["indices"] = [torch.LongTensor([int(i) for i in l]).to(dev) for l in cluster if len(l) > 0]
