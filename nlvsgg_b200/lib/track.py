"""Drop-in for lib/track.py:get_sequence (:127-262): writes ``entry["indices"]`` (list of LongTensor index groups).

predcls / sgdet (wks): group the boxes by (ground-truth / arg-max) class (:128-152; singletons collected in indices[0]
for sgdet).
sgcls: frame-by-frame Hungarian tracking.  The whole loop of the reference — matcher cost per frame (lib/matcher.py),
assignment, the tau = 0.5 accept rule, cluster / track bookkeeping, 50-frame expiry — runs as ONE kernel launch per call
(`nlv_track_sequence`, csrc/track.cu: one CTA per video, any number of videos per launch); the host only lays out the
frame table before and groups the per-detection cluster ids after."""
import ctypes
from typing import List, Sequence

import numpy as np
import torch

from .. import _C
from ..ops import _ptr, _stream
from .matcher import HungarianMatcher  # noqa: F401

MAX_GAP = 50          # lib/track.py:55


def _groups_by_value(values: torch.Tensor):
    v = values.detach().cpu().numpy()
    return [np.nonzero(v == u)[0] for u in np.unique(v)]


def frame_number(name: str) -> int:
    return int(name.split("/")[1].split(".")[0])          # lib/track.py:175


def track_videos(boxes: Sequence[torch.Tensor], feats: Sequence[torch.Tensor], dists: Sequence[torch.Tensor],
                 keys: Sequence[Sequence[int]], shape, weights=(1.0, 1.0, 1.0, 1.0)) -> List[torch.Tensor]:
    """Cluster id of every detection for a batch of videos (ids count up per video in creation order).
    boxes[v] f32[N_v,5] (frame index, x1, y1, x2, y2; frames ascending), feats[v] f32[N_v,F], dists[v] f32[N_v,C], keys[v] the
    frame number of each key frame.  One launch for all videos."""
    dev = boxes[0].device
    if dev.type != "cuda":
        raise RuntimeError("track_videos (nlvsgg_b200) runs on CUDA only; there is no CPU fallback")
    V = len(boxes)
    n_det = [int(b.shape[0]) for b in boxes]
    det_off = np.concatenate(([0], np.cumsum(n_det))).astype(np.int32)
    frame_off = np.concatenate(([0], np.cumsum([len(k) for k in keys]))).astype(np.int32)
    starts, cost_off, max_det = [], [], 0
    cost_elems = 0
    for v in range(V):
        fid = boxes[v][:, 0].detach().cpu().numpy().astype(np.int64)
        per = np.bincount(fid, minlength=len(keys[v]))[:len(keys[v])] if len(fid) else np.zeros(len(keys[v]), np.int64)
        assert per.sum() == n_det[v] and (len(fid) < 2 or np.all(np.diff(fid) >= 0)), "boxes must be grouped by ascending frame index"
        starts.append(np.concatenate(([0], np.cumsum(per))))
        md = int(per.max()) if len(per) else 0
        max_det = max(max_det, md)
        cost_off.append(cost_elems)
        cost_elems += md * n_det[v]
    N = int(det_off[-1])
    b = (boxes[0] if V == 1 else torch.cat(list(boxes))).contiguous().float()
    f = (feats[0] if V == 1 else torch.cat(list(feats))).contiguous().float()
    d = dists[0] if V == 1 else torch.cat(list(dists))
    cls = d.argmax(1).to(torch.int32).contiguous()
    i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    t_det, t_frame, t_start = i32(det_off), i32(frame_off), i32(np.concatenate(starts))
    t_key = i32(np.concatenate([np.asarray(k, dtype=np.int64) for k in keys]) if frame_off[-1] else np.zeros(0))
    t_cost = torch.from_numpy(np.asarray(cost_off, dtype=np.int64)).to(dev)
    lib = _C.lib()
    lib.nlv_track_sequence_workspace.restype = ctypes.c_longlong
    ws_bytes = int(lib.nlv_track_sequence_workspace(ctypes.c_longlong(N), int(f.shape[1]), int(d.shape[1]), ctypes.c_longlong(max(cost_elems, 1))))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cluster = torch.full((N,), -1, dtype=torch.int32, device=dev)
    ncl = torch.zeros(V, dtype=torch.int32, device=dev)
    status = torch.zeros(V, dtype=torch.int32, device=dev)
    w, h = shape
    F = ctypes.c_float
    _C.check(lib.nlv_track_sequence(_ptr(b), _ptr(f), int(f.shape[1]), _ptr(cls), int(d.shape[1]), _ptr(t_det), _ptr(t_frame), _ptr(t_start),
                                    _ptr(t_key), V, ctypes.c_longlong(N), max_det, F(w), F(h), F(weights[0]), F(weights[1]), F(weights[2]),
                                    F(weights[3]), MAX_GAP, _ptr(t_cost), ctypes.c_longlong(max(cost_elems, 1)), _ptr(ws), _ptr(cluster),
                                    _ptr(ncl), _ptr(status), _stream()), "track_sequence")
    if int(status.max().item()) != 0:
        raise RuntimeError("track_sequence: a frame has more detections / live tracks than the 1024-wide assignment state")
    return [cluster[det_off[v]:det_off[v + 1]] for v in range(V)]


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
    keys = [frame_number(annotation[0]["frame"]) for annotation in gt_annotation]
    weights = (matcher.cost_class, matcher.cost_feature, matcher.cost_bbox, matcher.cost_giou)
    cluster = track_videos([entry["boxes"]], [entry["features"]], [entry["distribution"]], [keys], shape, weights)[0]
    # clusters in creation order; members in detection order (a cluster gains at most one member per frame)
    order = torch.sort(cluster.long(), stable=True)[1]
    sizes = torch.bincount(cluster.long()).tolist() if cluster.numel() else []
    entry["indices"] = [g for g in torch.split(order, sizes) if g.numel() > 0]
