"""TEST INFRASTRUCTURE — restatement of the DSG-DETR tracking cost (lib/matcher.py:49-78,102-150) in plain torch fp32.
Only tests/ may import this module."""
import torch


def cosine_distance(x, y):
    """matcher.py:70-78 — rows are divided by (norm + 1e-12) first, then multiplied."""
    x = x / (x.norm(dim=1, keepdim=True) + 1e-12)
    y = y / (y.norm(dim=1, keepdim=True) + 1e-12)
    return 1 - x @ y.t()


def xywh_to_cxcywh(b):
    return torch.stack((b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2, b[:, 2], b[:, 3]), 1)


def cxcywh_to_xyxy(b):
    return torch.stack((b[:, 0] - 0.5 * b[:, 2], b[:, 1] - 0.5 * b[:, 3], b[:, 0] + 0.5 * b[:, 2], b[:, 1] + 0.5 * b[:, 3]), 1)


def giou(a, b):
    """matcher.py:34-68 (torchvision box_area = (x2-x1)*(y2-y1), no +1)."""
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[:, :2]); rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2 - inter
    iou = inter / union
    lt = torch.min(a[:, None, :2], b[:, :2]); rb = torch.max(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


def matcher_cost(out, tgt, w_class=0.5, w_feat=1.0, w_bbox=1.0, w_giou=0.5):
    """-> (C, cost_dist, cost_feat); out/tgt = {"boxes" (xywh), "features", "dists"} as HungarianMatcher.forward takes."""
    ob, tb = xywh_to_cxcywh(out["boxes"]), xywh_to_cxcywh(tgt["boxes"])
    cd = cosine_distance(out["dists"], tgt["dists"])
    cf = cosine_distance(out["features"], tgt["features"])
    cb = torch.cdist(ob, tb, p=1)
    cg = -giou(cxcywh_to_xyxy(ob), cxcywh_to_xyxy(tb))
    return w_class * cd + w_feat * cf + w_bbox * cb + w_giou * cg, cd, cf
