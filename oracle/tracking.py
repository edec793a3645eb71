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


def lsap(cost):
    """TEST INFRASTRUCTURE.  Restatement of scipy.optimize.linear_sum_assignment (SciPy 1.18, the rectangular shortest
    augmenting path solver of Crouse 2016 that lib/matcher.py:147-149 calls): float64 duals, the problem transposed when
    it has more rows than columns, and the column choice rule (lowest reduced cost; an equal cost replaces the choice only
    when that column is unassigned) with the swap-remove `remaining` list.  csrc/track.cu:lsap_warp follows this line by
    line; tests/test_cpu_tracking.py pins it against scipy itself, ties included.  -> (row_ind, col_ind)."""
    import numpy as np
    c = np.asarray(cost, dtype=np.float64)
    tr = c.shape[1] < c.shape[0]
    if tr:
        c = c.T
    nr, nc = c.shape
    u, v = np.zeros(nr), np.zeros(nc)
    col4row, row4col = -np.ones(nr, dtype=np.int64), -np.ones(nc, dtype=np.int64)
    path = -np.ones(nc, dtype=np.int64)
    for cur in range(nr):
        remaining = [nc - it - 1 for it in range(nc)]
        SR, SC = np.zeros(nr, bool), np.zeros(nc, bool)
        spc = np.full(nc, np.inf)
        min_val, i, sink = 0.0, cur, -1
        while sink == -1:
            index, lowest = -1, np.inf
            SR[i] = True
            for it, j in enumerate(remaining):
                r = min_val + c[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            remaining[index] = remaining[-1]
            remaining.pop()
        u[cur] += min_val
        for r_ in range(nr):
            if SR[r_] and r_ != cur:
                u[r_] += min_val - spc[col4row[r_]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if tr:
        order = np.argsort(col4row, kind="stable")
        return col4row[order], order
    return np.arange(nr), col4row
