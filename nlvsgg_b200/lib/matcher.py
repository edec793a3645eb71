"""Drop-in for lib/matcher.py:HungarianMatcher (:81-150), entirely on the device: the four cost terms are one kernel
(``nlv_track_cost``) and the linear-sum assignment the reference hands to scipy on the host (:147-149) is ``nlv_lsap`` (one
warp, the same shortest-augmenting-path algorithm with float64 duals and scipy's tie rules, so assignments are identical)."""
import ctypes

import torch
from torch import nn

from .. import _C
from ..ops import _ptr, _stream


def track_cost(out_boxes, tgt_boxes, out_feat, tgt_feat, out_dist, tgt_dist, w_class, w_feat, w_bbox, w_giou):
    """-> (C, cost_dist, cost_feat) each f32[n_det, n_trk] on the device of the inputs."""
    dev = out_feat.device
    if dev.type != "cuda":
        raise RuntimeError("track_cost (nlvsgg_b200) runs on CUDA only; there is no CPU fallback")
    f = lambda t: t.to(dev).contiguous().float()
    ob, tb, of, tf, od, td = f(out_boxes), f(tgt_boxes), f(out_feat), f(tgt_feat), f(out_dist), f(tgt_dist)
    n, m = ob.shape[0], tb.shape[0]
    C = torch.empty(n, m, device=dev)
    cd, cf = torch.empty_like(C), torch.empty_like(C)
    F = ctypes.c_float
    _C.check(_C.lib().nlv_track_cost(_ptr(ob), _ptr(tb), _ptr(of), _ptr(tf), of.shape[1], _ptr(od), _ptr(td), od.shape[1], n, m,
                                     F(w_class), F(w_feat), F(w_bbox), F(w_giou), _ptr(C), _ptr(cd), _ptr(cf), _stream()),
             "track_cost")
    return C, cd, cf


def linear_sum_assignment(cost: torch.Tensor):
    """scipy.optimize.linear_sum_assignment for a CUDA float matrix -> (row_ind, col_ind) int64 tensors on the device,
    rows ascending (scipy's output order)."""
    if cost.device.type != "cuda":
        raise RuntimeError("linear_sum_assignment (nlvsgg_b200) runs on CUDA only; there is no CPU fallback")
    c = cost.contiguous().float()
    n, m = c.shape
    match = torch.full((n,), -1, dtype=torch.int32, device=c.device)
    if n and m:
        _C.check(_C.lib().nlv_lsap(_ptr(c), n, m, m, _ptr(match), _stream()), "lsap")
    rows = torch.nonzero(match >= 0)[:, 0]
    return rows, match[rows].long()


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_feature: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class, self.cost_feature, self.cost_bbox, self.cost_giou = cost_class, cost_feature, cost_bbox, cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def forward(self, outputs, targets):
        """outputs/targets: {"boxes": xywh (normalised), "features": [n,2048], "dists": [n,C]} (matcher.py:124-149).
        Returns (row_ind, col_ind) as numpy arrays like the reference, and the two cost vectors of the matched pairs."""
        C, cd, cf = track_cost(outputs["boxes"], targets["boxes"], outputs["features"], targets["features"], outputs["dists"],
                               targets["dists"], self.cost_class, self.cost_feature, self.cost_bbox, self.cost_giou)
        row_ind, col_ind = linear_sum_assignment(C)
        return row_ind.cpu().numpy(), col_ind.cpu().numpy(), cd[row_ind, col_ind].cpu(), cf[row_ind, col_ind].cpu()
