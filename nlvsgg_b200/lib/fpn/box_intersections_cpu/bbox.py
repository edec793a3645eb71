"""Drop-in for lib/fpn/box_intersections_cpu/bbox.pyx:15-19: ``bbox_overlaps(boxes, query_boxes) -> f64[N,K]``
(float64, +1 pixel convention) computed on the GPU; numpy in -> numpy out, tensors in -> tensor out."""
import numpy as np
import torch

from .... import ops


def bbox_overlaps(boxes, query_boxes):
    if torch.is_tensor(boxes):
        return ops.bbox_overlaps(boxes, query_boxes)
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float64)).cuda()
    q = torch.from_numpy(np.ascontiguousarray(query_boxes, dtype=np.float64)).cuda()
    return ops.bbox_overlaps(b, q).cpu().numpy()
