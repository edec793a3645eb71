"""Drop-in for lib/draw_rectangles/draw_rectangles.pyx: ``draw_union_boxes(bbox_pairs f32[N,8], pooling_size) ->
f32[N,2,ps,ps]`` with the reference's host numpy signature (device tensors are accepted too and stay on device)."""
import numpy as np
import torch

from ... import ops


def draw_union_boxes(bbox_pairs, pooling_size, padding=0):
    assert padding == 0, "Padding>0 not supported yet"       # draw_rectangles.pyx:20
    if torch.is_tensor(bbox_pairs):
        return ops.draw_union_boxes(bbox_pairs, int(pooling_size))
    bp = torch.from_numpy(np.ascontiguousarray(bbox_pairs, dtype=np.float32)).cuda()
    return ops.draw_union_boxes(bp, int(pooling_size)).cpu().numpy()
