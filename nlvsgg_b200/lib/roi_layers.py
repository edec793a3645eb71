"""Drop-in for fasterRCNN/lib/model/roi_layers (roi_align.py:12-59, nms.py): ``ROIAlign(output_size, spatial_scale,
sampling_ratio)`` module with autograd, and ``nms(dets, scores, threshold)`` returning kept indices ascending."""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _C
from ..ops import _ptr, _stream


class _ROIAlign(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        if not input.is_cuda:
            raise RuntimeError("ROIAlign (nlvsgg_b200) runs on CUDA only; there is no CPU fallback")
        ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
        x = input.contiguous().float()
        r = roi.contiguous().float()
        b, c, h, w = x.shape
        out = torch.empty(r.shape[0], c, ph, pw, device=x.device, dtype=torch.float32)
        _C.check(_C.lib().nlv_roi_align_fwd(_ptr(x), b, c, h, w, _ptr(r), r.shape[0], ctypes.c_float(spatial_scale), ph, pw,
                                            int(sampling_ratio), _ptr(out), _stream()), "roi_align_fwd")
        ctx.save_for_backward(r)
        ctx.cfg = (ph, pw, float(spatial_scale), int(sampling_ratio), tuple(x.shape))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (r,) = ctx.saved_tensors
        ph, pw, scale, sr, (b, c, h, w) = ctx.cfg
        g = grad_output.contiguous().float()
        din = torch.zeros(b, c, h, w, device=g.device, dtype=torch.float32)
        _C.check(_C.lib().nlv_roi_align_bwd(_ptr(g), _ptr(r), r.shape[0], ctypes.c_float(scale), ph, pw, b, c, h, w, sr, _ptr(din),
                                            _stream()), "roi_align_bwd")
        return din, None, None, None, None


roi_align = _ROIAlign.apply


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size, self.spatial_scale, self.sampling_ratio = output_size, spatial_scale, sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio})")


def nms(dets: torch.Tensor, scores: torch.Tensor, threshold: float, strict: bool = True) -> torch.Tensor:
    """Kept original indices, ascending (nms.cu:127-130).  strict=True is the reference CUDA rule (IoU > thr)."""
    if not dets.is_cuda:
        raise RuntimeError("nms (nlvsgg_b200) runs on CUDA only; there is no CPU fallback")
    n = dets.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dets.device)
    d = dets[:, :4].contiguous().float()
    order = scores.float().sort(0, descending=True)[1].contiguous()
    words = (n + 63) // 64
    ws = torch.empty(n * words, dtype=torch.int64, device=d.device)
    flags = torch.zeros(n, dtype=torch.uint8, device=d.device)
    _C.check(_C.lib().nlv_nms(_ptr(d), _ptr(order), n, ctypes.c_float(threshold), 1 if strict else 0, _ptr(ws), _ptr(flags),
                              _stream()), "nms")
    return flags.nonzero().squeeze(1)
