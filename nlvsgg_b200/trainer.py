"""Fused training step: forward, fused losses, hand-written backward, (NCCL gradient allreduce), global-norm clip and
AdamW over flat fp32 parameter / gradient / moment buffers — the kernel-level equivalent of one iteration of
tools/train_STTran.py:129-195 for a batch of videos (loss = mean over videos of the per-video loss)."""
from __future__ import annotations

from typing import Dict, List

import torch

from . import engine as E
from . import model as M
from . import ops

F32 = torch.float32


class GradSink(dict):
    """grads[name] = tensor  ->  copied into the parameter's slice of the flat gradient buffer."""

    def __init__(self, views: Dict[str, torch.Tensor]):
        super().__init__()
        self.views = views
        self.seen = set()

    def __setitem__(self, name, value):
        v = self.views[name]
        assert value.numel() == v.numel(), name
        ops.convert(value.contiguous().reshape(1, -1), F32, out=v.reshape(1, -1))
        self.seen.add(name)


class Trainer:
    def __init__(self, state: Dict[str, torch.Tensor], mode: str = "sgdet", arch: str = "sttran", precision: str = "bf16",
                 lr: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 5.0,
                 device="cuda"):
        dev = torch.device(device)
        self.mode, self.arch, self.dev = mode, arch, dev
        self.k = E.Kernels(precision)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        is_param = lambda n: not (n.endswith("running_mean") or n.endswith("running_var") or n.endswith("num_batches_tracked")
                                  or n.endswith(".pe"))
        self.param_names = [n for n in state if is_param(n) and "encoder_tran" not in n]
        total = sum(state[n].numel() for n in self.param_names)
        self.flat_p = torch.empty(total, device=dev, dtype=F32)
        self.flat_g = torch.zeros(total, device=dev, dtype=F32)
        self.flat_m = torch.zeros(total, device=dev, dtype=F32)
        self.flat_v = torch.zeros(total, device=dev, dtype=F32)
        self.P: Dict[str, torch.Tensor] = {}
        self.gviews: Dict[str, torch.Tensor] = {}
        off = 0
        for n in self.param_names:
            t = state[n]
            self.P[n] = self.flat_p[off:off + t.numel()].view(t.shape)
            self.P[n].copy_(t)
            self.gviews[n] = self.flat_g[off:off + t.numel()].view(t.shape)
            off += t.numel()
        for n, t in state.items():
            if n not in self.P and "encoder_tran" not in n:
                self.P[n] = t.to(dev).clone()
        self.total_sq = torch.zeros(1, device=dev, dtype=F32)
        self.step_count = 0
        self.n_params = total
        self.last_sink = None

    def forward_backward(self, batch: M.Batch):
        """batch: device-resident collated batch.  Plan + labels are (re)built from its host metadata every call —
        they replace the reference's per-frame python loops and are part of the step."""
        dsg = self.arch == "dsg"
        plan = M.make_plan(batch, self.dev, self.mode, dsg)
        labels = M.make_labels(batch, self.dev, self.mode)
        fwd = M.dsg_forward if dsg else M.sttran_forward
        bwd = M.dsg_backward if dsg else M.sttran_backward
        out, ctx = fwd(self.k, self.P, batch, plan, self.mode, True, True)
        loss, d26, dobj = M.fused_loss(out, batch, labels, self.mode)
        self.last_sink = GradSink(self.gviews)
        bwd(self.k, self.P, batch, plan, self.mode, ctx, d26, dobj, grads=self.last_sink)
        return loss, out

    def optimizer_step(self):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            ws = torch.distributed.get_world_size()
            chunk = 32 * 1024 * 1024  # 128 MB buckets over NVLink
            works = [torch.distributed.all_reduce(self.flat_g[i:i + chunk], async_op=True)
                     for i in range(0, self.n_params, chunk)]
            for w in works:
                w.wait()
            self.flat_g.mul_(1.0 / ws)
        self.step_count += 1
        self.total_sq.zero_()
        ops.sumsq(self.flat_g, self.total_sq)
        ops.adamw_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps,
                       self.wd, self.step_count, self.total_sq, self.max_norm)
        self.k._wcache.clear()   # operand copies of the weights are stale now

    def step(self, batch: M.Batch):
        loss, _ = self.forward_backward(batch)
        self.optimizer_step()
        return loss

    def step_from_host(self, host_batch: M.Batch):
        """End-to-end step: pinned host buffers -> device inside the step."""
        return self.step(M.upload(host_batch, self.dev))
