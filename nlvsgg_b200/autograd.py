"""torch.autograd glue: one Function per whole model, so the reference training scripts
(loss.backward(); clip_grad_norm_; optimizer.step()) drive the hand-written backward unchanged.

Forward and backward are ONE C call each (csrc/step.cu).  The operand (bf16) copies of the weights are re-made from the
live fp32 parameters inside every forward call — the reference optimiser updates through `p.data` (lib/AdamW.py:69,112),
which no version counter sees, so nothing derived from a parameter is ever cached across calls."""
from __future__ import annotations

from typing import Dict, List

import torch

from . import engine as E
from . import model as M
from . import ops


def module_tensors(module: torch.nn.Module) -> Dict[str, torch.Tensor]:
    P = dict(module.named_parameters())
    P.update(dict(module.named_buffers()))
    return P


class _WholeModelFn(torch.autograd.Function):
    """outputs: (object logits [N,37] or empty, attention logits [R,3], spatial probs [R,6], contacting probs [R,17])"""

    @staticmethod
    def forward(ctx, runner, names: List[str], *params):
        want_ctx = runner.want_ctx   # decided by the caller: grad mode is always off inside Function.forward
        out, saved = runner.fwd(want_ctx)
        att, spa, con = out["att"], out["spa"], out["con"]
        obj = out.get("distribution")
        if obj is None:
            obj = torch.empty(0, device=att.device)
        ctx.runner, ctx.saved, ctx.names = runner, saved, names
        ctx.spa, ctx.con = spa, con
        ctx.has_obj = "distribution" in out
        return obj, att, spa, con

    @staticmethod
    def backward(ctx, dobj, datt, dspa, dcon):
        spa, con = ctx.spa, ctx.con
        d26 = ops.heads_activation_bwd(datt, dspa, dcon, spa, con, spa.shape[0])
        dobj_l = dobj.contiguous() if (ctx.has_obj and dobj is not None) else None
        if ctx.has_obj and dobj_l is None:
            dobj_l = torch.zeros(ctx.runner.plan.N, 37, device=spa.device)
        grads = ctx.runner.bwd(ctx.saved, d26, dobj_l)
        ctx.saved = None
        if ctx.runner.plan.Mg == 0:
            # a video without a sliding window never reaches the temporal decoder / position embedding: the reference leaves
            # their .grad None (lib/AdamW.py:66 then skips them)
            grads = {n: g for n, g in grads.items() if "global_attention" not in n and "position_embedding" not in n}
        return (None, None) + tuple(grads.get(n) for n in ctx.names)


class Runner:
    """Binds a parameter dict, a batch and a plan to the sequencer's forward / backward."""

    def __init__(self, kernels: E.Kernels, P, batch, plan, mode: str, training: bool, arch: str = "sttran"):
        self.k, self.P, self.batch, self.plan, self.mode, self.training, self.arch = kernels, P, batch, plan, mode, training, arch
        self.want_ctx = False

    def fwd(self, want_ctx: bool):
        f = M.sttran_forward if self.arch == "sttran" else M.dsg_forward
        return f(self.k, self.P, self.batch, self.plan, self.mode, self.training, want_ctx, activations=True, with_backward=want_ctx,
                 fresh_ws=True)

    def bwd(self, saved, d26, dobj):
        f = M.sttran_backward if self.arch == "sttran" else M.dsg_backward
        return f(self.k, self.P, self.batch, self.plan, self.mode, saved, d26, dobj)


def run_module(module: torch.nn.Module, kernels: E.Kernels, entries, mode: str, arch: str):
    """Shared forward of the drop-in modules: returns (obj_logits|None, att, spa, con, batch)."""
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("nlvsgg_b200 models run on CUDA only (sm_100a kernels); there is no CPU fallback")
    batch, plan = M.make_batch(entries, dev, mode, dsg=(arch == "dsg"))
    P = {k: v.detach() for k, v in module_tensors(module).items()}
    runner = Runner(kernels, P, batch, plan, mode, module.training, arch)
    names = [n for n, p in module.named_parameters()]
    params = [p for n, p in module.named_parameters()]
    runner.want_ctx = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    kernels.seed += 1
    obj, att, spa, con = _WholeModelFn.apply(runner, names, *params)
    if module.training:   # BatchNorm bookkeeping the kernels do not touch
        for n, buf in module.named_buffers():
            if n.endswith("num_batches_tracked") and "encoder_tran" not in n:
                if not (mode == "predcls" and n.startswith("object_classifier.")):
                    buf += len(entries)
    return (obj if obj.numel() or mode != "predcls" else None), att, spa, con, batch


def run_object_head(module: torch.nn.Module, kernels: E.Kernels, entry) -> torch.Tensor:
    """Object-classifier logits [N,37] of lib/sttran.py:98-102 alone (inference): the sequencer stops after the head
    (NLV_RUN_OBJECT_ONLY); the entry needs no pairs yet."""
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("nlvsgg_b200 models run on CUDA only (sm_100a kernels); there is no CPU fallback")
    e = {k: entry[k] for k in ("boxes", "features", "distribution", "labels", "scores") if k in entry}
    n = e["boxes"].shape[0]
    if "labels" not in e:
        e["labels"] = torch.zeros(n, dtype=torch.int64, device=dev)
    if "scores" not in e:
        e["scores"] = torch.zeros(n, device=dev)
    e["pair_idx"] = torch.zeros(0, 2, dtype=torch.int64, device=dev)
    e["im_idx"] = torch.zeros(0, device=dev)
    e["union_feat"] = torch.zeros(0, 2048, 7, 7, device=dev)
    batch, plan = M.make_batch([e], dev, "sgcls")
    P = {k: v.detach() for k, v in module_tensors(module).items()}
    desc = M._desc(kernels, P, "sttran", "sgcls")
    out, _ = E.run_forward(kernels, desc, P, batch, plan, False, False, fresh_ws=True, object_only=True)
    return out["distribution"]
