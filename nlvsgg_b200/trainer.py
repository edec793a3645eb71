"""Fused training step: forward, fused losses, hand-written backward, (NCCL gradient allreduce), global-norm clip and
AdamW over flat fp32 parameter / gradient / moment buffers — the kernel-level equivalent of one iteration of
tools/train_STTran.py:129-195 for a batch of videos (loss = mean over videos of the per-video loss)."""
from __future__ import annotations

from typing import Dict, List

import torch

from . import dist as D
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
        names = [n for n in state if is_param(n) and "encoder_tran" not in n]
        # groups that the kernels consume as ONE matrix are laid out adjacently so a single view serves them
        groups = [["subj_fc.weight", "obj_fc.weight"], ["subj_fc.bias", "obj_fc.bias"],
                  ["a_rel_compress.weight", "s_rel_compress.weight", "c_rel_compress.weight"],
                  ["a_rel_compress.bias", "s_rel_compress.bias", "c_rel_compress.bias"]]
        grouped = {n for g in groups for n in g}
        self.param_names = [n for g in groups for n in g] + [n for n in names if n not in grouped]
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in self.param_names:
            if state[n].dim() >= 2:
                off += (-off) % 8             # 16-byte aligned rows for the bf16 mirror (TMA)
            self.offsets[n] = off
            off += state[n].numel()
        total, pad = off, (-off) % 8
        self.flat_p = torch.zeros(total + pad, device=dev, dtype=F32)
        self.flat_g = torch.zeros(total + pad, device=dev, dtype=F32)
        self.flat_m = torch.zeros(total + pad, device=dev, dtype=F32)
        self.flat_v = torch.zeros(total + pad, device=dev, dtype=F32)
        self.flat_pb = torch.zeros(total + pad, device=dev, dtype=torch.bfloat16) if precision == "bf16" else None
        self.P: Dict[str, torch.Tensor] = {}
        self.gviews: Dict[str, torch.Tensor] = {}
        for n in self.param_names:
            t, off = state[n], self.offsets[n]
            self.P[n] = self.flat_p[off:off + t.numel()].view(t.shape)
            self.P[n].copy_(t)
            self.gviews[n] = self.flat_g[off:off + t.numel()].view(t.shape)
        for n, t in state.items():
            if n not in self.P and "encoder_tran" not in n:
                self.P[n] = t.to(dev).clone()
        self._install_mirror()
        self.total_sq = torch.zeros(1, device=dev, dtype=F32)
        self.step_count = 0
        self.n_params = self.flat_p.numel()
        self.last_sink = None

    def _install_mirror(self):
        """Operand views the kernels can use without per-step conversion: the bf16 mirror of the flat parameter buffer
        (rewritten by the AdamW kernel) and fp32 views of parameter groups that are consumed as one matrix."""
        mir = self.k.mirror
        o = self.offsets
        f32view = lambda first, rows, cols: self.flat_p[o[first]:o[first] + rows * cols].view(rows, cols)
        mir["heads.weight"] = f32view("a_rel_compress.weight", 26, 1936)
        mir["heads.bias"] = self.flat_p[o["a_rel_compress.bias"]:o["a_rel_compress.bias"] + 26]
        mir["subjobj.bias"] = self.flat_p[o["subj_fc.bias"]:o["subj_fc.bias"] + 1024]
        if self.flat_pb is None:
            mir["subjobj.weight"] = f32view("subj_fc.weight", 1024, 2048)
            return
        from . import ops as _ops
        _ops.convert(self.flat_p.view(1, -1), torch.bfloat16, out=self.flat_pb.view(1, -1))
        bview = lambda n, shape: self.flat_pb[o[n]:o[n] + int(__import__('math').prod(shape))].view(shape)
        mir["subjobj.weight"] = bview("subj_fc.weight", (1024, 2048))
        for n in self.param_names:
            t = self.P[n]
            if t.dim() == 2 and t.shape[1] % 8 == 0 and t.shape[0] >= 64 and not n.endswith("_rel_compress.weight") \
                    and n not in ("vr_fc.weight", "subj_fc.weight", "obj_fc.weight", "obj_embed.weight", "obj_embed2.weight",
                                  "object_classifier.obj_embed.weight", "object_classifier.pos_embed.1.weight",
                                  "object_classifier.decoder_lin.3.weight"):
                mir[n] = bview(n, tuple(t.shape))
        mir["union_func1.weight"] = bview("union_func1.weight", (256, 2048))

    def forward_backward(self, batch: M.Batch, plan=None):
        """batch: device-resident collated batch.  Plan + labels are (re)built from its host metadata every call —
        they replace the reference's per-frame python loops and are part of the step."""
        dsg = self.arch == "dsg"
        if plan is None:
            plan = M.make_plan(batch, self.dev, self.mode, dsg, with_labels=True)
        labels = plan.labels
        fwd = M.dsg_forward if dsg else M.sttran_forward
        bwd = M.dsg_backward if dsg else M.sttran_backward
        out, ctx = fwd(self.k, self.P, batch, plan, self.mode, True, True)
        loss, d26, dobj = M.fused_loss(out, batch, labels, self.mode)
        self.last_sink = GradSink(self.gviews)
        bwd(self.k, self.P, batch, plan, self.mode, ctx, d26, dobj, grads=self.last_sink)
        return loss, out

    def optimizer_step(self):
        D.allreduce_mean_(self.flat_g)      # no-op on one rank
        self.step_count += 1
        self.total_sq.zero_()
        ops.sumsq(self.flat_g, self.total_sq)
        ops.adamw_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps,
                       self.wd, self.step_count, self.total_sq, self.max_norm, p_bf16=self.flat_pb)
        self.k._wcache.clear()   # derived operand copies (permuted / padded weights) are stale now; mirrored ones are not

    def step(self, batch: M.Batch):
        loss, _ = self.forward_backward(batch)
        self.optimizer_step()
        return loss

    def step_from_host(self, host_batch: M.Batch):
        """End-to-end step: pinned host buffers -> device inside the step (copies on the compute stream)."""
        return self.step(M.upload(host_batch, self.dev))

    # ---- input pipelining: the next batch's H2D copies run on a side stream while this batch computes ----
    def prefetch(self, host_batch: M.Batch):
        """Enqueue, on the copy stream, everything step needs from the host: the batch descriptors + labels (one small
        pinned buffer, first) and the input tensors.  Nothing is uploaded on the compute stream afterwards — a small
        copy there would queue behind this multi-GB transfer in the copy engine and stall the whole step."""
        if not hasattr(self, "copy_stream"):
            self.copy_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            plan = M.make_plan(host_batch, self.dev, self.mode, self.arch == "dsg", with_labels=True, consumer_stream=main)
            b = M.upload(host_batch, self.dev, rasterise=False, consumer_stream=main)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return b, plan, ev

    def step_pipelined(self, handle, next_host=None):
        """Run the step whose inputs were prefetched (`handle`) and start the transfers of the next batch.
        Returns (loss, next handle)."""
        b, plan, ev = handle
        nxt = self.prefetch(next_host) if next_host is not None else None
        torch.cuda.current_stream(self.dev).wait_event(ev)
        loss, _ = self.forward_backward(M.ensure_masks(b), plan)
        self.optimizer_step()
        return loss, nxt
