"""Fused training step: forward, fused losses, hand-written backward (ONE C call each, csrc/step.cu), NCCL gradient
all-reduce, global-norm clip and AdamW over flat fp32 parameter / gradient / moment buffers — the kernel-level equivalent
of one iteration of tools/train_STTran.py:129-195 for a batch of videos (loss = mean over videos of the per-video loss).

Reference behaviours kept:
  * `check_valid_iter` (lib/utils.py:3-11, train_STTran.py:191): a step whose loss or gradient norm is not finite is
    skipped — decided ON THE DEVICE (no host sync), agreed across ranks (the flag is max-reduced; NaN gradients poison
    the all-reduced norm on every rank alike); `Trainer.steps_applied()` / `steps_skipped()` read the counters back.
  * lib/AdamW.py:66 skips parameters without a gradient: parameters the mode never reaches are not in the flat buffers
    (object classifier in predcls); when a batch has no sliding window at all (every video a single frame) the temporal
    decoder and the frame position embedding are left untouched for that step.
  * tools/train_STTran.py:148-150: a pair with several attention labels trains on one drawn at random per step
    (`label_rng`; None = the first label, deterministic).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _C
from . import dist as D
from . import engine as E
from . import model as M
from . import ops

F32 = torch.float32

_OP_SUFFIX = ("in_proj_weight", "out_proj.weight", "linear1.weight", "linear2.weight")
_OP_NAMES = ("object_classifier.decoder_lin.0.weight", "union_func1.weight", "subj_fc.weight", "obj_fc.weight")


class Trainer:
    def __init__(self, state: Dict[str, torch.Tensor], mode: str = "sgdet", arch: str = "sttran", precision: str = "bf16",
                 lr: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 5.0,
                 device="cuda", dropout: float = 0.0, label_seed: Optional[int] = None):
        dev = torch.device(device)
        self.mode, self.arch, self.dev = mode, arch, dev
        self.k = E.Kernels(precision, dropout=dropout)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.label_rng = np.random.default_rng(label_seed) if label_seed is not None else None
        state = {n: t for n, t in state.items() if "encoder_tran" not in n and not n.startswith("object_classifier.positional_encoder")}
        probe = E.ModelDesc(self.k, state, arch, mode)
        self.param_names = list(probe.grad_names)
        self.offsets: Dict[str, int] = dict(probe.grad_off)
        total = probe.grad_elems
        self.flat_p = torch.zeros(total, device=dev, dtype=F32)
        self.flat_g = torch.zeros(total, device=dev, dtype=F32)
        self.flat_m = torch.zeros(total, device=dev, dtype=F32)
        self.flat_v = torch.zeros(total, device=dev, dtype=F32)
        self.flat_pb = torch.zeros(total, device=dev, dtype=torch.bfloat16) if precision == "bf16" else None
        self.P: Dict[str, torch.Tensor] = {}
        self.gviews: Dict[str, torch.Tensor] = {}
        for n in self.param_names:
            t, off = state[n], self.offsets[n]
            self.P[n] = self.flat_p[off:off + t.numel()].view(t.shape)
            self.P[n].copy_(t)
            self.gviews[n] = self.flat_g[off:off + t.numel()].view(t.shape)
        for n, t in state.items():      # buffers (BN running statistics, positional encoding) and unreached parameters
            if n not in self.P:
                self.P[n] = t.to(dev).clone().contiguous()
        self._install_mirror()
        self.desc = E.ModelDesc(self.k, self.P, arch, mode, static=True)
        assert self.desc.grad_off == self.offsets
        self._goff = self.desc.grad_offsets()
        self.total_sq = torch.zeros(1, device=dev, dtype=F32)
        self.state = torch.zeros(2, device=dev, dtype=torch.int32)     # [steps applied, steps skipped]
        self.skip_flag = torch.zeros(1, device=dev, dtype=torch.int32)
        self.n_params = self.flat_p.numel()
        # the parameters a window-less batch does not reach: position embedding + temporal decoder (contiguous slot ranges)
        self._tail_ranges = None
        if arch == "sttran":
            pos = "glocal_transformer.position_embedding.weight"
            dec0 = [n for n in self.param_names if n.startswith("glocal_transformer.global_attention.")]
            if dec0:
                self._tail_ranges = ((self.offsets[pos], self.offsets[pos] + (self.P[pos].numel() + 7) // 8 * 8), (self.offsets[dec0[0]], total))
        # gradient all-reduce under the backward pass (world > 1): the temporal / global layers are the tail of the flat gradient
        # buffer and the first to be finished by the backward pass; the sequencer calls back when their last kernel is enqueued
        # and their buckets go out while the spatial encoder, the pair-feature stack and the object head are still running
        tail = [n for n in self.param_names if n.startswith("glocal_transformer.global_attention." if arch == "sttran" else "global_transformer.")]
        self._tail_off = self.offsets[tail[0]] if tail else None
        self._tail_works = None
        self._hook_error = None
        self._hook_cb = ctypes.CFUNCTYPE(None, ctypes.c_void_p)(self._on_tail_grads)
        import os
        self.overlap_allreduce = os.environ.get("NLV_ALLREDUCE_OVERLAP", "1") != "0"      # debugging switch: 0 = one all-reduce after backward
        self.last = None
        from .plan import Stager
        self.stager = Stager(dev)
        self._loss_ring, self._loss_i = torch.zeros(8, device=dev, dtype=F32), 0
        # every step's loss is also copied to pinned host memory right behind the kernel that produced it (stream order), so
        # a training loop can read step i's value while step i+1 is already enqueued: `loss_value(ticket)`
        self._loss_host = torch.zeros(8, dtype=F32).pin_memory() if dev.type == "cuda" else torch.zeros(8, dtype=F32)
        self._loss_ev = [None] * 8
        self.last_ticket = None

    def _install_mirror(self):
        """bf16 operand mirror of the flat parameter buffer (rewritten by the AdamW kernel): the GEMM weights are used from
        it without a per-step conversion."""
        if self.flat_pb is None:
            return
        ops.convert(self.flat_p.view(1, -1), torch.bfloat16, out=self.flat_pb.view(1, -1))
        for n in self.param_names:
            if n in _OP_NAMES or (n.endswith(_OP_SUFFIX) and ("attention" in n or "transformer" in n)):
                o = self.offsets[n]
                self.k.mirror[n] = self.flat_pb[o:o + self.P[n].numel()].view(self.P[n].shape[0], -1)

    def _on_tail_grads(self, _user):
        try:
            self._tail_works = D.allreduce_sum_async(self.flat_g[self._tail_off:])
        except Exception as e:      # an exception cannot cross the C frame: re-raised by optimizer_step
            self._hook_error = e

    # ---- one step ------------------------------------------------------------------------------------------------
    def forward_backward(self, batch: M.Batch, plan=None):
        """batch: device-resident collated batch.  Plan + labels are (re)built from its host metadata every call —
        they replace the reference's per-frame python loops and are part of the step."""
        dsg = self.arch == "dsg"
        if plan is None:
            plan = M.make_plan(batch, self.dev, self.mode, dsg, with_labels=True, label_rng=self.label_rng, stager=self.stager)
        self.k.seed += 1
        out, sess = E.run_forward(self.k, self.desc, self.P, batch, plan, True, True, labels=plan.labels, with_loss=True,
                                  with_backward=True, grad_base=self.flat_g, grad_offsets=self._goff)
        self._tail_works = None
        if self.overlap_allreduce and self._tail_off is not None and D.world() > 1:
            _C.check(_C.lib().nlv_session_set_tail_hook(ctypes.c_void_p(sess.h), self._hook_cb, None), "session_set_tail_hook")
        E.run_backward(sess)
        plan.consumed()
        self.last = (out, plan)
        # the workspace is reused by the next step: the loss a caller may hold on to goes to a small ring of its own
        slot = self._loss_ring[self._loss_i % 8:self._loss_i % 8 + 1]
        self._loss_i += 1
        ops.convert(out["loss"].view(1, 1), F32, out=slot.view(1, 1))
        j = (self._loss_i - 1) % 8
        self._loss_host[j:j + 1].copy_(slot, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._loss_ev[j] = ev
        self.last_ticket = (j, ev)
        return slot, out

    def loss_value(self, ticket=None) -> float:
        """Host value of a step's loss (ticket = `last_ticket` taken right after that step was enqueued).  Waits only for
        that step's device-to-host copy, not for work enqueued after it."""
        j, ev = ticket if ticket is not None else self.last_ticket
        ev.synchronize()
        return float(self._loss_host[j])

    def optimizer_step(self):
        lib = _C.lib()
        st = ops._stream()
        out, plan = self.last if self.last is not None else (None, None)
        w = D.world()
        if out is not None:
            _C.check(lib.nlv_flag_nonfinite(ops._ptr(out["loss"]), 1, ops._ptr(self.skip_flag), st), "flag_nonfinite")
        if self._hook_error is not None:
            e, self._hook_error = self._hook_error, None
            raise e
        if self._tail_works is not None:       # the tail is already on its way: reduce the head (+ the skip flag), then wait for both
            D.allreduce_sum_(self.flat_g[:self._tail_off], flag=self.skip_flag)
            for wk in self._tail_works:
                wk.wait()
            self._tail_works = None
        else:
            D.allreduce_sum_(self.flat_g, flag=self.skip_flag)      # no-op on one rank; the mean's 1/world is folded into AdamW
        ops.zero_(self.total_sq)
        ops.sumsq(self.flat_g, self.total_sq)
        ranges = [(0, self.n_params)]
        if self._tail_ranges is not None and plan is not None and plan.Mg == 0:
            (p0, p1), (d0, _) = self._tail_ranges
            ranges = [(0, p0), (p1, d0)]
        for a, b in ranges:
            if b <= a:
                continue
            pb = ctypes.c_void_p(self.flat_pb.data_ptr() + 2 * a) if self.flat_pb is not None else None
            _C.check(lib.nlv_adamw_step_state(
                ctypes.c_void_p(self.flat_p.data_ptr() + 4 * a), ctypes.c_void_p(self.flat_g.data_ptr() + 4 * a),
                ctypes.c_void_p(self.flat_m.data_ptr() + 4 * a), ctypes.c_void_p(self.flat_v.data_ptr() + 4 * a),
                ctypes.c_longlong(b - a), ctypes.c_float(self.lr), ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]),
                ctypes.c_float(self.eps), ctypes.c_float(self.wd), ops._ptr(self.state), ops._ptr(self.total_sq), ops._ptr(self.skip_flag),
                ctypes.c_float(1.0 / w), ctypes.c_float(self.max_norm), pb, st), "adamw_step_state")
        _C.check(lib.nlv_adamw_finish(ops._ptr(self.state), ops._ptr(self.total_sq), ops._ptr(self.skip_flag), st), "adamw_finish")

    def steps_applied(self) -> int:
        return int(self.state[0].item())

    def steps_skipped(self) -> int:
        return int(self.state[1].item())

    @property
    def step_count(self) -> int:
        return self.steps_applied()

    def step(self, batch: M.Batch):
        loss, _ = self.forward_backward(batch)
        self.optimizer_step()
        return loss

    def step_from_host(self, host_batch: M.Batch):
        """End-to-end step: pinned host buffers -> device inside the step (copies on the compute stream)."""
        return self.step(M.upload(host_batch, self.dev, rasterise=False))

    # ---- input pipelining: the next batch's H2D copies run on a side stream while this batch computes ----
    def prefetch(self, host_batch: M.Batch):
        """Enqueue, on the copy stream, everything step needs from the host: the batch descriptors + labels (one small
        pinned buffer, first) and the input tensors.  Nothing is uploaded on the compute stream afterwards — a small
        copy there would queue behind this multi-GB transfer in the copy engine and stall the whole step."""
        if not hasattr(self, "copy_stream"):
            self.copy_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            plan = M.make_plan(host_batch, self.dev, self.mode, self.arch == "dsg", with_labels=True, consumer_stream=main,
                               label_rng=self.label_rng, stager=self.stager)
        b = M.upload_from_side_stream(host_batch, self.dev, self.copy_stream)     # memory from the compute stream's pool
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        return b, plan, ev

    def step_pipelined(self, handle, next_host=None):
        """Run the step whose inputs were prefetched (`handle`) and start the transfers of the next batch.
        Returns (loss, next handle)."""
        b, plan, ev = handle
        nxt = self.prefetch(next_host) if next_host is not None else None
        torch.cuda.current_stream(self.dev).wait_event(ev)
        loss, _ = self.forward_backward(b, plan)
        self.optimizer_step()
        return loss, nxt
