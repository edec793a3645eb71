"""TEST INFRASTRUCTURE — CPU timing of the oracle restatement (bench.py's cpu_baseline / --impl reference legs).

One CPU "step" = what one iteration of tools/train_STTran.py:129-195 does for each video of the sample (forward,
CE+CE+BCE+BCE loss, backward), followed by one clip_grad_norm_(5) + AdamW (lib/AdamW.py:52-114) update; gradients of
the sample's videos are averaged, as the CUDA step does for its batch.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List

import torch

from oracle import model as omodel


def adamw_update(params: Dict[str, torch.Tensor], state: dict, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, wd=1e-2, max_norm=5.0):
    """clip_grad_norm_ + lib/AdamW.py:52-114 (decay applied before the moment update, :69)."""
    grads = [p.grad for p in params.values() if p.grad is not None]
    total = torch.norm(torch.stack([torch.norm(g.detach(), 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    step_size = lr * math.sqrt(bc2) / bc1
    with torch.no_grad():
        for n, p in params.items():
            if p.grad is None:
                continue
            g = p.grad * coef
            p.mul_(1 - lr * wd)
            m = state.setdefault("m." + n, torch.zeros_like(p))
            v = state.setdefault("v." + n, torch.zeros_like(p))
            m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
            v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
            p.add_(m.div(v.sqrt().add_(eps)).mul_(-step_size))
            p.grad = None


def cpu_train_step(sd: Dict[str, torch.Tensor], entries: List[dict], mode: str, arch: str, opt_state: dict) -> float:
    """Runs one CPU step over `entries`; returns the loss (mean over videos)."""
    fwd = omodel.sttran_forward if arch == "sttran" else omodel.dsg_forward
    params = {k: v for k, v in sd.items() if v.is_floating_point() and "running_" not in k and not k.endswith(".pe")
              and "encoder_tran" not in k}
    for p in params.values():
        p.requires_grad_(True)
    total = 0.0
    for e in entries:
        pred = fwd(sd, e, mode, training=True)
        loss = omodel.training_loss(pred, e, mode) / len(entries)
        loss.backward()
        total += float(loss.detach())
    adamw_update(params, opt_state)
    return total


def time_cpu_steps(sd, entries, mode, arch, steps: int, warmup: int):
    """Returns (seconds per step, frames per step)."""
    frames = sum(int(e["im_idx"].max().item()) + 1 if e["im_idx"].numel() else 0 for e in entries)
    st = {}
    for _ in range(warmup):
        cpu_train_step(sd, entries, mode, arch, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_train_step(sd, entries, mode, arch, st)
    return (time.perf_counter() - t0) / max(steps, 1), frames
