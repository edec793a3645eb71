"""TEST INFRASTRUCTURE — pins the oracle restatement against the reference itself.

Runs only in the build container (needs /root/reference).  Compares oracle/model.py,
oracle/cref.c and oracle/evaluator.py with the reference's own code on seeded synthetic
inputs and prints a table; exits non-zero on any mismatch.  The same comparisons are frozen
into tests/golden/ by oracle/make_golden.py so they can be replayed where the reference is absent.

    python -m oracle.validate_oracle
"""
from __future__ import annotations

import copy
import sys

import numpy as np
import torch

from nlvsgg_b200 import synth
from oracle import cref, model as omodel, ref_harness as H


def _to_sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def _clone_entry(e):
    return {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in e.items()}


def check_sttran(ref, mode, seed, frames, k, empty_p, training):
    m = H.build_reference_sttran(ref, mode)
    sd = synth.make_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    m.train(training)
    entry, _ = synth.synth_video(seed, frames, k, mode, draw_fn=cref.draw_union_boxes, empty_frame_prob=empty_p)
    sd2 = {k_: v.clone() for k_, v in sd.items()}
    with torch.no_grad():
        want = m(_clone_entry(entry))
        got = omodel.sttran_forward(sd2, entry, mode, training=training)
    worst = 0.0
    for key in ("attention_distribution", "spatial_distribution", "contacting_distribution", "distribution"):
        if key in got and key in want:
            d = (got[key] - want[key]).abs().max().item() / max(want[key].abs().max().item(), 1e-12)
            worst = max(worst, d)
    if training:  # running stats must have been updated identically
        after = m.state_dict()
        for key in after:
            if "running" in key:
                d = (after[key] - sd2[key]).abs().max().item()
                worst = max(worst, d)
    return worst


def check_dsg(ref, mode, seed, frames, k):
    m = H.build_reference_dsg(ref, mode)
    sd = synth.make_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    m.eval()
    entry, _ = synth.synth_video(seed, frames, k, mode, draw_fn=cref.draw_union_boxes)
    with torch.no_grad():
        e = _clone_entry(entry)
        m(e)
        got = omodel.dsg_forward(sd, entry, mode, training=False)
    worst = 0.0
    for key in ("attention_distribution", "spatial_distribution", "contacting_distribution", "distribution"):
        if key in got and key in e:
            d = (got[key] - e[key]).abs().max().item() / max(e[key].abs().max().item(), 1e-12)
            worst = max(worst, d)
    return worst


def main():
    ref = H.load_reference()
    torch.manual_seed(0)
    ok = True
    rows = []
    for (mode, seed, frames, k, ep, tr) in [("predcls", 0, 20, 6, 0.0, False), ("sgdet", 1, 20, 6, 0.0, False),
                                            ("sgdet", 2, 12, 5, 0.3, False), ("sgdet", 3, 8, 5, 0.0, True),
                                            ("sgdet", 4, 1, 6, 0.0, False), ("sgdet", 5, 9, 5, 0.4, True)]:
        w = check_sttran(ref, mode, seed, frames, k, ep, tr)
        rows.append((f"sttran {mode} seed={seed} frames={frames} empty_p={ep} train={tr}", w))
        ok &= w < 2e-5
    for (mode, seed, frames, k) in [("sgdet", 6, 12, 6)]:
        w = check_dsg(ref, mode, seed, frames, k)
        rows.append((f"dsg_detr {mode} seed={seed} frames={frames}", w))
        ok &= w < 2e-5
    for name, w in rows:
        print(f"{name:70s} max rel diff {w:.3e}")
    print("ORACLE OK" if ok else "ORACLE MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
