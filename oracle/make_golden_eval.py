"""TEST INFRASTRUCTURE — evaluator golden vectors (called from oracle/make_golden.py; build container only).

For seeded synthetic videos: random relation logits -> (softmax | sigmoid), synthetic GT; the reference
SceneGraphEvaluator (as shipped, and with the stable-sort patches) produces the per-frame recall lists.
Inputs are small, so both inputs and outputs are stored (tests/golden/eval_*.pt)."""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from nlvsgg_b200 import synth
from oracle import ref_harness as H

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("eval_predcls", "predcls", 41, 12, 6, 0.0, False), ("eval_sgdet", "sgdet", 42, 12, 6, 0.2, False),
         ("eval_sgdet_saturated", "sgdet", 43, 8, 5, 0.0, True), ("eval_predcls_tiny", "predcls", 44, 6, 3, 0.0, True)]


synth_pred = synth.synth_pred      # the generator lives with the other synthetic inputs (nlvsgg_b200/synth.py)


def run_reference_eval(evmod, mode, pred, gt):
    ev = evmod.SceneGraphEvaluator(mode=mode, AG_object_classes=synth.AG_OBJECT_CLASSES, AG_all_predicates=synth.AG_RELATIONS,
                                   AG_attention_predicates=synth.AG_ATTENTION, AG_spatial_predicates=synth.AG_SPATIAL,
                                   AG_contacting_predicates=synth.AG_CONTACTING, iou_threshold=0.5, constraint="with")
    ev.register_container()
    ev.evaluate_scene_graph(gt, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in pred.items()})
    ev.calculate_mean_recall()
    return copy.deepcopy(ev.result_dict)


def main(ref=None):
    ref = ref or H.load_reference()
    stable = H.load_stable_evaluator(ref)
    for (name, mode, seed, frames, k, ep, sat) in CASES:
        pred, gt = synth_pred(mode, seed, frames, k, ep, sat)
        shipped = run_reference_eval(ref.evaluation_recall, mode, pred, gt)
        canon = run_reference_eval(stable, mode, pred, gt)
        diff = sum(int(a != b) for t in ("_recall", "_recall_nogc", "_semi_recall") for kk in (10, 20, 50)
                   for a, b in zip(shipped[mode + t][kk], canon[mode + t][kk]))
        torch.save({"name": name, "mode": mode, "seed": seed, "frames": frames, "mean_boxes": k, "empty_frame_prob": ep,
                    "saturate": sat, "pred": pred, "gt": gt, "result_shipped": shipped, "result_canonical": canon,
                    "frames_differing_between_shipped_and_canonical_sort": diff}, os.path.join(GOLDEN, name + ".pt"))
        r20 = np.mean(canon[mode + "_recall"][20])
        print(f"wrote {name}: frames={len(gt)} R@20={r20:.4f} nogc R@20={np.mean(canon[mode + '_recall_nogc'][20]):.4f} "
              f"shipped-vs-canonical differing entries: {diff}")


if __name__ == "__main__":
    main()
