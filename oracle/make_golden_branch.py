"""TEST INFRASTRUCTURE — golden vectors for the non-wks sgdet TEST branch (lib/sttran.py:185-283), produced by running the
REAL reference ObjectClassifier(is_wks=False).eval() on synthetic detector output (nlvsgg_b200/synth.py:synth_detections)
with the reference's own CPU nms (oracle/_ref) and its bit-equal RoIAlign stand-in (oracle/ref_harness.py).
    python -m oracle.make_golden_branch        # writes tests/golden/branch_sgdet_*.pt
Inputs are regenerated from the seed by the tests; only outputs are stored (large tensors as float64 digests + heads)."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nlvsgg_b200 import synth  # noqa: E402
from oracle import ref_harness as H, ref_native  # noqa: E402

CASES = {"branch_sgdet_a": dict(seed=11, frames=5, mean_boxes=9, fmap_channels=64, fmap_hw=(17, 30)),
         "branch_sgdet_b": dict(seed=12, frames=3, mean_boxes=14, fmap_channels=32, fmap_hw=(20, 34))}


def digest(t: torch.Tensor) -> dict:
    d = t.double()
    return {"shape": tuple(t.shape), "sum": float(d.sum()), "abs_sum": float(d.abs().sum()), "sq_sum": float((d * d).sum()),
            "head": t.flatten()[:64].clone()}


def main():
    ref = H.load_reference()
    ref.sttran.nms = lambda d, s, t: ref_native.nms(d, s, t)
    oc = ref.sttran.ObjectClassifier(mode="sgdet", obj_classes=ref.obj_classes, is_wks=False).eval()
    for name, cfg in CASES.items():
        entry = synth.synth_detections(**cfg)
        with torch.no_grad():
            out = oc(copy.deepcopy(entry))
        gold = {"cfg": cfg, "nms_rule": "cpu (IoU >= thr suppresses)"}
        for k in ("boxes", "distribution", "pred_labels", "pred_scores", "pair_idx", "im_idx", "human_idx", "union_box"):
            gold[k] = out[k].clone()
        for k in ("features", "union_feat", "spatial_masks"):
            gold[k] = digest(out[k])
        torch.save(gold, os.path.join(ROOT, "tests", "golden", name + ".pt"))
        print(name, "boxes", tuple(entry["boxes"].shape), "->", tuple(out["boxes"].shape), "pairs", out["pair_idx"].shape[0])


if __name__ == "__main__":
    main()
