"""Oracle evaluator (numpy restatement) against golden vectors from the reference SceneGraphEvaluator
(canonical tie order = the reference with its sorts made stable; see oracle/evaluator.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from nlvsgg_b200 import synth
from oracle import evaluator as oe
from tests import golden_util as G

EVAL_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(G.GOLDEN, "eval_*.pt")))


def make_oracle(mode):
    ev = oe.Evaluator(mode, synth.AG_OBJECT_CLASSES, synth.AG_RELATIONS, synth.AG_ATTENTION, synth.AG_SPATIAL,
                      synth.AG_CONTACTING, iou_threshold=0.5, constraint="with")
    ev.register_container()
    return ev


def assert_same_results(got, want, mode):
    for t in ("_recall", "_recall_nogc", "_semi_recall"):
        for k in (10, 20, 50):
            assert got[mode + t][k] == want[mode + t][k], (t, k)          # per-frame floats, bit-identical
    for t in ("_mean_recall", "_ng_mean_recall"):
        for k in (10, 20, 50):
            assert got[mode + t + "_collect"][k] == want[mode + t + "_collect"][k], (t, k)
            assert got[mode + t][k] == want[mode + t][k]


@pytest.mark.parametrize("name", EVAL_CASES)
def test_oracle_evaluator_matches_reference(name):
    case = G.load_case(name)
    mode = case["mode"]
    pred = dict(case["pred"])
    pred["attention_distribution"] = torch.softmax(pred["attention_distribution"], dim=1)   # evaluation_recall.py:400
    ev = make_oracle(mode)
    ev.evaluate_scene_graph(case["gt"], pred)
    ev.calculate_mean_recall()
    assert_same_results(ev.result_dict, case["result_canonical"], mode)


def test_perfect_predictions_give_full_recall():
    """Known answer (lib/assign_pseudo_label.py:1391-1415 entry_to_pred idea): GT boxes + one-hot GT predicates."""
    entry, gt = synth.synth_video(9, 5, 4, "predcls", union_feat=False)
    R = entry["pair_idx"].shape[0]
    att = torch.full((R, 3), 1e-4); spa = torch.full((R, 6), 0.01); con = torch.full((R, 17), 0.01)  # exact zeros would hit the row-type quirk of :282-296
    for r in range(R):
        att[r, entry["attention_gt"][r][0]] = 1.0
        spa[r, entry["spatial_gt"][r]] = 0.9
        con[r, entry["contacting_gt"][r]] = 0.9
    pred = {k: v for k, v in entry.items() if torch.is_tensor(v)}
    pred.update(attention_distribution=att / att.sum(1, keepdim=True), spatial_distribution=spa, contacting_distribution=con)
    ev = make_oracle("predcls")
    ev.evaluate_scene_graph(gt, pred)
    assert all(v == 1.0 for v in ev.result_dict["predcls_recall_nogc"][50])
    assert all(v == 1.0 for v in ev.result_dict["predcls_semi_recall"][50])
