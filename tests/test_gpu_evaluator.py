"""Recall@K kernel (through the drop-in SceneGraphEvaluator) vs golden vectors from the reference evaluator and vs the
numpy oracle: per-frame recall floats and mean-recall collections must be bit-identical."""
import copy

import numpy as np
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G
from tests.test_cpu_evaluator import EVAL_CASES, assert_same_results, make_oracle

pytestmark = pytest.mark.gpu


def make_cuda_eval(mode):
    from nlvsgg_b200.lib.evaluation_recall import SceneGraphEvaluator
    ev = SceneGraphEvaluator(mode=mode, AG_object_classes=synth.AG_OBJECT_CLASSES, AG_all_predicates=synth.AG_RELATIONS,
                             AG_attention_predicates=synth.AG_ATTENTION, AG_spatial_predicates=synth.AG_SPATIAL,
                             AG_contacting_predicates=synth.AG_CONTACTING, iou_threshold=0.5, constraint="with")
    ev.register_container()
    return ev


def to_cuda(pred):
    return {k: (v.cuda() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in pred.items()}


@pytest.mark.parametrize("name", EVAL_CASES)
def test_cuda_evaluator_matches_reference_golden(cuda_lib, name):
    case = G.load_case(name)
    ev = make_cuda_eval(case["mode"])
    ev.evaluate_scene_graph(case["gt"], to_cuda(case["pred"]))
    ev.calculate_mean_recall()
    assert_same_results(ev.result_dict, case["result_canonical"], case["mode"])


@pytest.mark.parametrize("mode,seed,frames,k,ep", [("predcls", 101, 30, 7, 0.1), ("sgdet", 102, 40, 9, 0.0), ("sgdet", 103, 5, 3, 0.5),
                                                    ("predcls", 104, 25, 20, 0.0)])
def test_cuda_evaluator_matches_oracle(cuda_lib, mode, seed, frames, k, ep):
    from oracle.make_golden_eval import synth_pred
    pred, gt = synth_pred(mode, seed, frames, k, ep, saturate=(seed % 2 == 0))
    dpred = to_cuda(pred)
    ev = make_cuda_eval(mode)
    ev.evaluate_scene_graph(gt, dpred)
    ev.calculate_mean_recall()
    opred = {kk: (v.cpu() if torch.is_tensor(v) else v) for kk, v in dpred.items()}   # softmaxed by the CUDA evaluator, as :400 does
    oe = make_oracle(mode)
    oe.evaluate_scene_graph(gt, opred)
    oe.calculate_mean_recall()
    assert_same_results(ev.result_dict, oe.result_dict, mode)


def test_batched_videos_one_launch_equals_per_video(cuda_lib):
    from oracle.make_golden_eval import synth_pred
    vids = [synth_pred("sgdet", 200 + i, 6 + i, 5, 0.2, False) for i in range(5)]
    a, b = make_cuda_eval("sgdet"), make_cuda_eval("sgdet")
    for pred, gt in vids:
        a.evaluate_scene_graph(gt, to_cuda(pred))
    b.evaluate_videos([(gt, to_cuda(pred)) for pred, gt in vids])
    a.calculate_mean_recall(); b.calculate_mean_recall()
    assert_same_results(a.result_dict, b.result_dict, "sgdet")


def test_empty_prediction_frames_give_zero_recall(cuda_lib):
    from oracle.make_golden_eval import synth_pred
    pred, gt = synth_pred("sgdet", 300, 4, 4, 0.0, False)
    empty = {k: (v[:0] if torch.is_tensor(v) and k not in ("boxes", "labels", "scores", "pred_labels", "pred_scores", "features", "distribution") else v)
             for k, v in pred.items()}
    ev = make_cuda_eval("sgdet")
    ev.evaluate_scene_graph(gt, to_cuda(empty))
    assert ev.result_dict["sgdet_recall"][20] == [0.0] * len(gt)
