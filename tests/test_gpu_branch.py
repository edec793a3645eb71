"""Non-weakly-supervised sgdet TEST branch (lib/sttran.py:185-283) through the drop-in module on CUDA: detections ->
per-class NMS -> labels / human / pairs -> union RoIAlign + masks -> relation heads.

  * entry production vs golden vectors written by the reference's own ObjectClassifier(is_wks=False).eval()
    (oracle/make_golden_branch.py): indices, boxes, scores, union boxes bit-exact; RoIAlign / mask digests exact.
    The golden was produced with the reference's CPU nms (IoU >= thr suppresses), so the module runs with
    nms_strict=False here; the default (True) is the reference's CUDA rule and is checked against the C oracle.
  * the relation outputs of the full forward vs the CPU oracle model on the same produced entry.
"""
import copy

import numpy as np
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G
from tests.test_cpu_branch import CASES, check_against_golden

pytestmark = pytest.mark.gpu


def _module(precision="fp32"):
    from nlvsgg_b200.lib.sttran import STTran
    m = STTran("sgdet", 3, 6, 17, synth.AG_OBJECT_CLASSES, 1, 3, "wk", False, 2048, precision=precision)
    m.load_state_dict(synth.make_state_dict(G.sttran_template(), 9))
    return m.cuda().eval()


def _cuda(entry):
    return {k: (v.cuda() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in entry.items()}


@pytest.mark.parametrize("name", CASES)
def test_branch_entry_production_matches_reference_golden(cuda_lib, name):
    gold = G.load_case(name)
    m = _module()
    m.object_classifier.nms_strict = False
    out = m.object_classifier.sgdet_test_branch(_cuda(synth.synth_detections(**gold["cfg"])))
    host = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in out.items()}
    check_against_golden({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in host.items()}, gold)


@pytest.mark.parametrize("strict", [True, False])
def test_branch_matches_oracle_for_both_nms_rules(cuda_lib, strict):
    from oracle import detector_branch as DB
    entry = synth.synth_detections(21, frames=6, mean_boxes=12, fmap_channels=32, fmap_hw=(17, 30))
    want = DB.sgdet_test_branch({k: v.numpy() for k, v in entry.items()}, nms_strict=strict)
    m = _module()
    m.object_classifier.nms_strict = strict
    out = m.object_classifier.sgdet_test_branch(_cuda(entry))
    for k in ("pred_labels", "pair_idx", "human_idx"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k
    for k in ("boxes", "distribution", "pred_scores", "im_idx", "union_box", "union_feat", "spatial_masks"):
        assert np.array_equal(out[k].cpu().numpy(), want[k].astype(np.float32)), k


def test_full_forward_of_the_branch_matches_oracle_model(cuda_lib):
    from oracle import detector_branch as DB, model as omodel
    entry = synth.synth_detections(31, frames=5, mean_boxes=8, fmap_channels=2048, fmap_hw=(17, 30))
    m = _module("fp32")
    with torch.no_grad():
        pred = m(_cuda(entry))
    prod = DB.sgdet_test_branch({k: v.numpy() for k, v in entry.items()}, nms_strict=True)
    oentry = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in prod.items()}
    oentry["labels"], oentry["scores"] = oentry["pred_labels"], oentry["pred_scores"]
    sd = synth.make_state_dict(G.sttran_template(), 9)
    with torch.no_grad():
        want = omodel.sttran_forward(sd, oentry, "predcls", training=False)
    assert np.array_equal(pred["pair_idx"].cpu().numpy(), prod["pair_idx"])
    for k in ("attention_distribution", "spatial_distribution", "contacting_distribution"):
        assert G.rel_err(pred[k].float().cpu(), want[k]) < 1e-4, k
    # the detector's distribution is kept (the classifier head is not applied on this branch)
    assert np.array_equal(pred["distribution"].cpu().numpy(), prod["distribution"])


@pytest.mark.parametrize("name", ["sgcls_test_branch_a", "sgcls_test_branch_b"])
def test_sgcls_test_branch_matches_reference_golden(cuda_lib, name):
    """STTran(mode='sgcls').eval() (lib/sttran.py:105-170) against goldens written by the reference itself with the
    un-vendored union-feature extractor replaced by a stand-in (oracle/make_golden_r2.py): the classifier head stops the
    sequencer after the logits, the branch builds labels / humans / duplicate clean-up / pairs / union boxes / masks on the
    device (all exact), the relation path then runs on the inferred labels."""
    from nlvsgg_b200.lib.sttran import STTran
    from oracle.make_golden_r2 import sgcls_entry, standin_union_features
    case = G.load_case(name)
    entry = sgcls_entry(case["seed"], case["frames"], case["k"])
    m = STTran("sgcls", 3, 6, 17, synth.AG_OBJECT_CLASSES, 1, 3, "wk", True, 2048, precision="fp32")
    m.load_state_dict(synth.make_state_dict(G.sttran_template(), case["seed"]))
    m = m.cuda().eval()
    fm = entry["fmaps"]
    m.object_classifier.union_feature_extractor = lambda e, f, boxes: standin_union_features(fm, f, boxes.cpu()).cuda()
    with torch.no_grad():
        out = m(_cuda({k: v for k, v in entry.items() if k != "fmaps"}))
    want = case["outputs"]
    for k in ("pred_labels", "pair_idx", "im_idx", "union_box"):
        assert torch.equal(out[k].cpu(), want[k]), k
    assert torch.equal(out["spatial_masks"].cpu().float(), want["spatial_masks"].float())
    assert tuple(out["union_feat"].shape) == tuple(want["union_feat_shape"])
    for k in ("distribution", "pred_scores", "attention_distribution", "spatial_distribution", "contacting_distribution"):
        assert G.rel_err(out[k].cpu(), want[k]) < 1e-3, k
