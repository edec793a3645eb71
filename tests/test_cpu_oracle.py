"""The oracle restatement against the golden vectors produced by the reference itself
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cref, model as omodel
from nlvsgg_b200 import synth
from tests import golden_util as G


def test_state_dict_template_matches_reference_names():
    # the golden train case lists every parameter that received a gradient in the reference
    case = G.load_case("sttran_sgdet_train")
    tmpl = G.sttran_template()
    params = {k for k in tmpl if "running_" not in k and "num_batches" not in k}
    assert set(case["grads"].keys()) | set(case["no_grad_params"]) == params
    assert sum(tmpl[k].numel() for k in params) == 103660503


def test_native_draw_union_boxes_golden():
    z = np.load(G.GOLDEN + "/native_draw_union_boxes.npz")
    assert np.array_equal(cref.draw_union_boxes(z["box_pairs"], 27), z["out"])


def test_native_bbox_overlaps_golden():
    z = np.load(G.GOLDEN + "/native_bbox_overlaps.npz")
    out = cref.bbox_overlaps(z["boxes"], z["query"])
    assert np.array_equal(out, z["out"])
    assert out[0, 0] == 1.0 and abs(out[1, 1] - 1.0 / 3.0) < 1e-15 and out[2, 2] == 0.0


def test_nms_cpu_vs_cuda_threshold_semantics():
    # two boxes with IoU exactly 0.5 under the +1 convention: kept by the CUDA rule (>), dropped by the CPU rule (>=)
    d = np.array([[0, 0, 9, 9], [0, 0, 9, 4], [100, 100, 120, 120]], dtype=np.float32)
    s = np.array([0.9, 0.8, 0.7], dtype=np.float32)
    assert cref.nms(d, s, 0.5, strict=False).tolist() == [0, 2]
    assert cref.nms(d, s, 0.5, strict=True).tolist() == [0, 1, 2]
    assert cref.nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5).tolist() == []


def test_roi_align_matches_torchvision():
    import torchvision
    rng = np.random.default_rng(0)
    inp = rng.standard_normal((2, 5, 38, 67)).astype(np.float32)
    rois = np.array([[0, 10, 20, 300, 200], [1, 0, 0, 1071, 607], [0, 5, 5, 6, 6], [1, -50, -50, 2000, 900]], np.float32)
    tv = torchvision.ops.roi_align(torch.from_numpy(inp), torch.from_numpy(rois), (7, 7), 1 / 16., 0, aligned=False).numpy()
    assert np.array_equal(cref.roi_align_forward(inp, rois, 1 / 16., 7, 7, 0), tv)


@pytest.mark.parametrize("name", G.model_cases("sttran_"))
def test_oracle_sttran_matches_reference_golden(name):
    case = G.load_case(name)
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    sd = synth.make_state_dict(G.sttran_template(), case["seed"])
    if case["training"]:
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
        pred = omodel.sttran_forward(sd, entry, case["mode"], training=True)
        loss = omodel.training_loss(pred, entry, case["mode"])
        loss.backward()
        assert abs(loss.item() - case["loss"]) <= 1e-5 * abs(case["loss"])
        for k, dg in case["grads"].items():
            g = sd[k].grad.double().flatten()
            assert abs(g.sum().item() - dg["sum"]) <= 2e-4 * dg["abs_sum"] + 1e-7, k
            assert abs(g.abs().sum().item() - dg["abs_sum"]) <= 2e-4 * dg["abs_sum"] + 1e-7, k
        for k, v in case["running"].items():
            assert G.rel_err(sd[k].detach(), v) < 1e-5, k
    else:
        with torch.no_grad():
            pred = omodel.sttran_forward(sd, entry, case["mode"], training=False)
    for k, want in case["outputs"].items():
        assert G.rel_err(pred[k].detach(), want) < 2e-5, k


def test_oracle_dsg_matches_reference_golden():
    case = G.load_case("dsg_sgdet_eval")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    # template = sttran's shared part + dsg-specific transformer names, taken from the fixture-free builder below
    sd = synth.make_state_dict(G.dsg_template(), case["seed"])
    with torch.no_grad():
        pred = omodel.dsg_forward(sd, entry, case["mode"], training=False)
    for k, want in case["outputs"].items():
        assert G.rel_err(pred[k], want) < 2e-5, k


def test_oracle_additive_int_mask_matches_reference_golden():
    """lib/transformer_wk.py:154 under torch 1.10.1 (the int key_padding_mask is ADDED to the logits): golden written by the
    reference with that reading of the mask; the plain restatement must NOT match it (the mode changes the numbers)."""
    case = G.load_case("additive_sttran_eval")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    sd = synth.make_state_dict(G.sttran_template(), case["seed"])
    with torch.no_grad():
        pred = omodel.sttran_forward(sd, entry, "sgdet", training=False, additive_mask=True)
        plain = omodel.sttran_forward(sd, entry, "sgdet", training=False)
    for k, want in case["outputs"].items():
        assert G.rel_err(pred[k], want) < 2e-5, k
    assert G.rel_err(plain["attention_distribution"], case["outputs"]["attention_distribution"]) > 1e-3


@pytest.mark.parametrize("name", ["sgcls_test_branch_a", "sgcls_test_branch_b"])
def test_oracle_sgcls_test_branch_matches_reference_golden(name):
    """lib/sttran.py:105-170 (sgcls, eval) with the un-vendored union-feature extractor replaced by the same stand-in the
    golden was written with: labels, humans, duplicate clean-up, pairs, union boxes and masks are exact."""
    from oracle.make_golden_r2 import sgcls_entry, standin_union_features
    case = G.load_case(name)
    entry = sgcls_entry(case["seed"], case["frames"], case["k"])
    sd = synth.make_state_dict(G.sttran_template(), case["seed"])
    with torch.no_grad():
        logits = omodel.object_classifier(entry, sd, "sgcls", False)["distribution"]
        got = omodel.sgcls_test_branch(entry, logits, lambda f, b: standin_union_features(entry["fmaps"], f, b), cref.draw_union_boxes)
    want = case["outputs"]
    for k in ("pred_labels", "pair_idx"):
        assert torch.equal(got[k], want[k]), k
    assert torch.equal(got["im_idx"], want["im_idx"]) and torch.equal(got["union_box"], want["union_box"])
    assert torch.equal(got["spatial_masks"].float(), want["spatial_masks"].float())
    assert G.rel_err(got["distribution"], want["distribution"]) < 1e-5 and G.rel_err(got["pred_scores"], want["pred_scores"]) < 1e-5
    assert tuple(got["union_feat"].shape) == tuple(want["union_feat_shape"])
    assert abs(got["union_feat"].double().sum().item() - want["union_feat_digest"][0].item()) <= 1e-6 * want["union_feat_digest"][1].item()


def test_oracle_transformer_mode_both_matches_reference_golden():
    """lib/transformer_wk.py mode='both' (:197-207; lib/sttran.py itself only uses 'latter'): golden written by the reference module."""
    z = G.load_case("transformer_both")
    sd = synth.make_state_dict({k: v for k, v in G.sttran_template().items() if k.startswith("glocal_transformer.")}, z["seed"])
    g = torch.Generator().manual_seed(z["seed"])
    x = torch.randn(len(z["im_idx"]), 1936, generator=g).requires_grad_(True)
    pe = sd["glocal_transformer.position_embedding.weight"].requires_grad_(True)
    out = omodel.glocal_transformer(x, z["im_idx"], sd, mode="both")
    out.square().sum().backward()
    assert G.rel_err(out.detach(), z["out"]) < 2e-5
    assert G.rel_err(x.grad, z["dx"]) < 2e-4 and G.rel_err(pe.grad, z["dpos"]) < 2e-4
    latter = omodel.glocal_transformer(x.detach(), z["im_idx"], sd)
    assert G.rel_err(latter, z["out"]) > 1e-3                     # the two modes differ on the frames inside the video
