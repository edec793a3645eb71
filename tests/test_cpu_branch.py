"""Oracle restatement of the non-wks sgdet TEST branch (oracle/detector_branch.py) vs golden vectors produced by the
reference's own ObjectClassifier(is_wks=False).eval() (oracle/make_golden_branch.py): indices bit-exact, floats exact."""
import numpy as np
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G

CASES = ["branch_sgdet_a", "branch_sgdet_b"]


def digest_close(t, d, tol=1e-6):
    t = torch.as_tensor(t)
    assert tuple(t.shape) == tuple(d["shape"])
    x = t.double()
    assert abs(float(x.sum()) - d["sum"]) <= tol * (d["abs_sum"] + 1e-30)
    assert abs(float((x * x).sum()) - d["sq_sum"]) <= tol * (d["sq_sum"] + 1e-30)
    assert torch.allclose(t.flatten()[:64].float(), d["head"].float(), rtol=tol, atol=tol)


def check_against_golden(out, gold, float_tol=0.0):
    for k in ("pred_labels", "pair_idx", "human_idx"):
        assert np.array_equal(np.asarray(out[k]), gold[k].numpy()), k
    for k in ("boxes", "distribution", "pred_scores", "im_idx", "union_box"):
        a, b = np.asarray(out[k], dtype=np.float32), gold[k].numpy()
        assert a.shape == b.shape, k
        assert np.array_equal(a, b) if float_tol == 0.0 else np.allclose(a, b, rtol=float_tol, atol=float_tol), k
    for k in ("features", "union_feat", "spatial_masks"):
        digest_close(out[k], gold[k])


@pytest.mark.parametrize("name", CASES)
def test_oracle_branch_matches_reference_golden(name):
    from oracle import detector_branch as DB
    gold = G.load_case(name)
    entry = synth.synth_detections(**gold["cfg"])
    out = DB.sgdet_test_branch({k: v.numpy() for k, v in entry.items()}, nms_strict=False)
    assert out["boxes"].shape[0] < entry["boxes"].shape[0] + 8      # NMS really removed detections
    check_against_golden(out, gold)
