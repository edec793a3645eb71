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


def _book_loop(ev, masks, rels_per_frame):
    """The per-frame / per-bit bookkeeping of lib/evaluation_recall.py:203-204,69-87,146-165 written as plain loops."""
    from nlvsgg_b200.lib.evaluation_recall import KS, _bits
    m = ev.mode
    for f in range(masks.shape[0]):
        rel = rels_per_frame[f]
        G = rel.shape[0]
        for pi, key in enumerate(("_recall", "_recall_nogc", "_semi_recall")):
            for ki, k in enumerate(KS):
                n = int(sum(bin(int(w)).count("1") for w in masks[f, pi, ki]))
                ev.result_dict[m + key][k].append(float(n) / float(G))
        for pi, key in ((0, "_mean_recall"), (1, "_ng_mean_recall")):
            for ki, k in enumerate(KS):
                hit, cnt = [0] * ev.num_rel, [0] * ev.num_rel
                for g in range(G):
                    cnt[int(rel[g, 2])] += 1
                    cnt[0] += 1
                for g in _bits(masks[f, pi, ki]):
                    hit[int(rel[g, 2])] += 1
                    hit[0] += 1
                for n in range(ev.num_rel):
                    if cnt[n] > 0:
                        ev.result_dict[m + key + "_collect"][k][n].append(float(hit[n] / cnt[n]))


def test_vectorised_booking_equals_the_reference_loops():
    """The numpy bookkeeping of the CUDA evaluator's host side (no GPU needed: it consumes integer match sets)."""
    import numpy as np
    from nlvsgg_b200.lib.evaluation_recall import PackedGT, SceneGraphEvaluator
    mk = lambda: SceneGraphEvaluator("sgdet", synth.AG_OBJECT_CLASSES, synth.AG_RELATIONS, synth.AG_ATTENTION, synth.AG_SPATIAL,
                                     synth.AG_CONTACTING, iou_threshold=0.5, constraint="with")
    a, b = mk(), mk()
    a.register_container(); b.register_container()
    _, gt = synth.synth_video(77, 9, 6, "sgdet", draw_fn=None, union_feat=False)
    rel, cls, box, nrel, nbox = PackedGT.pack_video(gt, a._ia, a._is, a._ic)
    # pack_video == the oracle's per-frame packing
    names = synth.AG_RELATIONS
    o = 0
    for f, fg in enumerate(gt):
        gb, gc, gr = oe.build_frame_gt(fg, synth.AG_ATTENTION, synth.AG_SPATIAL, synth.AG_CONTACTING, names)
        assert np.array_equal(rel[o:o + len(gr)], gr.astype(np.int32)) and nrel[f] == len(gr) and nbox[f] == len(gc)
        o += len(gr)
    rng = np.random.default_rng(3)
    F = len(gt)
    rel_off = np.concatenate(([0], np.cumsum(nrel)))
    masks = np.zeros((F, 3, 3, 8), dtype=np.uint32)
    for f in range(F):
        for pi in range(3):
            for ki in range(3):
                sel = rng.random(int(nrel[f])) < 0.4
                for g in np.nonzero(sel)[0]:
                    masks[f, pi, ki, g // 32] |= np.uint32(1) << np.uint32(g % 32)
    a._book(masks, {"rel": rel, "rel_off": rel_off})
    _book_loop(b, masks, [rel[rel_off[f]:rel_off[f + 1]] for f in range(F)])
    a.calculate_mean_recall(); b.calculate_mean_recall()
    assert a.result_dict == b.result_dict
