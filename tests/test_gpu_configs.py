"""Parity at the BASELINE.json configuration shapes that are not the bench workload (C3, C4, C5).

Full-size runs cannot be compared with the CPU oracle end to end in seconds, so each test anchors a full-size CUDA run
with size-independent properties plus an oracle comparison on a slice the oracle finishes quickly:

  C4  one 200-frame x 20-box video (3800 pairs, 2-frame windows of 38 tokens: several attention chunks / work items per
      segment).  STTran's temporal decoder only looks one frame back (mode 'latter', lib/transformer_wk.py:209-215), so in
      eval mode the outputs of frames f0+1..f1 of the full video equal those of the sub-video [f0..f1]: the full CUDA run
      is compared with the ORACLE run on a 9-frame slice.
  C3  DSG-DETR sgdet, 8 videos per rank: one batched launch sequence == 8 single-video runs (per-video BatchNorm
      statistics, per-video class sequences), and one video against the oracle.
  C5  Recall@K over 1737 videos with the test split's frame-count shape (mean 31, max 121): one launch for the whole
      split, bit-identical to the numpy oracle on a sample of videos, invariant to video order, idempotent.
"""
import copy

import numpy as np
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _cuda(entry):
    return {k: (v.cuda() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in entry.items()}


def _slice_entry(entry, f0, f1):
    """Sub-video of frames f0..f1 (inclusive), frame ids rebased to 0."""
    fb = entry["boxes"][:, 0].long()
    keep_b = (fb >= f0) & (fb <= f1)
    new_index = torch.cumsum(keep_b.long(), 0) - 1
    keep_p = (entry["im_idx"].long() >= f0) & (entry["im_idx"].long() <= f1)
    out = {}
    for k in ("boxes", "labels", "scores", "features", "distribution"):
        if k in entry:
            out[k] = entry[k][keep_b].clone()
    out["boxes"][:, 0] -= f0
    out["pair_idx"] = new_index[entry["pair_idx"][keep_p]]
    out["im_idx"] = entry["im_idx"][keep_p] - f0
    for k in ("union_feat", "spatial_masks"):
        out[k] = entry[k][keep_p].clone()
    idx = torch.nonzero(keep_p).flatten().tolist()
    for k in ("attention_gt", "spatial_gt", "contacting_gt"):
        out[k] = [entry[k][i] for i in idx]
    return out, keep_p


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("bf16", 6e-2)])
def test_c4_long_video_matches_oracle_on_a_slice(cuda_lib, precision, tol):
    from oracle import cref, model as omodel
    from nlvsgg_b200 import engine as E, model as M
    frames, boxes, f0, f1 = 200, 20, 90, 98
    entry, _ = synth.synth_video(404, frames, boxes, "sgdet", draw_fn=cref.draw_union_boxes, fixed_boxes=boxes, with_gt=False)
    assert entry["pair_idx"].shape[0] == frames * (boxes - 1)          # 3800 pairs, 199 windows x 38 tokens
    sd = synth.make_state_dict(G.sttran_template(), 4)
    sub, keep_p = _slice_entry(entry, f0, f1)
    with torch.no_grad():
        want = omodel.sttran_forward(sd, sub, "sgdet", training=False)
    P = {k: v.cuda() for k, v in sd.items()}
    k = E.Kernels(precision)
    batch, plan = M.make_batch([_cuda(entry)], "cuda", "sgdet")
    out, _ = M.sttran_forward(k, P, batch, plan, "sgdet", False, False)
    # rows of the slice that are NOT in its first frame see exactly the same context as in the full video
    later = (sub["im_idx"] > 0)
    rows_full = torch.nonzero(keep_p).flatten()[later]
    got = {"attention_distribution": out["logits26"][:, :3], "spatial_distribution": torch.sigmoid(out["logits26"][:, 3:9]),
           "contacting_distribution": torch.sigmoid(out["logits26"][:, 9:])}
    for key, g in got.items():
        assert G.rel_err(g[rows_full.cuda()].float().cpu(), want[key][later]) < tol, key
    # object classifier rows are per box: all boxes of the slice
    keep_b = (entry["boxes"][:, 0] >= f0) & (entry["boxes"][:, 0] <= f1)
    assert G.rel_err(out["distribution"][keep_b.cuda()].float().cpu(), want["distribution"]) < tol


def test_c4_training_step_is_finite_and_deterministic(cuda_lib):
    """Forward + backward of the long video through the fused trainer twice from the same state: identical loss, finite grads."""
    from nlvsgg_b200 import model as M, shapes
    from nlvsgg_b200.trainer import Trainer
    entry, _ = synth.synth_video(405, 200, 20, "sgdet", fixed_boxes=20, with_gt=False)
    losses = []
    for _ in range(2):
        tr = Trainer({k: v.cuda() for k, v in synth.make_state_dict(shapes.sttran_template(), 5).items()}, "sgdet", "sttran", "bf16",
                     device=torch.device("cuda"))
        batch, _ = M.make_batch([_cuda(entry)], "cuda", "sgdet")
        loss = tr.step(batch)
        losses.append(float(loss))
    assert np.isfinite(losses[0]) and abs(losses[0] - losses[1]) <= 2e-3 * abs(losses[0])   # split-K atomics reorder fp32 sums


def test_c3_dsg_batch_of_8_videos_equals_single_videos_and_oracle(cuda_lib):
    from oracle import cref, model as omodel
    from nlvsgg_b200 import engine as E, model as M
    sd = synth.make_state_dict(G.dsg_template(), 6)
    P = {k: v.cuda() for k, v in sd.items()}
    k = E.Kernels("bf16x3")
    cpu_entries = [synth.synth_video(600 + i, 20 + 2 * i, 7, "sgdet", draw_fn=cref.draw_union_boxes, with_gt=False)[0] for i in range(8)]
    entries = [_cuda(e) for e in cpu_entries]
    singles = []
    for e in entries:
        b, pl = M.make_batch([e], "cuda", "sgdet", dsg=True)
        out, _ = M.dsg_forward(k, {n: t.clone() for n, t in P.items()}, b, pl, "sgdet", False, False)
        singles.append(out)
    b, pl = M.make_batch(entries, "cuda", "sgdet", dsg=True)
    out, _ = M.dsg_forward(k, {n: t.clone() for n, t in P.items()}, b, pl, "sgdet", False, False)
    assert G.rel_err(out["logits26"].cpu(), torch.cat([s["logits26"] for s in singles]).cpu()) < 1e-5
    assert G.rel_err(out["distribution"].cpu(), torch.cat([s["distribution"] for s in singles]).cpu()) < 1e-5
    with torch.no_grad():
        want = omodel.dsg_forward(sd, cpu_entries[3], "sgdet", training=False)
    got = singles[3]["logits26"].float().cpu()
    assert G.rel_err(got[:, :3], want["attention_distribution"]) < 1e-3
    assert G.rel_err(torch.sigmoid(got[:, 3:9]), want["spatial_distribution"]) < 1e-3
    assert G.rel_err(torch.sigmoid(got[:, 9:]), want["contacting_distribution"]) < 1e-3


def _split_frame_counts(n_videos, rng):
    """Frame counts shaped like datasets/AG/ag_test_id.pkl: mean ~31, long tail up to 121."""
    c = np.clip(np.round(rng.gamma(shape=3.2, scale=9.8, size=n_videos)), 3, 121).astype(int)
    c[0] = 121
    return c


def test_c5_recall_over_a_test_split_shape(cuda_lib):
    from oracle.make_golden_eval import synth_pred
    from tests.test_cpu_evaluator import assert_same_results, make_oracle
    from tests.test_gpu_evaluator import make_cuda_eval, to_cuda
    rng = np.random.default_rng(5)
    n_videos = 1737
    counts = _split_frame_counts(n_videos, rng)
    # 40 distinct synthetic videos tiled over the split (generation is python-loop bound), every video its own frame count
    base = [synth_pred("sgdet", 5000 + i, int(counts[i]), 6, 0.05, saturate=(i % 7 == 0)) for i in range(40)]
    order = [i % 40 for i in range(n_videos)]
    items = [(base[j][1], to_cuda(base[j][0])) for j in order]
    ev = make_cuda_eval("sgdet")
    ev.evaluate_videos(items)
    ev.calculate_mean_recall()
    n_frames = sum(len(base[j][1]) for j in order)
    key = "sgdet_recall"
    assert all(len(ev.result_dict[key][k]) == n_frames for k in (10, 20, 50))
    # bit-identical to the numpy oracle on the 40 distinct videos (the first 40 items)
    oe = make_oracle("sgdet")
    sub = make_cuda_eval("sgdet")
    for j in range(40):
        pred, gt = base[j]
        dp = to_cuda(pred)
        sub.evaluate_scene_graph(gt, dp)
        oe.evaluate_scene_graph(gt, {kk: (v.cpu() if torch.is_tensor(v) else v) for kk, v in dp.items()})
    sub.calculate_mean_recall(); oe.calculate_mean_recall()
    assert_same_results(sub.result_dict, oe.result_dict, "sgdet")
    first = sum(len(base[j][1]) for j in range(40))
    for name in ("sgdet_recall", "sgdet_recall_nogc", "sgdet_semi_recall"):
        for k in (10, 20, 50):
            assert ev.result_dict[name][k][:first] == sub.result_dict[name][k], (name, k)
            # the split is a tiling of those 40 videos: every later copy repeats the same per-frame values
            period = ev.result_dict[name][k][:first]
            full = ev.result_dict[name][k]
            assert full[first:2 * first] == period and full[43 * first:] == period[:len(full) - 43 * first]
    # order invariance: reversing the videos reverses the per-video blocks, the dataset mean is the same float
    ev2 = make_cuda_eval("sgdet")
    ev2.evaluate_videos([(base[j][1], to_cuda(base[j][0])) for j in reversed(order)])
    for k in (10, 20, 50):
        assert np.mean(sorted(ev2.result_dict[key][k])) == np.mean(sorted(ev.result_dict[key][k]))


def test_packed_bf16_feature_format_is_bit_identical_in_bf16_mode(cuda_lib):
    """a loader that stores `features` / `union_feat` as bf16 halves the host -> device bytes; in the bf16
    compute mode the device rounds those tensors to bf16 first anyway, so logits and loss must not change by one bit."""
    from nlvsgg_b200 import engine as E, model as M
    entries = [synth.synth_video(700 + i, 6 + i, 5, "sgdet", with_gt=False)[0] for i in range(3)]
    sd = synth.make_state_dict(G.sttran_template(), 7)
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        P = {k: v.cuda() for k, v in sd.items()}
        hb = M.collate(entries, "sgdet", pin=True, feat_dtype=dt)
        assert hb.union_feat.dtype == dt and hb.features.dtype == dt
        plan = M.make_plan(hb, "cuda", "sgdet")
        batch = M.upload(hb, "cuda")
        out, _ = M.sttran_forward(E.Kernels("bf16"), P, batch, plan, "sgdet", True, False)
        outs.append(out)
    assert M.input_bytes(M.collate(entries, "sgdet", feat_dtype=torch.bfloat16)) < 0.51 * M.input_bytes(M.collate(entries, "sgdet"))
    assert torch.equal(outs[0]["logits26"], outs[1]["logits26"])
    assert torch.equal(outs[0]["distribution"], outs[1]["distribution"])
