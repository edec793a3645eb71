"""Fused training step (trainer.Trainer): fused losses, flat-buffer gradients, clip + AdamW kernel."""
import copy

import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _trainer(seed, precision="fp32", tmpl=None, arch="sttran"):
    from nlvsgg_b200.trainer import Trainer
    sd = synth.make_state_dict(tmpl or G.sttran_template(), seed)
    return Trainer({k: v.cuda() for k, v in sd.items()}, "sgdet", arch, precision), sd


def _batch(entries):
    from nlvsgg_b200 import model as M
    return M.upload(M.collate(entries, "sgdet"), "cuda")


@pytest.mark.parametrize("name", ["sttran_sgdet_train", "sttran_sgdet_train_gaps"])
def test_fused_step_matches_reference_golden(cuda_lib, name):
    """loss, every gradient and the BN running stats of the fused path vs the reference's own numbers."""
    from oracle import cref
    case = G.load_case(name)
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    tr, _ = _trainer(case["seed"])
    loss, _ = tr.forward_backward(_batch([entry]))
    assert abs(loss.item() - case["loss"]) <= 1e-5 * abs(case["loss"])
    assert set(tr.param_names) == set(case["grads"]), set(tr.param_names) ^ set(case["grads"])
    for n in tr.param_names:
        dg = case["grads"][n]
        g = tr.gviews[n].double().flatten().cpu()
        ref = dg["full"].double() if "full" in dg else dg["head"].double()
        got = g if "full" in dg else g[:64]
        if ref.abs().max().item() < 1e-6:
            assert got.abs().max().item() < 1e-4, n
            continue
        assert (got - ref).norm().item() <= 2e-3 * ref.norm().item(), n
        assert abs(g.abs().sum().item() - dg["abs_sum"]) <= 2e-3 * dg["abs_sum"], n
    for k, want in case["running"].items():
        assert G.rel_err(tr.P[k].cpu(), want) < 1e-4, k


def test_batch_loss_and_grads_are_the_mean_over_videos(cuda_lib):
    from oracle import cref
    entries = [synth.synth_video(s, f, 5, "sgdet", draw_fn=cref.draw_union_boxes, empty_frame_prob=p)[0]
               for s, f, p in ((61, 5, 0.0), (62, 1, 0.0), (63, 7, 0.3))]
    singles, losses = [], []
    for e in entries:
        tr, _ = _trainer(7)
        loss, _ = tr.forward_backward(_batch([e]))
        singles.append(tr.flat_g.clone()); losses.append(loss.item())
    tr, _ = _trainer(7)
    loss, _ = tr.forward_backward(_batch(entries))
    assert abs(loss.item() - sum(losses) / 3) <= 1e-5 * abs(loss.item())
    want = sum(singles) / 3
    assert (tr.flat_g - want).norm().item() <= 1e-4 * want.norm().item()


def test_clip_and_adamw_kernel_matches_reference_optimizer(cuda_lib):
    """lib/AdamW.py:52-114 after clip_grad_norm_(5), two consecutive steps, vs the CPU restatement."""
    from oracle import baseline, cref
    entry, _ = synth.synth_video(71, 4, 4, "sgdet", draw_fn=cref.draw_union_boxes)
    tr, sd = _trainer(9)
    tr.lr = 1e-3            # large enough for the update to be visible in fp32
    params = {n: sd[n].clone().requires_grad_(True) for n in tr.param_names}
    st = {}
    for _ in range(2):
        tr.forward_backward(_batch([entry]))
        for n in tr.param_names:
            params[n].grad = tr.gviews[n].detach().cpu().clone()
        baseline.adamw_update(params, st, lr=1e-3)
        tr.optimizer_step()
    for n in tr.param_names:
        a, b = tr.P[n].cpu(), params[n].detach()
        assert (a - b).abs().max().item() <= 2e-6 + 1e-5 * b.abs().max().item(), n


def test_bf16_mirror_tracks_master_weights(cuda_lib):
    from oracle import cref
    entry, _ = synth.synth_video(72, 4, 4, "sgdet", draw_fn=cref.draw_union_boxes)
    tr, _ = _trainer(9, "bf16")
    tr.lr = 1e-2
    tr.step(_batch([entry]))
    n = "glocal_transformer.global_attention.layers.1.linear1.weight"
    assert torch.equal(tr.k.mirror[n], tr.P[n].bfloat16())
    assert torch.equal(tr.k.mirror["subj_fc.weight"], tr.P["subj_fc.weight"].bfloat16())
    assert tr.steps_applied() == 1 and tr.steps_skipped() == 0


def test_non_finite_step_is_skipped_on_device(cuda_lib):
    """check_valid_iter (lib/utils.py:3-11, tools/train_STTran.py:191): a NaN loss leaves parameters and moments untouched."""
    from oracle import cref
    entry, _ = synth.synth_video(73, 4, 4, "sgdet", draw_fn=cref.draw_union_boxes)
    tr, _ = _trainer(9, "bf16")
    tr.lr = 1e-2
    tr.step(_batch([entry]))
    before = tr.flat_p.clone(), tr.flat_m.clone(), tr.flat_pb.clone()
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in entry.items()}
    bad["features"][0, 0] = float("nan")
    loss = tr.step(_batch([bad]))
    assert not torch.isfinite(loss).all()
    assert tr.steps_applied() == 1 and tr.steps_skipped() == 1
    assert torch.equal(tr.flat_p, before[0]) and torch.equal(tr.flat_m, before[1]) and torch.equal(tr.flat_pb, before[2])
    tr.step(_batch([entry]))                      # and training continues
    assert tr.steps_applied() == 2 and not torch.equal(tr.flat_p, before[0])


def test_windowless_batch_leaves_the_temporal_decoder_untouched(cuda_lib):
    """Single-frame videos never reach the temporal decoder: the reference's AdamW skips parameters whose grad is None
    (lib/AdamW.py:66) — no weight decay, no moment decay — and no gradient of an earlier batch may be re-applied."""
    from oracle import cref
    multi, _ = synth.synth_video(74, 4, 4, "sgdet", draw_fn=cref.draw_union_boxes)
    single, _ = synth.synth_video(75, 1, 5, "sgdet", draw_fn=cref.draw_union_boxes)
    tr, _ = _trainer(9, "fp32")
    tr.lr = 1e-2
    tr.step(_batch([multi]))
    dec = "glocal_transformer.global_attention.layers.2.linear1.weight"
    pos = "glocal_transformer.position_embedding.weight"
    enc = "glocal_transformer.local_attention.layers.0.linear1.weight"
    snap = {n: tr.P[n].clone() for n in (dec, pos, enc)}
    o = tr.offsets[dec]
    m_before = tr.flat_m[o:o + 16].clone()
    tr.step(_batch([single]))
    assert torch.equal(tr.P[dec], snap[dec]) and torch.equal(tr.P[pos], snap[pos])
    assert torch.equal(tr.flat_m[o:o + 16], m_before)
    assert not torch.equal(tr.P[enc], snap[enc])
    assert tr.gviews[dec].abs().max().item() == 0.0      # the flat gradient buffer is re-zeroed by every backward
