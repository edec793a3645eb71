"""STTran forward / backward through the drop-in module on CUDA vs (a) golden vectors produced by the reference
itself and (b) the CPU oracle, on identical seeded inputs and weights.

Tolerances (relative to the max |reference| of each tensor):
  fp32 / bf16x3 modes : 1e-3  (north-star bound for fp32-accumulated logits)
  bf16 mode           : 6e-2  (bf16 operand rounding through 7 transformer layers; reported, not a parity claim)
"""
import copy

import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-3, "bf16x3": 1e-3, "bf16": 6e-2}
OUT_KEYS = ("attention_distribution", "spatial_distribution", "contacting_distribution", "distribution")


def _build(case, precision, training):
    from nlvsgg_b200.lib.sttran import STTran
    m = STTran(case["mode"], 3, 6, 17, synth.AG_OBJECT_CLASSES, 1, 3, "wk", True, 2048, precision=precision)
    sd = synth.make_state_dict(G.sttran_template(), case["seed"])
    m.load_state_dict(sd)          # same names/shapes as the reference: drop-in checkpoint compatibility
    m = m.cuda()
    m.train(training)
    m.kernels.dropout = 0.0      # parity runs with the dropout probability forced to 0 (masks cannot be RNG-matched)
    return m


def _entry_cuda(entry):
    return {k: (v.cuda() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in entry.items()}


def _reference_style_loss(pred):
    """tools/train_STTran.py:143-189 (bce_loss=True) with torch ops, as the unchanged training script would run it."""
    import torch.nn as nn
    ce, bce = nn.CrossEntropyLoss(), nn.BCELoss()
    dev = pred["attention_distribution"].device
    att_mask = torch.tensor([len(i) > 0 for i in pred["attention_gt"]], device=dev)
    att_label = torch.tensor([int(i[0]) for i in pred["attention_gt"] if len(i) >= 1], dtype=torch.int64, device=dev)
    R = len(pred["spatial_gt"])
    spa = torch.zeros(R, 6, device=dev)
    con = torch.zeros(R, 17, device=dev)
    for i in range(R):
        spa[i, pred["spatial_gt"][i]] = 1.0
        con[i, pred["contacting_gt"][i]] = 1.0
    loss = ce(pred["distribution"], pred["labels"]) + ce(pred["attention_distribution"][att_mask], att_label)
    sm, cm = (spa > 0).sum(-1) != 0, (con > 0).sum(-1) != 0
    loss = loss + bce(pred["spatial_distribution"][sm], spa[sm]) + bce(pred["contacting_distribution"][cm], con[cm])
    return loss


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("name", [n for n in G.model_cases("sttran_") if "train" not in n])
def test_sttran_eval_matches_reference(cuda_lib, name, precision):
    from oracle import cref
    case = G.load_case(name)
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    m = _build(case, precision, False)
    with torch.no_grad():
        pred = m(_entry_cuda(entry))
    for k, want in case["outputs"].items():
        err = G.rel_err(pred[k].cpu(), want)
        assert err < TOL[precision], f"{k}: rel err {err:.3e}"
    assert pred["pred_labels"].cpu().equal(entry["labels"])


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_additive_int_mask_mode_matches_reference(cuda_lib, precision):
    """kernels.additive_mask: the int key_padding_mask of lib/transformer_wk.py:154 as torch 1.10.1 read it (+1 on the logits
    of the padded keys instead of masking them) — golden written by the reference under that reading; inference only."""
    from oracle import cref
    case = G.load_case("additive_sttran_eval")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    m = _build(case, precision, False)
    m.kernels.additive_mask = True
    with torch.no_grad():
        pred = m(_entry_cuda(entry))
    for k, want in case["outputs"].items():
        assert G.rel_err(pred[k].cpu(), want) < TOL[precision], f"{k}: rel err {G.rel_err(pred[k].cpu(), want):.3e}"
    m.kernels.additive_mask = False
    with torch.no_grad():
        plain = m(_entry_cuda(entry))
    assert G.rel_err(plain["attention_distribution"].cpu(), case["outputs"]["attention_distribution"]) > 1e-3
    m.kernels.additive_mask = True
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(_entry_cuda(entry))


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("name", [n for n in G.model_cases("sttran_") if "train" in n])
def test_sttran_train_step_matches_reference(cuda_lib, name, precision):
    from oracle import cref
    case = G.load_case(name)
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    m = _build(case, precision, True)
    pred = m(_entry_cuda(entry))
    loss = _reference_style_loss(pred)
    loss.backward()
    tol = TOL[precision]
    assert abs(loss.item() - case["loss"]) <= tol * abs(case["loss"]), f"loss {loss.item()} vs {case['loss']}"
    for k, want in case["outputs"].items():
        assert G.rel_err(pred[k].detach().cpu(), want) < tol, k
    sd = m.state_dict()
    for k, want in case["running"].items():
        assert G.rel_err(sd[k].cpu(), want) < tol, k
    # gradient check: relative L2 error per tensor (digest: first 64 entries + |g| sum for the large ones).
    # bf16x3 is looser than its forward error because ReLU masks of near-zero activations flip on these tiny
    # (27-pair) batches; structurally-zero gradients (bias in front of a BatchNorm) are checked absolutely.
    # bf16 is the throughput mode, not a parity claim: a 20-pair batch gives single weight slices whose bf16 rounding
    # noise reaches ~0.35 rel-L2 (run-to-run with split-K atomics); the parity bars are the fp32 / bf16x3 rows.
    gtol = {"fp32": 2e-3, "bf16x3": 3e-2, "bf16": 1.0}[precision]
    bad, errs = [], []
    for n, p in m.named_parameters():
        assert p.grad is not None, f"{n} received no gradient"
        dg = case["grads"][n]
        g = p.grad.detach().double().flatten().cpu()
        ref = dg["full"].double() if "full" in dg else dg["head"].double()
        got = g if "full" in dg else g[:64]
        if ref.abs().max().item() < 1e-6:
            err = 0.0 if got.abs().max().item() < 1e-4 else float("inf")
        else:
            err = (got - ref).norm().item() / (ref.norm().item() + 1e-30)
            if "full" not in dg:
                err = max(err, abs(g.abs().sum().item() - dg["abs_sum"]) / (dg["abs_sum"] + 1e-30))
        errs.append(err)
        if err > gtol:
            bad.append((n, err))
    assert not bad, f"gradient mismatches (rel L2): {bad[:12]}"
    if precision == "bf16":
        # the throughput mode is not a parity claim, but its gradients must still be the right ones: a single 64-entry slice of a
        # weight gradient can carry ~0.5 of bf16 rounding noise on a 20-pair batch, the typical tensor must not
        errs.sort()
        assert errs[len(errs) // 2] < 0.1, f"median gradient error {errs[len(errs) // 2]:.3f}"


def test_sttran_matches_oracle_intermediates(cuda_lib):
    """fp32 mode vs the CPU oracle on a fresh seed with empty frames: pair tokens, transformer output, logits."""
    from oracle import cref, model as omodel
    from nlvsgg_b200 import engine as E, model as M
    seed = 21
    entry, _ = synth.synth_video(seed, 10, 5, "sgdet", draw_fn=cref.draw_union_boxes, empty_frame_prob=0.25)
    sd = synth.make_state_dict(G.sttran_template(), seed)
    with torch.no_grad():
        want = omodel.sttran_forward(sd, entry, "sgdet", training=False, return_tokens=True)
    P = {k: v.cuda() for k, v in sd.items()}
    k = E.Kernels("fp32")
    batch, plan = M.make_batch([_entry_cuda(entry)], "cuda", "sgdet")
    out, _ = M.sttran_forward(k, P, batch, plan, "sgdet", False, False)
    assert G.rel_err(out["distribution"].cpu(), want["distribution"]) < 1e-4
    assert G.rel_err(out["rel_tokens"].cpu(), want["rel_features"]) < 1e-4
    assert G.rel_err(out["rel_out"].cpu(), want["global_output"]) < 1e-4


def test_batched_videos_equal_single_videos(cuda_lib):
    """A 3-video batch (per-video BatchNorm statistics, per-video windows) reproduces the 3 single-video results."""
    from oracle import cref
    from nlvsgg_b200 import engine as E, model as M
    sd = synth.make_state_dict(G.sttran_template(), 3)
    P = {k: v.cuda() for k, v in sd.items()}
    k = E.Kernels("fp32")
    entries = [_entry_cuda(synth.synth_video(s, f, 5, "sgdet", draw_fn=cref.draw_union_boxes, empty_frame_prob=p)[0])
               for s, f, p in ((31, 6, 0.0), (32, 1, 0.0), (33, 7, 0.3))]
    singles = []
    for e in entries:
        Pi = {n: t.clone() for n, t in P.items()}
        b, pl = M.make_batch([e], "cuda", "sgdet")
        out, _ = M.sttran_forward(k, Pi, b, pl, "sgdet", True, False)
        singles.append(out)
    Pb = {n: t.clone() for n, t in P.items()}
    b, pl = M.make_batch(entries, "cuda", "sgdet")
    out, _ = M.sttran_forward(k, Pb, b, pl, "sgdet", True, False)
    want26 = torch.cat([s["logits26"] for s in singles])
    wantobj = torch.cat([s["distribution"] for s in singles])
    assert G.rel_err(out["logits26"].cpu(), want26.cpu()) < 1e-5
    assert G.rel_err(out["distribution"].cpu(), wantobj.cpu()) < 1e-5


class _RefStyleAdamW:
    """The update rule of the reference optimizer (lib/AdamW.py:64-112), restated: decoupled decay and the Adam step are
    applied through `p.data` — invisible to autograd version counters, which is what a weight cache must survive."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params, self.lr, self.betas, self.eps, self.wd = list(params), lr, betas, eps, weight_decay
        self.state = {}

    def step(self):
        b1, b2 = self.betas
        for i, p in enumerate(self.params):
            if p.grad is None:
                continue
            p.data.mul_(1 - self.lr * self.wd)
            st = self.state.setdefault(i, {"step": 0, "m": torch.zeros_like(p.data), "v": torch.zeros_like(p.data)})
            st["step"] += 1
            st["m"].mul_(b1).add_(p.grad.data, alpha=1 - b1)
            st["v"].mul_(b2).addcmul_(p.grad.data, p.grad.data, value=1 - b2)
            denom = st["v"].sqrt().add_(self.eps)
            step_size = self.lr * (1 - b2 ** st["step"]) ** 0.5 / (1 - b1 ** st["step"])
            p.data.add_(st["m"].div(denom).mul_(-step_size))

    def zero_grad(self):
        for p in self.params:
            p.grad = None


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dropin_training_loop_with_reference_style_optimizer(cuda_lib, precision):
    """model(entry); loss.backward(); optimizer.step() twice, the optimizer writing through p.data like lib/AdamW.py:
    the second forward must see the updated weights (fp32: step-2 outputs equal the CPU oracle put through the same loop)."""
    from oracle import cref, model as omodel
    case = G.load_case("sttran_sgdet_train")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    m = _build(case, precision, True)
    opt = _RefStyleAdamW(m.parameters(), lr=2e-3)
    outs = []
    for _ in range(2):
        opt.zero_grad()
        pred = m(_entry_cuda(entry))
        outs.append({k: pred[k].detach().cpu().clone() for k in OUT_KEYS})
        _reference_style_loss(pred).backward()
        opt.step()
    # the same two iterations on the CPU oracle
    sd = synth.make_state_dict(G.sttran_template(), case["seed"])
    names = [n for n, _ in m.named_parameters()]
    for n in names:
        sd[n].requires_grad_(True)
    oopt = _RefStyleAdamW([sd[n] for n in names], lr=2e-3)
    wants = []
    for _ in range(2):
        oopt.zero_grad()
        pred = omodel.sttran_forward(sd, entry, "sgdet", training=True)
        wants.append({k: pred[k].detach().clone() for k in OUT_KEYS})
        omodel.training_loss(pred, entry, "sgdet").backward()
        oopt.step()
    moved = G.rel_err(wants[1]["attention_distribution"], wants[0]["attention_distribution"])
    assert moved > 5e-2, moved                                         # the step is large enough to matter
    assert G.rel_err(outs[1]["attention_distribution"], outs[0]["attention_distribution"]) > 0.5 * moved   # weights really moved
    if precision == "fp32":
        for k in OUT_KEYS:
            assert G.rel_err(outs[1][k], wants[1][k]) < 2e-3, k
