"""DSG-DETR (lib/dsg_detr.py) forward / backward through the drop-in module vs golden vectors from the reference."""
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G
from tests.test_gpu_sttran import TOL, _entry_cuda, _reference_style_loss

pytestmark = pytest.mark.gpu


def _build(case, precision, training):
    from nlvsgg_b200.lib.dsg_detr import STTran
    # tools/test_DSG_DETR.py:39-50 passes six extra kwargs the reference class does not accept; the drop-in ignores them
    m = STTran(case["mode"], 3, 6, 17, synth.AG_OBJECT_CLASSES, enc_layer_num=1, dec_layer_num=3, precision=precision)
    m.load_state_dict(synth.make_state_dict(G.dsg_template(), case["seed"]))
    m = m.cuda()
    m.train(training)
    m.kernels.dropout = 0.0      # parity runs with the dropout probability forced to 0 (masks cannot be RNG-matched)
    return m


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_dsg_eval_matches_reference(cuda_lib, precision):
    from oracle import cref
    case = G.load_case("dsg_sgdet_eval")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    m = _build(case, precision, False)
    with torch.no_grad():
        pred = m(_entry_cuda(entry))
    for k, want in case["outputs"].items():
        assert G.rel_err(pred[k].cpu(), want) < TOL[precision], k


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_dsg_train_step_matches_reference(cuda_lib, precision):
    from oracle import cref
    case = G.load_case("dsg_sgdet_train")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    m = _build(case, precision, True)
    pred = m(_entry_cuda(entry))
    loss = _reference_style_loss(pred)
    loss.backward()
    tol = TOL[precision]
    assert abs(loss.item() - case["loss"]) <= tol * abs(case["loss"])
    gtol = {"fp32": 2e-3, "bf16x3": 3e-2}[precision]
    bad = []
    for n, p in m.named_parameters():
        if n in case["no_grad_params"]:
            assert p.grad is None, n          # object-track encoder: unused in sgdet, as in the reference
            continue
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
        if err > gtol:
            bad.append((n, err))
    assert not bad, bad[:10]


def test_dsg_fused_trainer_matches_golden_loss(cuda_lib):
    from oracle import cref
    from nlvsgg_b200 import model as M
    from nlvsgg_b200.trainer import Trainer
    case = G.load_case("dsg_sgdet_train")
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    sd = synth.make_state_dict(G.dsg_template(), case["seed"])
    tr = Trainer({k: v.cuda() for k, v in sd.items()}, "sgdet", "dsg", "fp32")
    loss, _ = tr.forward_backward(M.upload(M.collate([entry], "sgdet"), "cuda"))
    assert abs(loss.item() - case["loss"]) <= 1e-5 * abs(case["loss"])
    tr.optimizer_step()
