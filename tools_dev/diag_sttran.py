"""Dev diagnostic: per-output and per-gradient relative errors vs the golden fixtures, per precision."""
import sys, os, copy, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import golden_util as G
from tests.test_gpu_sttran import _build, _entry_cuda, _reference_style_loss
from oracle import cref

res = {}
for name in G.model_cases("sttran_"):
    case = G.load_case(name)
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    for prec in ("fp32", "bf16x3", "bf16"):
        m = _build(case, prec, case["training"])
        if case["training"]:
            pred = m(_entry_cuda(entry)); loss = _reference_style_loss(pred); loss.backward()
        else:
            with torch.no_grad(): pred = m(_entry_cuda(entry))
        r = {k: G.rel_err(pred[k].detach().cpu(), w) for k, w in case["outputs"].items()}
        if case["training"]:
            r["loss"] = abs(loss.item() - case["loss"]) / abs(case["loss"])
            ge = {}
            for n, p in m.named_parameters():
                dg = case["grads"][n]; g = p.grad.detach().double().flatten().cpu()
                if "full" in dg:
                    ge[n] = (g - dg["full"].double()).abs().max().item() / (dg["full"].double().abs().max().item() + 1e-30)
                else:
                    rms = (dg["sq_sum"] / g.numel()) ** 0.5
                    ge[n] = max((g[:64] - dg["head"].double()).abs().max().item() / (dg["head"].double().abs().max().item() + 1e-30),
                                abs(g.abs().sum().item() - dg["abs_sum"]) / (dg["abs_sum"] + 1e-30))
            top = sorted(ge.items(), key=lambda kv: -kv[1])[:8]
            r["grad_top"] = top
        res[f"{name}/{prec}"] = r
        print(name, prec, json.dumps(r, default=str)[:1500], flush=True)
