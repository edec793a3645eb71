"""Dev: per-parameter gradient error of the fp32 drop-in vs the CPU oracle (full tensors), for golden train cases."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nlvsgg_b200 import synth
from tests import golden_util as G
from tests.test_gpu_sttran import _build, _entry_cuda, _reference_style_loss
from oracle import cref, model as omodel

for name in sys.argv[1:] or ["sttran_sgcls_train", "sttran_sgdet_train"]:
    if ":" in name:      # mode:seed:frames:k — a fresh synthetic case (oracle as the reference)
        mode, seed, frames, k = name.split(":")
        case = {"mode": mode, "seed": int(seed), "frames": int(frames), "mean_boxes": int(k), "empty_frame_prob": 0.0, "loss": float("nan")}
        e0, _ = synth.synth_video(case["seed"], case["frames"], case["mean_boxes"], mode, draw_fn=cref.draw_union_boxes)
        case["n_boxes"], case["n_pairs"] = int(e0["boxes"].shape[0]), int(e0["pair_idx"].shape[0])
    else:
        case = G.load_case(name)
    entry, _ = G.case_inputs(case, cref.draw_union_boxes)
    for prec in ("fp32",):
        m = _build(case, prec, True)
        pred = m(_entry_cuda(entry))
        loss = _reference_style_loss(pred)
        loss.backward()
        sd = synth.make_state_dict(G.sttran_template(), case["seed"])
        names = [n for n, _ in m.named_parameters()]
        for n in names:
            sd[n].requires_grad_(True)
        op = omodel.sttran_forward(sd, entry, case["mode"], training=True)
        ol = omodel.training_loss(op, entry, case["mode"])
        ol.backward()
        errs = []
        for n, p in m.named_parameters():
            g, r = p.grad.detach().double().cpu().flatten(), sd[n].grad.double().flatten()
            full = (g - r).norm().item() / (r.norm().item() + 1e-30)
            head = (g[:64] - r[:64]).norm().item() / (r[:64].norm().item() + 1e-30)
            errs.append((full, head, n))
        errs.sort(reverse=True)
        print(f"== {name} {prec}: loss {loss.item():.6f} oracle {ol.item():.6f} golden {case['loss']:.6f}; worst full-tensor rel-L2:")
        for full, head, n in (errs if os.environ.get("ALL") else errs[:8]):
            print(f"   full {full:.2e}  head64 {head:.2e}  {n}")
        print("   median full %.2e" % sorted(e[0] for e in errs)[len(errs) // 2])
        for k in ("attention_distribution", "spatial_distribution", "distribution"):
            print("   out", k, "%.2e" % G.rel_err(pred[k].detach().cpu(), op[k].detach()))
