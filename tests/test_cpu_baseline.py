"""The CPU arm of bench.py (oracle/baseline_batched.py): same numbers as the segment-wise oracle, and — in the build
container, where the reference can be imported — the same speed as the reference's own forward + backward."""
import copy
import time

import pytest
import torch

from nlvsgg_b200 import shapes, synth
from oracle import baseline_batched as BB, cref, model as omodel, ref_harness as H


@pytest.mark.parametrize("seed,frames,empty", [(41, 6, 0.0), (42, 9, 0.3), (43, 1, 0.0)])
def test_padded_formulation_equals_segment_oracle(seed, frames, empty):
    entry, _ = synth.synth_video(seed, frames, 5, "sgdet", draw_fn=cref.draw_union_boxes, empty_frame_prob=empty)
    sd = synth.make_state_dict(shapes.sttran_template(), seed)
    with torch.no_grad():
        a = BB.sttran_forward_padded({k: v.clone() for k, v in sd.items()}, entry, "sgdet", training=True)
        b = omodel.sttran_forward({k: v.clone() for k, v in sd.items()}, entry, "sgdet", training=True)
    for k in ("attention_distribution", "spatial_distribution", "contacting_distribution", "distribution"):
        assert (a[k] - b[k]).abs().max().item() <= 2e-5 * max(b[k].abs().max().item(), 1.0), k


@pytest.mark.skipif(not H.available(), reason="needs /root/reference (build container)")
def test_cpu_arm_is_as_fast_as_the_reference_itself():
    """the port must not understate the reference.  Same 30-frame video, forward + backward, same thread
    count: the restated step must take at most 1.3x the reference's own lib/sttran.py."""
    ref = H.load_reference()
    torch.set_num_threads(4)
    entry, _ = synth.synth_video(5, 30, 7, "sgdet", draw_fn=cref.draw_union_boxes)
    m = H.build_reference_sttran(ref, "sgdet")
    sd = synth.make_state_dict(m.state_dict(), 5)
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    m.train()

    def ref_step():
        e = {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in entry.items()}
        pred = m(e)
        loss = omodel.training_loss(pred, entry, "sgdet")
        m.zero_grad()
        loss.backward()

    sd2 = {k: v.clone() for k, v in sd.items()}
    params = [v for k, v in sd2.items() if v.is_floating_point() and "running_" not in k]

    def port_step():
        for p in params:
            p.requires_grad_(True)
            p.grad = None
        pred = BB.sttran_forward_padded(sd2, entry, "sgdet", training=True)
        omodel.training_loss(pred, entry, "sgdet").backward()

    def best(fn, n=3):
        fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)

    t_ref, t_port = best(ref_step), best(port_step)
    assert t_port <= 1.3 * t_ref, f"port {t_port:.3f}s vs reference {t_ref:.3f}s"
