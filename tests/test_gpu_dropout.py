"""Counter-based dropout (csrc/philox.cuh): every site regenerates its mask from (seed, stream, element), so the forward and
backward kernels can be checked against torch references that are handed the SAME masks (read back with nlv_dropout_mask*).
The reference's RNG stream cannot be matched, so model-level checks are statistical."""
import numpy as np
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu
P = 0.1


def _drop(stream=5, seed=1234567890123, p=P):
    from nlvsgg_b200 import _C
    return _C.Dropout.make(p, seed, stream)


def test_mask_statistics_and_determinism(cuda_lib):
    from nlvsgg_b200 import ops
    m = ops.dropout_mask(4096, 1936, _drop()).float()
    keep = m.mean().item()
    assert abs(keep - (1 - P)) < 2e-3                                   # 7.9M draws: sigma ~ 1e-4
    assert abs(m.mean(0).min().item() - (1 - P)) < 0.03 and abs(m.mean(1).min().item() - (1 - P)) < 0.04   # no dead row / column
    assert torch.equal(m, ops.dropout_mask(4096, 1936, _drop()).float())                                   # same key -> same mask
    other = ops.dropout_mask(4096, 1936, _drop(stream=6)).float()
    assert abs((m * other).mean().item() - (1 - P) ** 2) < 3e-3                                          # sites are independent
    seeded = ops.dropout_mask(4096, 1936, _drop(seed=99)).float()
    assert abs((m * seeded).mean().item() - (1 - P) ** 2) < 3e-3
    x = torch.randn(300, 1000, device="cuda")
    d = _drop(stream=9)
    y = ops.dropout_apply(x, d)
    mk = ops.dropout_mask(300, 1000, d).float()
    assert torch.equal(y, x * mk * torch.tensor(1.0 / (1.0 - P), device="cuda", dtype=torch.float32)) or \
        (y - x * mk / (1 - P)).abs().max().item() < 1e-6


@pytest.mark.parametrize("simt", [False, True])
def test_gemm_epilogue_dropout(cuda_lib, simt):
    from nlvsgg_b200 import ops
    g = torch.Generator().manual_seed(1)
    m, n, k = 300, 1936, 512
    a = torch.randn(m, k, generator=g).cuda().bfloat16()
    b = (torch.randn(n, k, generator=g) / k ** 0.5).cuda().bfloat16()
    bias = torch.randn(n, generator=g).cuda()
    res = torch.randn(m, n, generator=g).cuda()
    d = _drop(stream=17)
    out = torch.empty(m, n, device="cuda")
    ops.gemm(a, b, out, bias=bias, residual=res, relu=True, drop=d, force_simt=simt)
    mask = ops.dropout_mask(m, n, d).float()
    want = torch.relu(a.float() @ b.float().t() + bias) * mask / (1 - P) + res
    assert (out - want).abs().max().item() <= 2e-3 * want.abs().max().item()
    # gate + gate_scale: the backward of dropout(relu(.)) through the saved activation
    h = (torch.relu(torch.randn(m, n, generator=g)).cuda() * mask).bfloat16()
    out2 = torch.empty(m, n, device="cuda")
    ops.gemm(a, b, out2, gate=h, gate_scale=1 / (1 - P), force_simt=simt)
    want2 = (a.float() @ b.float().t()) * (h.float() > 0) / (1 - P)
    assert (out2 - want2).abs().max().item() <= 2e-3 * want2.abs().max().item()


def test_layernorm_bwd_masked_operand(cuda_lib):
    from nlvsgg_b200 import ops
    g = torch.Generator().manual_seed(2)
    rows, cols = 257, 1936
    x, dy = torch.randn(rows, cols, generator=g).cuda(), torch.randn(rows, cols, generator=g).cuda()
    w, b = torch.randn(cols, generator=g).cuda(), torch.randn(cols, generator=g).cuda()
    _, _, mean, rstd = ops.layernorm_fwd(x, w, b)
    d = _drop(stream=2)
    dx, dx2, _, _ = ops.layernorm_bwd(dy, x, mean, rstd, w, dx2_dtype=torch.float32, drop=d)
    dx_plain, _, _, _ = ops.layernorm_bwd(dy, x, mean, rstd, w)
    assert torch.equal(dx, dx_plain)                                          # the residual branch stays unmasked
    mask = ops.dropout_mask(rows, cols, d).float()
    assert (dx2 - dx * mask / (1 - P)).abs().max().item() < 1e-6
    _, dx2b, _, _ = ops.layernorm_bwd(dy, x, mean, rstd, w, dx2_dtype=torch.bfloat16, drop=d)
    assert torch.equal(dx2b, (dx * mask * (1.0 / (1.0 - P))).bfloat16()) or (dx2b.float() - dx * mask / (1 - P)).abs().max().item() < 2e-2


def test_attention_weight_dropout_fwd_bwd(cuda_lib, monkeypatch):
    """nn.MultiheadAttention(dropout=p): softmax over all keys, dropped weights hit V; gradients through the same mask."""
    from nlvsgg_b200 import ops
    from nlvsgg_b200.plan import work_items
    monkeypatch.setenv("NLV_ATTN_SIMT", "0")
    hd, heads = 242, 8
    lens = [1, 3, 6, 7, 9, 12, 16, 17, 24, 33, 38]
    starts = np.concatenate(([0], np.cumsum(lens)))[:-1]
    M = int(sum(lens))
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(M, 3 * hd * heads, generator=g) * 0.5).cuda().bfloat16()
    dout = torch.randn(M, hd * heads, generator=g).cuda().bfloat16()
    work = torch.from_numpy(work_items(starts, np.asarray(lens))).cuda()
    dm = hd * heads
    q, k, v = qkv[:, :dm], qkv[:, dm:2 * dm], qkv[:, 2 * dm:]
    d = _drop(stream=11)
    o, lse = ops.attn_fwd(q, k, v, hd, heads, work, work.shape[0], torch.bfloat16, drop=d)
    mask = ops.dropout_mask_attn(M, heads, max(lens), d).float()                # [row, head, key index inside the segment]
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = torch.zeros(M, dm, device="cuda")
    for s, l in zip(starts.tolist(), lens):
        for h in range(heads):
            c = slice(h * hd, (h + 1) * hd)
            a = torch.softmax((qr[s:s + l, c] @ kr[s:s + l, c].t()) / hd ** 0.5, -1)
            a = a * mask[s:s + l, h, :l] / (1 - P)
            ref[s:s + l, c] = a @ vr[s:s + l, c]
    assert (o.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    o_plain, _ = ops.attn_fwd(q, k, v, hd, heads, work, work.shape[0], torch.bfloat16)
    assert not torch.equal(o, o_plain)
    ref.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(q, k, v, o, dout, lse, hd, heads, work, work.shape[0], dqkv[:, :dm], dqkv[:, dm:2 * dm], dqkv[:, 2 * dm:], drop=d)
    for got, want, name in ((dqkv[:, :dm], qr.grad, "dq"), (dqkv[:, dm:2 * dm], kr.grad, "dk"), (dqkv[:, 2 * dm:], vr.grad, "dv")):
        err = (got.float() - want).abs().max().item() / want.abs().max().item()
        assert err <= 0.1, (name, err)
        rel = (got.float() - want).norm().item() / want.norm().item()
        assert rel <= 2e-2, (name, rel)


def test_training_step_with_dropout(cuda_lib):
    """Model level: dropout changes the training loss, is reproducible for a seed, draws new masks every step, leaves eval
    untouched, and the parity (fp32-grade) modes refuse it loudly."""
    from oracle import cref
    from nlvsgg_b200 import model as M
    from nlvsgg_b200.trainer import Trainer
    entries = [synth.synth_video(80 + i, 6, 5, "sgdet", draw_fn=cref.draw_union_boxes)[0] for i in range(3)]
    sd = synth.make_state_dict(G.sttran_template(), 3)

    def run(p, steps=1, lr=0.0):
        tr = Trainer({k: v.cuda() for k, v in sd.items()}, "sgdet", "sttran", "bf16", dropout=p, lr=lr, weight_decay=0.0)
        return [float(tr.step(M.upload(M.collate(entries, "sgdet"), "cuda", rasterise=False))) for _ in range(steps)], tr

    base, _ = run(0.0)
    a, tr = run(P, steps=3)
    b, _ = run(P, steps=3)
    assert a == b or max(abs(x - y) for x, y in zip(a, b)) < 1e-4 * abs(a[0])      # same seed sequence -> same masks (atomics reorder sums)
    assert abs(a[0] - base[0]) > 1e-4 * abs(base[0])                              # dropout is really on
    assert len({round(x, 4) for x in a}) == 3                                     # lr = 0: only the masks change between steps
    assert all(np.isfinite(x) for x in a) and torch.isfinite(tr.flat_g).all()
    assert abs(np.mean(a) - base[0]) < 0.2 * abs(base[0])
    with pytest.raises(RuntimeError, match="dropout"):
        Trainer({k: v.cuda() for k, v in sd.items()}, "sgdet", "sttran", "fp32", dropout=P).step(
            M.upload(M.collate(entries, "sgdet"), "cuda", rasterise=False))


def test_dropin_module_uses_dropout_only_in_train_mode(cuda_lib):
    from oracle import cref
    from nlvsgg_b200.lib.sttran import STTran
    entry, _ = synth.synth_video(91, 5, 5, "sgdet", draw_fn=cref.draw_union_boxes)
    e = lambda: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in entry.items()}
    m = STTran("sgdet", 3, 6, 17, synth.AG_OBJECT_CLASSES, 1, 3, "wk", True, 2048, precision="bf16")
    m.load_state_dict(synth.make_state_dict(G.sttran_template(), 4))
    m = m.cuda()
    assert m.kernels.dropout == 0.1                                    # the reference's p (lib/transformer.py:7,36)
    m.eval()
    with torch.no_grad():
        a = m(e())["attention_distribution"].clone()
        b = m(e())["attention_distribution"].clone()
    assert torch.equal(a, b)
    m.train()
    with torch.no_grad():
        c = m(e())["attention_distribution"].clone()
        d = m(e())["attention_distribution"].clone()
    assert not torch.equal(c, d)                                       # a fresh mask per forward call
