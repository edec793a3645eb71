"""Varlen attention kernels (forward + backward) against torch autograd on ragged segments of many lengths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, segs, hd, heads):
    out = torch.zeros_like(q)
    for s, l in segs:
        for h in range(heads):
            c = slice(h * hd, (h + 1) * hd)
            a = torch.softmax((q[s:s + l, c] @ k[s:s + l, c].t()) / hd ** 0.5, -1)
            out[s:s + l, c] = a @ v[s:s + l, c]
    return out


@pytest.mark.parametrize("dtype,tol,simt", [(torch.float32, 2e-5, True), (torch.bfloat16, 2e-2, True), (torch.bfloat16, 2e-2, False)])
def test_attention_fwd_bwd_ragged(cuda_lib, monkeypatch, dtype, tol, simt):
    """simt=True: the register-resident SIMT kernels (attn.cu); simt=False: the mma.sync tensor-core kernels for bf16 I/O (attn_mma.cu)."""
    from nlvsgg_b200 import ops
    monkeypatch.setenv("NLV_ATTN_SIMT", "1" if simt else "0")
    from nlvsgg_b200.plan import work_items
    hd, heads = 242, 8
    lens = [1, 2, 3, 5, 7, 8, 9, 12, 15, 16, 17, 24, 31, 32, 33, 38, 47, 64, 65, 100]
    starts = np.concatenate(([0], np.cumsum(lens)))[:-1]
    M = int(sum(lens))
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(M, 3 * hd * heads, generator=g) * 0.5).cuda()
    dout = torch.randn(M, hd * heads, generator=g).cuda()
    work = torch.from_numpy(work_items(starts, np.asarray(lens))).cuda()
    d = hd * heads
    x = qkv.to(dtype)
    q, k, v = x[:, :d], x[:, d:2 * d], x[:, 2 * d:]
    o, lse = ops.attn_fwd(q, k, v, hd, heads, work, work.shape[0], dtype)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qr, kr, vr, list(zip(starts.tolist(), lens)), hd, heads)
    assert (o.float() - ref).abs().max().item() <= tol * ref.abs().max().item() + 1e-6
    ref.backward(dout)
    dqkv = torch.empty_like(x)
    ops.attn_bwd(q, k, v, o, dout.to(dtype), lse, hd, heads, work, work.shape[0], dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
    for got, want, name in ((dqkv[:, :d], qr.grad, "dq"), (dqkv[:, d:2 * d], kr.grad, "dk"), (dqkv[:, 2 * d:], vr.grad, "dv")):
        err = (got.float() - want).abs().max().item() / want.abs().max().item()
        assert err <= (5 * tol if dtype == torch.bfloat16 else 1e-4), (name, err)
