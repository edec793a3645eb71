"""Packed feature files on the device: the unpack / create_dis kernels and bit-identity of the model fed from them."""
import ctypes

import numpy as np
import pytest
import torch

from nlvsgg_b200 import featfile as FF, model as M, synth
from tests import golden_util as G
from tests.test_cpu_featfile import _entries, _unpack_numpy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pack12", [True, False])
def test_union_unpack_kernel_matches_numpy(cuda_lib, tmp_path, pack12):
    from nlvsgg_b200 import _C
    entries = _entries()
    if pack12:                                   # values outside their rows' windows: exceptions patched after the decode
        entries[0]["union_feat"][0, 5, 3, 3] = -1.5
        entries[1]["union_feat"][2, 7, 0, 0] = 1e-30
    hb = FF.Loader(pin=False).load(FF.write_videos(str(tmp_path), entries, pack12=pack12))
    want = _unpack_numpy(hb)
    ref = M.collate(entries, "sgdet").union_feat.bfloat16().permute(0, 2, 3, 1).reshape(-1, 2048).contiguous()
    assert np.array_equal(want, ref.view(torch.int16).numpy().view(np.uint16))
    bm, off, vals = hb.union_bitmap.cuda(), hb.union_off.cuda(), hb.union_feat.cuda()
    out = torch.empty(want.shape, dtype=torch.bfloat16, device="cuda")
    vp = ctypes.c_void_p
    if pack12:
        assert hb.union_rows == 3
        hx, base = hb.union_hx.cuda(), hb.union_base.cuda()
        _C.check(_C.lib().nlv_union_unpack12(vp(bm.data_ptr()), vp(off.data_ptr()), vp(vals.data_ptr()), vp(hx.data_ptr()), vp(base.data_ptr()),
                                             ctypes.c_longlong(want.shape[0]), vp(out.data_ptr()), None), "union_unpack12")
        ep, ev = hb.union_exc_pos.cuda(), hb.union_exc_val.cuda()
        assert ep.numel() == 2
        _C.check(_C.lib().nlv_union_patch(vp(out.data_ptr()), vp(ep.data_ptr()), vp(ev.data_ptr()), ep.numel(), None), "union_patch")
    else:
        _C.check(_C.lib().nlv_union_unpack(vp(bm.data_ptr()), vp(off.data_ptr()), vp(vals.data_ptr()),
                                           ctypes.c_longlong(want.shape[0]), vp(out.data_ptr()), None), "union_unpack")
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().view(torch.int16).numpy().view(np.uint16), want)


def test_create_dis_kernel_is_bit_exact(cuda_lib):
    from nlvsgg_b200 import _C
    g = torch.Generator().manual_seed(3)
    conf = torch.cat((torch.rand(500, generator=g), torch.tensor([0.0, 1.0, 0.2, 1.0 / 36])))
    idx = torch.randint(0, 36, (conf.numel(),), generator=g, dtype=torch.int32)
    want = torch.stack([synth._create_dis(float(c), int(i)) for c, i in zip(conf, idx)])       # python-double (1 - conf) / 35
    other = want.gather(1, ((idx.long() + 1) % 36)[:, None])[:, 0].contiguous()
    want = torch.where(want == 0, other[:, None].expand(-1, 36), want)     # the reference's `d[d == 0] = (1 - conf) / 35` (conf == 0 rows)
    out = torch.empty(conf.numel(), 36, device="cuda")
    c, o, i = conf.cuda(), other.cuda(), idx.cuda()
    _C.check(_C.lib().nlv_create_dis(ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(o.data_ptr()), ctypes.c_void_p(i.data_ptr()),
                                     ctypes.c_longlong(conf.numel()), ctypes.c_void_p(out.data_ptr()), None), "create_dis")
    assert torch.equal(out.cpu(), want)
    # without `other`: the reference's own tensor arithmetic, d[d == 0] = (1 - conf) / 35 in fp32 (assign_pseudo_label.py:934-938)
    ref = torch.zeros(conf.numel(), 36)
    ref[torch.arange(conf.numel()), idx.long()] = conf
    ref = torch.where(ref == 0, ((1 - conf) / 35)[:, None].expand(-1, 36), ref)
    _C.check(_C.lib().nlv_create_dis(ctypes.c_void_p(c.data_ptr()), None, ctypes.c_void_p(i.data_ptr()),
                                     ctypes.c_longlong(conf.numel()), ctypes.c_void_p(out.data_ptr()), None), "create_dis")
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("sparse", [True, False])
def test_model_fed_from_packed_files_is_bit_identical_in_bf16_mode(cuda_lib, tmp_path, sparse):
    """bf16 storage + channels-last + zero suppression change no bit of the bf16-mode results: the device rounds the fp32
    entry tensors to bf16 before their first use anyway."""
    from nlvsgg_b200 import engine as E
    entries = _entries()
    sd = synth.make_state_dict(G.sttran_template(), 7)
    hb_ref = M.collate(entries, "sgdet")
    hb_pk = FF.Loader(pin=False).load(FF.write_videos(str(tmp_path), entries, sparse=sparse))
    outs = []
    for hb in (hb_ref, hb_pk):
        P = {k: v.cuda() for k, v in sd.items()}
        plan = M.make_plan(hb, "cuda", "sgdet")
        batch = M.upload(hb, "cuda", rasterise=False)
        out, _ = M.sttran_forward(E.Kernels("bf16"), P, batch, plan, "sgdet", True, False)
        outs.append(out)
    assert torch.equal(outs[0]["logits26"], outs[1]["logits26"])
    assert torch.equal(outs[0]["distribution"], outs[1]["distribution"])


def test_fused_trainer_step_from_packed_files(cuda_lib, tmp_path):
    from nlvsgg_b200.trainer import Trainer
    entries = _entries()
    sd = synth.make_state_dict(G.sttran_template(), 8)
    losses = []
    for packed in (False, True):
        tr = Trainer({k: v.cuda() for k, v in sd.items()}, "sgdet", "sttran", "bf16")
        hb = FF.Loader(pin=False).load(FF.write_videos(str(tmp_path), entries)) if packed else M.collate(entries, "sgdet")
        loss = tr.step(M.upload(hb, "cuda", rasterise=False))
        losses.append(float(loss))
    assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[0])      # identical logits; the loss scalar is an atomic float sum
