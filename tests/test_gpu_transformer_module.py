"""lib/transformer(_wk).py drop-in as a standalone module: forward(features, im_idx) and its gradients vs the oracle."""
import pytest
import torch

from nlvsgg_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("im_dtype", [torch.int64, torch.float32])   # sgdet producers give int64, predcls float32
def test_transformer_wk_module_matches_oracle(cuda_lib, im_dtype):
    from nlvsgg_b200.lib.transformer import transformer
    from oracle import model as omodel
    sd_full = synth.make_state_dict(G.sttran_template(), 11)
    sd = {k[len("glocal_transformer."):]: v for k, v in sd_full.items() if k.startswith("glocal_transformer.")}
    m = transformer(enc_layer_num=1, dec_layer_num=3, embed_dim=1936, nhead=8, dim_feedforward=2048, dropout=0.1, mode="latter",
                    precision="fp32")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(4)
    im_idx = torch.tensor([0, 0, 0, 2, 2, 3, 3, 3, 3, 6, 7, 7], dtype=im_dtype)          # frames 1, 4, 5 have no pairs
    x = torch.randn(len(im_idx), 1936, generator=g)
    xr = x.clone().requires_grad_(True)
    want = omodel.glocal_transformer(xr, im_idx, {k: v for k, v in sd_full.items()})
    want.square().sum().backward()
    xc = x.cuda().requires_grad_(True)
    out, gw, lw = m(xc, im_idx.cuda())
    assert gw is None and lw is None                                   # 3-tuple kept; weights are never consumed (sttran.py:401)
    assert G.rel_err(out.detach().cpu(), want.detach()) < 1e-4
    out.square().sum().backward()
    assert G.rel_err(xc.grad.cpu(), xr.grad) < 1e-3
    pe = m.position_embedding.weight.grad
    assert pe is not None and pe.abs().sum().item() > 0


def test_transformer_wk_mode_both_matches_reference_golden(cuda_lib):
    """mode='both' (lib/transformer_wk.py:197-207): forward, input gradient and position-embedding gradient against a golden
    written by the reference module itself (oracle/make_golden_r2.py)."""
    from nlvsgg_b200.lib.transformer_wk import transformer_wk
    z = G.load_case("transformer_both")
    sd = synth.make_state_dict({k: v for k, v in G.sttran_template().items() if k.startswith("glocal_transformer.")}, z["seed"])
    m = transformer_wk(enc_layer_num=1, dec_layer_num=3, embed_dim=1936, nhead=8, dim_feedforward=2048, dropout=0.1, mode="both", precision="fp32")
    m.load_state_dict({k[len("glocal_transformer."):]: v for k, v in sd.items()})
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(z["seed"])
    x = torch.randn(len(z["im_idx"]), 1936, generator=g).cuda().requires_grad_(True)
    out, _, _ = m(x, z["im_idx"].cuda())
    assert G.rel_err(out.detach().cpu(), z["out"]) < 1e-4
    out.square().sum().backward()
    assert G.rel_err(x.grad.cpu(), z["dx"]) < 1e-3
    assert G.rel_err(m.position_embedding.weight.grad.cpu(), z["dpos"]) < 1e-3
    with pytest.raises(ValueError):
        transformer_wk(mode="neither")
