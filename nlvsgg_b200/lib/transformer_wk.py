"""Drop-in for lib/transformer_wk.py / lib/transformer.py: the spatio-temporal transformer as a standalone module.
``forward(features, im_idx) -> (output, None, None)``: the attention-weight tensors the reference also returns are
never consumed (lib/sttran.py:401 discards them), so they are not materialised."""
import copy
import os

import torch
import torch.nn as nn

from .. import engine as E


class TransformerEncoderLayer(nn.Module):     # parameter container of lib/transformer.py:5-18
    def __init__(self, embed_dim=1936, nhead=4, dim_feedforward=2048, dropout=0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(embed_dim, nhead, dropout=dropout)
        self.linear1 = nn.Linear(embed_dim, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, embed_dim)
        self.norm1, self.norm2 = nn.LayerNorm(embed_dim), nn.LayerNorm(embed_dim)
        self.dropout1, self.dropout2 = nn.Dropout(dropout), nn.Dropout(dropout)


class TransformerDecoderLayer(nn.Module):     # lib/transformer.py:33-47
    def __init__(self, embed_dim=1936, nhead=4, dim_feedforward=2048, dropout=0.1):
        super().__init__()
        self.multihead2 = nn.MultiheadAttention(embed_dim, nhead, dropout=dropout)
        self.linear1 = nn.Linear(embed_dim, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, embed_dim)
        self.norm3 = nn.LayerNorm(embed_dim)
        self.dropout2, self.dropout3 = nn.Dropout(dropout), nn.Dropout(dropout)


class _Stack(nn.Module):
    def __init__(self, layer, n):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])
        self.num_layers = n


class _TransformerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernels, plan, P, names, want, features, *params):
        Pg = {"glocal_transformer." + n: t for n, t in P.items()}
        out, saved = E.sttran_transformer_fwd(kernels, Pg, plan, features.contiguous().float(), want)
        ctx.k, ctx.plan, ctx.Pg, ctx.saved, ctx.names = kernels, plan, Pg, saved, names
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = {}
        dx = E.sttran_transformer_bwd(ctx.k, ctx.Pg, ctx.plan, ctx.saved, dout.contiguous(), grads)
        ctx.saved = None
        return (None, None, None, None, None, dx) + tuple(grads.get("glocal_transformer." + n) for n in ctx.names)


class transformer_wk(nn.Module):
    """lib/transformer_wk.py:104-217 (mode 'latter'; frames / windows without pairs are skipped; a single-frame video
    returns the spatial-encoder output)."""

    def __init__(self, enc_layer_num=1, dec_layer_num=3, embed_dim=1936, nhead=8, dim_feedforward=2048, dropout=0.1,
                 mode=None, precision=None):
        super().__init__()
        assert embed_dim == 1936 and nhead == 8 and dim_feedforward == 2048, "kernels are specialised for d=1936, 8 heads"
        if mode not in (None, "latter"):
            raise NotImplementedError("only mode='latter' (the one lib/sttran.py:359 uses) is built")
        self.mode = mode
        self.local_attention = _Stack(TransformerEncoderLayer(embed_dim, nhead, dim_feedforward, dropout), enc_layer_num)
        self.global_attention = _Stack(TransformerDecoderLayer(embed_dim, nhead, dim_feedforward, dropout), dec_layer_num)
        self.position_embedding = nn.Embedding(2, embed_dim)
        nn.init.uniform_(self.position_embedding.weight)
        self._precision = precision
        self._kernels = None

    def forward(self, features, im_idx):
        if not features.is_cuda:
            raise RuntimeError("nlvsgg_b200 transformer runs on CUDA only; there is no CPU fallback")
        if self._kernels is None:
            self._kernels = E.Kernels(self._precision or os.environ.get("NLV_PRECISION", "bf16"))
        fid = im_idx.detach().cpu().numpy()
        plan = E.Plan([0], [fid], features.device)
        P = dict(self.named_parameters())
        names = list(P.keys())
        want = torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in P.values()))
        out = _TransformerFn.apply(self._kernels, plan, P, names, want, features, *P.values())
        return out, None, None


transformer = transformer_wk
