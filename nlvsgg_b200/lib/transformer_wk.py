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
    def forward(ctx, kernels, desc, plan, P, names, want, training, features, *params):
        x = features.detach().contiguous().float()
        out, sess = E.run_transformer_forward(kernels, desc, P, plan, x, training, want)
        ctx.desc, ctx.sess, ctx.P, ctx.names, ctx.plan = desc, sess, P, names, plan
        return out

    @staticmethod
    def backward(ctx, dout):
        dx, grads = E.run_transformer_backward(ctx.sess, ctx.desc, dout.contiguous().float(), ctx.P)
        ctx.sess = None
        if ctx.plan.Mg == 0:   # single-frame input: the temporal decoder is never reached (lib/transformer_wk.py:187-188)
            grads = {n: g for n, g in grads.items() if "global_attention" not in n and "position_embedding" not in n}
        return (None, None, None, None, None, None, None, dx) + tuple(grads.get(n) for n in ctx.names)


class transformer_wk(nn.Module):
    """lib/transformer_wk.py:104-217 (modes 'latter' and 'both'; frames / windows without pairs are skipped; a single-frame
    video returns the spatial-encoder output)."""

    def __init__(self, enc_layer_num=1, dec_layer_num=3, embed_dim=1936, nhead=8, dim_feedforward=2048, dropout=0.1,
                 mode=None, precision=None):
        super().__init__()
        assert embed_dim == 1936 and nhead == 8 and dim_feedforward == 2048, "kernels are specialised for d=1936, 8 heads"
        if mode not in (None, "latter", "both"):
            raise ValueError("mode must be 'latter' or 'both' (lib/transformer_wk.py:197-215)")
        self.mode = mode
        self.local_attention = _Stack(TransformerEncoderLayer(embed_dim, nhead, dim_feedforward, dropout), enc_layer_num)
        self.global_attention = _Stack(TransformerDecoderLayer(embed_dim, nhead, dim_feedforward, dropout), dec_layer_num)
        self.position_embedding = nn.Embedding(2, embed_dim)
        nn.init.uniform_(self.position_embedding.weight)
        self._precision = precision
        self._kernels = None
        self._desc = None

    def forward(self, features, im_idx):
        if not features.is_cuda:
            raise RuntimeError("nlvsgg_b200 transformer runs on CUDA only; there is no CPU fallback")
        if self._kernels is None:
            self._kernels = E.Kernels(self._precision or os.environ.get("NLV_PRECISION", "bf16"))
            self._kernels.transformer_both = self.mode == "both"
        fid = im_idx.detach().cpu().numpy()
        plan = E.Plan([0], [fid], features.device)
        params = dict(self.named_parameters())
        P = {n: t.detach() for n, t in params.items()}
        if self._desc is None:
            self._desc = E.ModelDesc(self._kernels, P, "sttran", "predcls", transformer_prefix="")
        names = list(params.keys())
        want = torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in params.values()))
        self._kernels.seed += 1
        out = _TransformerFn.apply(self._kernels, self._desc, plan, P, names, want, self.training, features, *params.values())
        return out, None, None


transformer = transformer_wk
