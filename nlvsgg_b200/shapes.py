"""Parameter / buffer names and shapes of the reference models (the checkpoint-compatibility surface,
SURVEY.md §8b): lib/sttran.py:STTran and lib/dsg_detr.py:STTran."""
import math

import torch


def _lin(t, p, o, i):
    t[p + ".weight"] = torch.empty(o, i)
    t[p + ".bias"] = torch.empty(o)


def _bn(t, p, c):
    for s in ("weight", "bias", "running_mean", "running_var"):
        t[f"{p}.{s}"] = torch.empty(c)
    t[p + ".num_batches_tracked"] = torch.empty((), dtype=torch.int64)


def _ln(t, p, c):
    t[p + ".weight"] = torch.empty(c)
    t[p + ".bias"] = torch.empty(c)


def _mha(t, p, d):
    t[p + ".in_proj_weight"] = torch.empty(3 * d, d)
    t[p + ".in_proj_bias"] = torch.empty(3 * d)
    _lin(t, p + ".out_proj", d, d)


def _common(t):
    oc = "object_classifier"
    t[oc + ".obj_embed.weight"] = torch.empty(36, 200)
    _bn(t, oc + ".pos_embed.0", 4)
    _lin(t, oc + ".pos_embed.1", 128, 4)
    _lin(t, oc + ".decoder_lin.0", 1024, 2376)
    _bn(t, oc + ".decoder_lin.1", 1024)
    _lin(t, oc + ".decoder_lin.3", 37, 1024)
    t["union_func1.weight"] = torch.empty(256, 2048, 1, 1)
    t["union_func1.bias"] = torch.empty(256)
    t["conv.0.weight"] = torch.empty(128, 2, 7, 7)
    t["conv.0.bias"] = torch.empty(128)
    _bn(t, "conv.2", 128)
    t["conv.4.weight"] = torch.empty(256, 128, 3, 3)
    t["conv.4.bias"] = torch.empty(256)
    _bn(t, "conv.6", 256)
    _lin(t, "subj_fc", 512, 2048)
    _lin(t, "obj_fc", 512, 2048)
    _lin(t, "vr_fc", 512, 12544)
    t["obj_embed.weight"] = torch.empty(37, 200)
    t["obj_embed2.weight"] = torch.empty(37, 200)
    _lin(t, "a_rel_compress", 3, 1936)
    _lin(t, "s_rel_compress", 6, 1936)
    _lin(t, "c_rel_compress", 17, 1936)


def sttran_template(enc_layers: int = 1, dec_layers: int = 3):
    t = {}
    _common(t)
    g = "glocal_transformer"
    for i in range(enc_layers):
        p = f"{g}.local_attention.layers.{i}"
        _mha(t, p + ".self_attn", 1936)
        _lin(t, p + ".linear1", 2048, 1936)
        _lin(t, p + ".linear2", 1936, 2048)
        _ln(t, p + ".norm1", 1936)
        _ln(t, p + ".norm2", 1936)
    for i in range(dec_layers):
        p = f"{g}.global_attention.layers.{i}"
        _mha(t, p + ".multihead2", 1936)
        _lin(t, p + ".linear1", 2048, 1936)
        _lin(t, p + ".linear2", 1936, 2048)
        _ln(t, p + ".norm3", 1936)
    t[g + ".position_embedding.weight"] = torch.empty(2, 1936)
    return t


def sinusoidal_pe(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding buffer of lib/dsg_detr.py:31-36, shape [1, max_len, d_model]."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def dsg_template():
    t = {}
    _common(t)

    def enc(p, d, ff):
        _mha(t, p + ".self_attn", d)
        _lin(t, p + ".linear1", ff, d)
        _lin(t, p + ".linear2", d, ff)
        _ln(t, p + ".norm1", d)
        _ln(t, p + ".norm2", d)
    for i in range(3):
        enc(f"object_classifier.encoder_tran.layers.{i}", 2376, 1024)
        enc(f"global_transformer.layers.{i}", 1936, 2048)
    enc("local_transformer.layers.0", 1936, 2048)
    t["object_classifier.positional_encoder.pe"] = sinusoidal_pe(600, 2376)
    t["positional_encoder.pe"] = sinusoidal_pe(400, 1936)
    return t
