"""Helpers shared by the CPU (oracle-vs-golden) and GPU (CUDA-vs-oracle/golden) tests."""
import glob
import os

import torch

from nlvsgg_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def model_cases(prefix):
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.pt")))


def load_case(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def case_inputs(case, draw_fn):
    entry, gt = synth.synth_video(case["seed"], case["frames"], case["mean_boxes"], case["mode"], draw_fn=draw_fn,
                                  empty_frame_prob=case["empty_frame_prob"])
    assert entry["boxes"].shape[0] == case["n_boxes"] and entry["pair_idx"].shape[0] == case["n_pairs"]
    return entry, gt


def sttran_template():
    """Parameter/buffer names and shapes of lib/sttran.py:STTran ([probed])."""
    t = {}
    def lin(p, o, i): t[p + ".weight"] = torch.empty(o, i); t[p + ".bias"] = torch.empty(o)
    def bn(p, c):
        for s in ("weight", "bias", "running_mean", "running_var"): t[f"{p}.{s}"] = torch.empty(c)
        t[p + ".num_batches_tracked"] = torch.empty((), dtype=torch.int64)
    def ln(p, c): t[p + ".weight"] = torch.empty(c); t[p + ".bias"] = torch.empty(c)
    def mha(p, d):
        t[p + ".in_proj_weight"] = torch.empty(3 * d, d); t[p + ".in_proj_bias"] = torch.empty(3 * d)
        lin(p + ".out_proj", d, d)
    oc = "object_classifier"
    t[oc + ".obj_embed.weight"] = torch.empty(36, 200)
    bn(oc + ".pos_embed.0", 4); lin(oc + ".pos_embed.1", 128, 4)
    lin(oc + ".decoder_lin.0", 1024, 2376); bn(oc + ".decoder_lin.1", 1024); lin(oc + ".decoder_lin.3", 37, 1024)
    t["union_func1.weight"] = torch.empty(256, 2048, 1, 1); t["union_func1.bias"] = torch.empty(256)
    t["conv.0.weight"] = torch.empty(128, 2, 7, 7); t["conv.0.bias"] = torch.empty(128); bn("conv.2", 128)
    t["conv.4.weight"] = torch.empty(256, 128, 3, 3); t["conv.4.bias"] = torch.empty(256); bn("conv.6", 256)
    lin("subj_fc", 512, 2048); lin("obj_fc", 512, 2048); lin("vr_fc", 512, 12544)
    t["obj_embed.weight"] = torch.empty(37, 200); t["obj_embed2.weight"] = torch.empty(37, 200)
    g = "glocal_transformer"
    p = g + ".local_attention.layers.0"
    mha(p + ".self_attn", 1936); lin(p + ".linear1", 2048, 1936); lin(p + ".linear2", 1936, 2048)
    ln(p + ".norm1", 1936); ln(p + ".norm2", 1936)
    for i in range(3):
        p = f"{g}.global_attention.layers.{i}"
        mha(p + ".multihead2", 1936); lin(p + ".linear1", 2048, 1936); lin(p + ".linear2", 1936, 2048)
        ln(p + ".norm3", 1936)
    t[g + ".position_embedding.weight"] = torch.empty(2, 1936)
    lin("a_rel_compress", 3, 1936); lin("s_rel_compress", 6, 1936); lin("c_rel_compress", 17, 1936)
    return t


def rel_err(got, want):
    return (got.double() - want.double()).abs().max().item() / max(want.double().abs().max().item(), 1e-12)
