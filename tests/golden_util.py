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


from nlvsgg_b200.shapes import dsg_template, sttran_template  # noqa: E402,F401


def rel_err(got, want):
    return (got.double() - want.double()).abs().max().item() / max(want.double().abs().max().item(), 1e-12)
