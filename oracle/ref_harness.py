"""TEST INFRASTRUCTURE — import harness for the *real* reference (rlqja1107/NL-VSGG).

Runs only where ``/root/reference`` exists (the build container); the GPU box never sees
it.  Nothing from the reference is copied into this repository: the reference ``lib/``
tree is copied to a scratch directory under ``/tmp`` (its Cython extensions build in
place and ``/root/reference`` is read-only), stubs are injected for the third-party
packages the reference imports but that are not installed, and the module objects are
returned to the caller.

Used by ``oracle/make_golden.py`` (writes ``tests/golden/*``) and by
``oracle/validate_oracle.py`` (checks the restatement in ``oracle/*.py`` against the
reference itself).  Stub list follows SURVEY.md §8(c) / Appendix A.
"""
from __future__ import annotations

import importlib
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
SCRATCH = os.environ.get("NLV_REF_SCRATCH", "/tmp/nlv_ref_scratch")

OBJ_CLASSES_FILE = os.path.join(REFERENCE_ROOT, "datasets/AG/object_classes.txt")
REL_CLASSES_FILE = os.path.join(REFERENCE_ROOT, "datasets/AG/relationship_classes.txt")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib"))


def _build_cython(scratch_lib: str) -> None:
    for sub, so_prefix in (("draw_rectangles", "draw_rectangles"),
                           ("fpn/box_intersections_cpu", "bbox")):
        d = os.path.join(scratch_lib, sub)
        if any(f.startswith(so_prefix) and f.endswith(".so") for f in os.listdir(d)):
            continue
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=d,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def prepare_scratch() -> str:
    """Copy reference/lib (+ dataloader-free bits) to the scratch dir and build Cython."""
    if not available():
        raise RuntimeError("reference tree not present; the harness only runs in the build container")
    dst = os.path.join(SCRATCH, "lib")
    if not os.path.isdir(dst):
        os.makedirs(SCRATCH, exist_ok=True)
        shutil.copytree(os.path.join(REFERENCE_ROOT, "lib"), dst)
        subprocess.run(["chmod", "-R", "u+w", SCRATCH], check=True)
    _build_cython(dst)
    return SCRATCH


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


_MHA_PATCHED = False


def patch_mha_int_mask(mode: str = "bool") -> None:
    """lib/transformer_wk.py:154 passes an int key_padding_mask, which torch>=2 rejects.
    mode='bool' -> true masking (lib/transformer.py:144 semantics, canonical);
    mode='additive' -> torch-1.10.1 emulation (mask value added to the logits)."""
    global _MHA_PATCHED
    orig = getattr(torch.nn.MultiheadAttention, "_nlv_orig_forward", None)
    if orig is None:
        orig = torch.nn.MultiheadAttention.forward
        torch.nn.MultiheadAttention._nlv_orig_forward = orig

    def fwd(self, query, key, value, key_padding_mask=None, **kw):
        if key_padding_mask is not None and key_padding_mask.dtype in (
                torch.int32, torch.int64, torch.uint8, torch.int16, torch.int8):
            key_padding_mask = (key_padding_mask.bool() if mode == "bool"
                                else key_padding_mask.to(query.dtype))
        return orig(self, query, key, value, key_padding_mask=key_padding_mask, **kw)

    torch.nn.MultiheadAttention.forward = fwd
    _MHA_PATCHED = True


def unpatch_mha() -> None:
    orig = getattr(torch.nn.MultiheadAttention, "_nlv_orig_forward", None)
    if orig is not None:
        torch.nn.MultiheadAttention.forward = orig


def load_reference(embed_seed: int = 1234, mha_mode: str = "bool"):
    """Return a namespace with the reference modules importable as ``ref.sttran`` etc."""
    scratch = prepare_scratch()
    if not hasattr(np, "float"):
        np.float = float  # lib/fpn/box_intersections_cpu/bbox.pyx:12
    if scratch not in sys.path:
        sys.path.insert(0, scratch)

    import torchvision

    class ROIAlign(torch.nn.Module):
        def __init__(self, output_size, spatial_scale, sampling_ratio):
            super().__init__()
            self.output_size, self.spatial_scale, self.sampling_ratio = output_size, spatial_scale, sampling_ratio

        def forward(self, x, rois):
            # bit-equal to the reference CPU RoIAlign (SURVEY.md §8c [probed])
            return torchvision.ops.roi_align(x, rois, self.output_size, self.spatial_scale,
                                             self.sampling_ratio, aligned=False)

    def nms(dets, scores, thr):
        raise RuntimeError("nms stub: build oracle/_ref for the first-party CPU nms")

    _stub("fasterRCNN"); _stub("fasterRCNN.lib"); _stub("fasterRCNN.lib.model")
    _stub("fasterRCNN.lib.model.roi_layers", ROIAlign=ROIAlign, nms=nms)
    _stub("lib.extract_bbox_features", extract_feature_given_bbox_base_feat_torch=None)
    for name in ("h5py", "tensorboardX", "termcolor", "yacs"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "easydict" not in sys.modules:
        try:
            importlib.import_module("easydict")
        except Exception:
            _stub("easydict", EasyDict=_EasyDict)

    def obj_edge_vectors(names, wv_type="glove.6B", wv_dir="data", wv_dim=200):
        g = torch.Generator().manual_seed(embed_seed)
        return torch.randn(len(names), wv_dim, generator=g)

    import lib.word_vectors as wv
    wv.obj_edge_vectors = obj_edge_vectors
    patch_mha_int_mask(mha_mode)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    ns = types.SimpleNamespace()
    ns.sttran = importlib.import_module("lib.sttran")
    ns.sttran.obj_edge_vectors = obj_edge_vectors
    ns.dsg_detr = importlib.import_module("lib.dsg_detr")
    ns.dsg_detr.obj_edge_vectors = obj_edge_vectors
    ns.transformer = importlib.import_module("lib.transformer")
    ns.transformer_wk = importlib.import_module("lib.transformer_wk")
    ns.evaluation_recall = importlib.import_module("lib.evaluation_recall")
    ns.draw_rectangles = importlib.import_module("lib.draw_rectangles.draw_rectangles")
    ns.bbox = importlib.import_module("lib.fpn.box_intersections_cpu.bbox")
    ns.box_utils = importlib.import_module("lib.fpn.box_utils")
    ns.matcher = importlib.import_module("lib.matcher")
    ns.track = importlib.import_module("lib.track")
    ns.AdamW = importlib.import_module("lib.AdamW")
    ns.obj_classes = ["__background__"] + open(OBJ_CLASSES_FILE).read().split()
    ns.rel_classes = open(REL_CLASSES_FILE).read().split()
    return ns


def build_reference_sttran(ref, mode: str):
    return ref.sttran.STTran(mode=mode, attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                             obj_classes=ref.obj_classes, enc_layer_num=1, dec_layer_num=3,
                             transformer_mode="wk", is_wks=True, feat_dim=2048, conf=None)


def build_reference_dsg(ref, mode: str):
    return ref.dsg_detr.STTran(mode=mode, attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                               obj_classes=ref.obj_classes)


def load_stable_evaluator(ref):
    """Scratch copy of lib/evaluation_recall.py with the two one-word stable-sort patches of SURVEY.md Appendix A.10
    (argsort(kind='stable') at :670-672 and a stable argsort_desc).  Nothing is written into the repository."""
    src = os.path.join(SCRATCH, "lib", "evaluation_recall.py")
    dst = os.path.join(SCRATCH, "lib", "evaluation_recall_stable.py")
    txt = open(src).read().replace("sorted_scores.argsort()[::-1]", 'sorted_scores.argsort(kind="stable")[::-1]')
    open(dst, "w").write(txt)
    mod = importlib.import_module("lib.evaluation_recall_stable")
    mod.argsort_desc = lambda s: np.column_stack(np.unravel_index(np.argsort(-s.ravel(), kind="stable"), s.shape))
    return mod
