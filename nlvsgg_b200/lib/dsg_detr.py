"""Drop-in for lib/dsg_detr.py:STTran (DSG-DETR): same constructor (the six extra keyword arguments that
tools/test_DSG_DETR.py:39-50 passes are accepted and ignored), same ``forward(entry)`` contract, same state_dict
names.  The sgdet path (is_wks hard-coded, dsg_detr.py:89,277-288) and predcls-without-object-tracks are built; the
object-track encoder (``object_classifier.encoder_tran``, used only by predcls/sgcls with ``entry['indices']``) is
carried as parameters for checkpoint compatibility."""
import os

import torch
import torch.nn as nn

from .. import autograd as A
from .. import engine as E
from ..shapes import sinusoidal_pe
from .word_vectors import obj_edge_vectors


class PositionalEncoding(nn.Module):
    """lib/dsg_detr.py:25-48 (buffer only; the addition happens inside the row-gather kernel)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", sinusoidal_pe(max_len, d_model))


class ObjectClassifier(nn.Module):
    def __init__(self, mode="sgdet", obj_classes=None):
        super().__init__()
        self.classes, self.mode, self.is_wks = obj_classes, mode, True
        embed_vecs = obj_edge_vectors(obj_classes[1:], wv_type="glove.6B", wv_dir="data", wv_dim=200)
        self.obj_embed = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=0.01 / 10.0), nn.Linear(4, 128), nn.ReLU(inplace=True),
                                       nn.Dropout(0.1))
        d_model = 2048 + 200 + 128
        self.positional_encoder = PositionalEncoding(d_model, 0.1, 600)
        encoder_layer = nn.TransformerEncoderLayer(d_model=d_model, dim_feedforward=1024, nhead=8, batch_first=True)
        self.encoder_tran = nn.TransformerEncoder(encoder_layer, num_layers=3)
        self.decoder_lin = nn.Sequential(nn.Linear(d_model, 1024), nn.BatchNorm1d(1024), nn.ReLU(),
                                         nn.Linear(1024, len(self.classes)))


class STTran(nn.Module):
    def __init__(self, mode="sgdet", attention_class_num=None, spatial_class_num=None, contact_class_num=None,
                 obj_classes=None, precision=None, **_ignored):
        super().__init__()
        assert mode in ("sgdet", "sgcls", "predcls")
        if mode == "sgcls":
            # lib/dsg_detr.py:185-275 feeds 2376-d object-track rows into subj_fc = Linear(2048, 512) (:181, :486): the
            # reference itself raises a shape error on this path
            raise NotImplementedError("DSG-DETR sgcls is unreachable in the reference (lib/dsg_detr.py:181 vs :486)")
        self.obj_classes, self.mode = obj_classes, mode
        self.attention_class_num, self.spatial_class_num, self.contact_class_num = \
            attention_class_num, spatial_class_num, contact_class_num
        assert (attention_class_num, spatial_class_num, contact_class_num) == (3, 6, 17)
        self.object_classifier = ObjectClassifier(mode=mode, obj_classes=obj_classes)
        self.union_func1 = nn.Conv2d(2048, 256, 1, 1)
        self.conv = nn.Sequential(
            nn.Conv2d(2, 256 // 2, kernel_size=7, stride=2, padding=3, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256 // 2, momentum=0.01), nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(256 // 2, 256, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256, momentum=0.01))
        self.subj_fc, self.obj_fc, self.vr_fc = nn.Linear(2048, 512), nn.Linear(2048, 512), nn.Linear(256 * 7 * 7, 512)
        embed_vecs = obj_edge_vectors(obj_classes, wv_type="glove.6B", wv_dir="data", wv_dim=200)
        self.obj_embed = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.obj_embed2 = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed2.weight.data = embed_vecs.clone()
        d_model = 1936
        self.positional_encoder = PositionalEncoding(d_model, max_len=400)
        self.global_transformer = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=d_model, dim_feedforward=2048, nhead=8, batch_first=True), num_layers=3)
        self.local_transformer = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=d_model, dim_feedforward=2048, nhead=8, batch_first=True), num_layers=1)
        self.a_rel_compress = nn.Linear(d_model, attention_class_num)
        self.s_rel_compress = nn.Linear(d_model, spatial_class_num)
        self.c_rel_compress = nn.Linear(d_model, contact_class_num)
        self.kernels = E.Kernels(precision or os.environ.get("NLV_PRECISION", "bf16"), dropout=0.1)

    def forward(self, entry):
        """lib/dsg_detr.py:514-572: mutates ``entry`` and returns it."""
        obj, att, spa, con, batch = A.run_module(self, self.kernels, [entry], self.mode, "dsg")
        entry["pred_labels"] = entry["labels"]
        if self.mode != "predcls":
            entry["distribution"] = obj
            entry["pred_scores"] = entry["scores"]
        if "spatial_masks" not in entry:
            entry["spatial_masks"] = batch.spatial_masks
        entry["attention_distribution"] = att
        entry["spatial_distribution"] = spa
        entry["contacting_distribution"] = con
        return entry
