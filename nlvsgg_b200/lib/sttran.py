"""Drop-in for lib/sttran.py (reference lines cited inline).  Same constructor, same ``forward(entry)`` dict
contract, same state_dict names; the torch submodules below only *hold* the parameters — every FLOP of
forward and backward runs in the nlv_b200 kernels (nlvsgg_b200/engine.py)."""
import os

import torch
import torch.nn as nn

from .. import autograd as A
from .. import engine as E
from .transformer_wk import transformer_wk
from .word_vectors import obj_edge_vectors


class ObjectClassifier(nn.Module):
    """Parameter container of lib/sttran.py:20-51 (predcls pass-through :90-92, sgdet/wks head :173-184)."""

    def __init__(self, mode="sgdet", obj_classes=None, is_wks=True):
        super().__init__()
        self.classes, self.mode, self.is_wks = obj_classes, mode, is_wks
        embed_vecs = obj_edge_vectors(obj_classes[1:], wv_type="glove.6B", wv_dir="data", wv_dim=200)
        self.obj_embed = nn.Embedding(len(obj_classes) - 1, 200)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=0.01 / 10.0), nn.Linear(4, 128), nn.ReLU(inplace=True),
                                       nn.Dropout(0.1))
        self.obj_dim = 2048
        self.decoder_lin = nn.Sequential(nn.Linear(self.obj_dim + 200 + 128, 1024), nn.BatchNorm1d(1024), nn.ReLU(),
                                         nn.Linear(1024, len(self.classes)))


class STTran(nn.Module):
    def __init__(self, mode="sgdet", attention_class_num=None, spatial_class_num=None, contact_class_num=None,
                 obj_classes=None, enc_layer_num=None, dec_layer_num=None, transformer_mode=None, is_wks=True,
                 feat_dim=2048, motifs_path=None, conf=None, precision=None):
        super().__init__()
        assert mode in ("sgdet", "sgcls", "predcls")          # lib/sttran.py:328
        self.conf, self.obj_classes, self.mode, self.is_wks = conf, obj_classes, mode, is_wks
        self.attention_class_num, self.spatial_class_num, self.contact_class_num = \
            attention_class_num, spatial_class_num, contact_class_num
        self.transformer_mode, self.motifs_path = transformer_mode, motifs_path
        if mode == "sgcls" or (mode == "sgdet" and not is_wks):
            raise NotImplementedError("only the predcls and sgdet/is_wks branches of lib/sttran.py are built (SURVEY.md §8f-3)")
        assert (attention_class_num, spatial_class_num, contact_class_num) == (3, 6, 17) and feat_dim == 2048
        self.object_classifier = ObjectClassifier(mode=mode, obj_classes=obj_classes, is_wks=is_wks)
        self.union_func1 = nn.Conv2d(feat_dim, 256, 1, 1)
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
        self.glocal_transformer = transformer_wk(enc_layer_num=enc_layer_num, dec_layer_num=dec_layer_num, embed_dim=1936,
                                                 nhead=8, dim_feedforward=2048, dropout=0.1, mode="latter")
        self.a_rel_compress = nn.Linear(1936, attention_class_num)
        self.s_rel_compress = nn.Linear(1936, spatial_class_num)
        self.c_rel_compress = nn.Linear(1936, contact_class_num)
        self.kernels = E.Kernels(precision or os.environ.get("NLV_PRECISION", "bf16"))

    def forward(self, entry):
        """lib/sttran.py:375-411: mutates and returns ``entry``."""
        obj, att, spa, con, batch = A.run_module(self, self.kernels, [entry], self.mode, "sttran")
        entry["pred_labels"] = entry["labels"]                      # :91 / :183
        if self.mode != "predcls":
            entry["distribution"] = obj                            # :182  logits [N,37]
            entry["pred_scores"] = entry["scores"]                  # :184
        if "spatial_masks" not in entry:
            entry["spatial_masks"] = batch.spatial_masks
        entry["attention_distribution"] = att                       # :404 (logits)
        entry["spatial_distribution"] = spa                         # :408
        entry["contacting_distribution"] = con                      # :409
        return entry
