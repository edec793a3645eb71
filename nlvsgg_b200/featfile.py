"""Packed per-video feature files.

The reference keeps one directory per frame with `dets.npy` + `feat.npy` (lib/assign_pseudo_label.py:27-45, python
lists re-assembled per frame every epoch) and rebuilds the `entry` dict on the host (:1196-1384).  Here a video is ONE
file whose sections are already the arrays the kernels consume, so a loader only has to `readinto` pinned staging
buffers at running offsets — no per-frame python, no dtype or layout conversion on the host:

    header   : magic, json (sizes, flags, section table)
    boxes    : f32 [N,5]            labels : i32 [N]          scores : f32 [N]
    dist     : (conf f32[N], other f32[N], idx i32[N]) when the distribution is a `create_dis` one (assign_pseudo_label.py:934-938;
               rebuilt bit-exactly on the device, csrc/util.cu), else f32 [N,36]
    features : bf16 [N,2048]
    pair_idx : i32 [R,2]            im_idx : i32 [R]
    union    : channels-last rows [R*49, 2048] bf16 — the operand layout of the union 1x1-conv GEMM, so the NCHW ->
               rows transposition of the fp32 entry contract disappears from the step — stored zero-suppressed
               (res5 output is post-ReLU): occupancy u64 [R*49,32] + the stored values, or dense when that is smaller.
               Stored values take 12 bits: low bytes u8 [nnz] + 4-bit codes u8 [nnz/2] = high byte (sign + 7 exponent bits)
               minus the row's base u8 [R*49]; a row's window of 16 high bytes = 32 binades below its largest value (post-ReLU
               activations use ~13).  Values outside the window (vanishing magnitudes, negative values) are listed as
               exceptions (position u32, value u16) and patched after the decode — lossless; more than 0.1 % exceptions:
               plain bf16 [nnz]
    labels   : attention values / list lengths (CSR), spatial / contacting multi-hot u32 [R]

bf16 storage: in the bf16 compute mode the device rounds both feature tensors to bf16 before their first use anyway, so
results are bit-identical to feeding the fp32 entry (tests/test_gpu_featfile.py).
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Optional

import numpy as np
import torch

from . import model as M

MAGIC = b"NLVF0002"
_ALIGN = 64


def _bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


def pack_union(union_feat: torch.Tensor, sparse: Optional[bool] = None, pack12: bool = True):
    """[R,2048,7,7] (any float dtype) -> dict of numpy sections for the channels-last (optionally zero-suppressed) layout."""
    R = union_feat.shape[0]
    rows = _bf16_bits(union_feat.permute(0, 2, 3, 1).reshape(R * 49, 2048))          # [R*49, 2048] u16
    nz = rows != 0                                                                    # -0.0 (0x8000) is kept: lossless
    nnz = int(nz.sum())
    sparse_bytes = R * 49 * 256 + nnz * 2
    if sparse is None:
        sparse = sparse_bytes < 0.9 * rows.size * 2
    if not sparse:
        return {"union_dense": rows}, {"union": "dense", "union_nnz": nnz}
    bitmap = np.packbits(nz, axis=1, bitorder="little").view(np.uint64).reshape(R * 49, 32)
    vals = rows[nz]
    rownnz = nz.sum(1).astype(np.uint32)
    if pack12 and nnz:
        # 12 bits per stored value: low byte + (high byte - the row's base), the base chosen so that the row's LARGEST positive
        # values fit the 16-step window [base, base + 15] (32 binades).  The few values outside it — magnitudes below 2^-29 of the
        # row's maximum, negative values — keep a zero slot in the stream and are listed as exceptions (position, value) that a
        # small kernel writes over the decoded rows.  Too many exceptions (> 0.1 % of the values): plain 16-bit values.
        hi, lo = (vals >> 8).astype(np.int16), (vals & 0xFF).astype(np.uint8)
        rr, cc = np.nonzero(nz)
        top = np.zeros(R * 49, dtype=np.int16)
        pos_mask = hi < 0x80
        np.maximum.at(top, rr[pos_mask], hi[pos_mask])
        base = np.maximum(top - 15, 0).astype(np.int16)
        code = hi - base[rr]
        exc = (code < 0) | (code > 15)
        n_exc = int(exc.sum())
        if n_exc <= max(16, nnz // 1000):
            exc_pos = (rr[exc].astype(np.int64) * 2048 + cc[exc]).astype(np.uint32)      # element index inside the video's rows
            exc_val = vals[exc].astype(np.uint16)
            code = np.where(exc, 0, code).astype(np.uint8)
            lo = np.where(exc, 0, lo).astype(np.uint8)
            if nnz & 1:                                   # the 4-bit plane of a video ends on a byte: one padding value, owned by
                lo, code = np.append(lo, np.uint8(0)), np.append(code, np.uint8(0))    # the last row (its bitmap ignores it)
                rownnz = rownnz.copy()
                rownnz[-1] += 1
            hx = (code[0::2] | (code[1::2] << 4)).astype(np.uint8)
            return ({"union_bitmap": bitmap, "union_lo": lo, "union_hx": hx, "union_base": base.astype(np.uint8), "union_rownnz": rownnz,
                     "union_exc_pos": exc_pos, "union_exc_val": exc_val},
                    {"union": "sparse12", "union_nnz": nnz, "union_nnz_stored": int(lo.size), "union_nexc": n_exc})
    return {"union_bitmap": bitmap, "union_vals": vals, "union_rownnz": rownnz}, {"union": "sparse", "union_nnz": nnz}


def _is_create_dis(dist: torch.Tensor):
    """(conf, other, idx) if every row of `dist` has the create_dis shape — one entry `conf`, the 35 others one common
    value (lib/assign_pseudo_label.py:934-938) — else None.  `other` is stored as the producer computed it, so the
    device rebuild is bit-exact whatever arithmetic (python double / fp32 tensor) produced (1 - conf) / 35."""
    n = dist.shape[0]
    if n == 0:
        return torch.zeros(0), torch.zeros(0), torch.zeros(0, dtype=torch.int64)
    conf, idx = dist.max(1)
    other = dist.gather(1, ((idx + 1) % 36)[:, None])[:, 0]
    rebuilt = other[:, None].expand(-1, 36).clone()
    nz = conf != 0
    rebuilt[torch.arange(n)[nz], idx[nz]] = conf[nz]
    return (conf, other, idx) if torch.equal(rebuilt, dist) else None


def write_video(path: str, entry: dict, sparse: Optional[bool] = None, pack12: bool = True) -> dict:
    """entry (the reference contract, lib/assign_pseudo_label.py:1368-1382, CPU tensors) -> one packed file.  Returns the header."""
    N, R = int(entry["boxes"].shape[0]), int(entry["pair_idx"].shape[0])
    sec: Dict[str, np.ndarray] = {
        "boxes": entry["boxes"].float().contiguous().numpy(),
        "labels": entry["labels"].to(torch.int32).numpy(),
        "scores": entry["scores"].float().numpy() if "scores" in entry else np.ones(N, np.float32),
        "features": _bf16_bits(entry["features"]),
        "pair_idx": entry["pair_idx"].to(torch.int32).contiguous().numpy(),
        "im_idx": entry["im_idx"].to(torch.int32).numpy(),
    }
    meta = {"n_boxes": N, "n_pairs": R, "im_idx_float": bool(entry["im_idx"].is_floating_point())}
    if "distribution" in entry:
        cd = _is_create_dis(entry["distribution"].float())
        if cd is not None:
            sec["dist_conf"], sec["dist_other"] = cd[0].numpy().astype(np.float32), cd[1].numpy().astype(np.float32)
            sec["dist_idx"] = cd[2].numpy().astype(np.int32)
            meta["dist"] = "create_dis"
        else:
            sec["dist_full"] = entry["distribution"].float().contiguous().numpy()
            meta["dist"] = "full"
    else:
        meta["dist"] = "none"
    usec, umeta = pack_union(entry["union_feat"], sparse, pack12)
    sec.update(usec)
    meta.update(umeta)
    if entry.get("attention_gt") is not None:
        avals, alens, sb, cb = M._label_csr([(entry["attention_gt"], entry["spatial_gt"], entry["contacting_gt"])])
        sec.update(att_vals=avals.astype(np.int32), att_lens=alens.astype(np.int32), spa_bits=sb, con_bits=cb)
        meta["has_labels"] = True
    table, off = {}, 0
    for k, a in sec.items():
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        table[k] = [off, int(a.nbytes), str(a.dtype), list(a.shape)]
        off += a.nbytes
    meta["sections"] = table
    hdr = json.dumps(meta).encode()
    pre = len(MAGIC) + 8 + len(hdr)
    pad = (-pre) % _ALIGN
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", len(hdr), pad) + hdr + b"\0" * pad)
        pos = 0
        for k, a in sec.items():
            o = table[k][0]
            f.write(b"\0" * (o - pos))
            f.write(np.ascontiguousarray(a).tobytes())
            pos = o + a.nbytes
    return meta


def read_header(path: str):
    with open(path, "rb") as f:
        head = f.read(len(MAGIC) + 8)
        if head[:len(MAGIC)] != MAGIC:
            raise ValueError(f"{path}: not a packed feature file")
        n, pad = struct.unpack("<II", head[len(MAGIC):])
        meta = json.loads(f.read(n))
    meta["_data0"] = len(MAGIC) + 8 + n + pad
    return meta


class _Pinned:
    """Grow-only pinned byte buffers keyed by name (the loader's staging memory; reused batch after batch)."""

    def __init__(self, pin: bool):
        self.pin, self.buf = pin, {}

    def get(self, name: str, nbytes: int) -> torch.Tensor:
        b = self.buf.get(name)
        if b is None or b.numel() < nbytes:
            b = self.buf[name] = torch.empty(int(nbytes * 1.25) + 64, dtype=torch.uint8, pin_memory=self.pin)
        return b[:nbytes]


class Loader:
    """Reads packed files straight into pinned staging buffers and hands out a host Batch (model.Batch) whose tensors are views
    of them.  `depth` rotating buffer sets so that a batch can still be in flight to the device while the next one is read."""

    def __init__(self, pin: bool = True, depth: int = 2):
        self.sets = [_Pinned(pin and torch.cuda.is_available()) for _ in range(depth)]
        self.i = 0

    def load(self, paths: List[str], mode: str = "sgdet") -> M.Batch:
        st = self.sets[self.i % len(self.sets)]
        self.i += 1
        metas = [read_header(p) for p in paths]
        nb = np.asarray([m["n_boxes"] for m in metas], dtype=np.int64)
        nr = np.asarray([m["n_pairs"] for m in metas], dtype=np.int64)
        N, R = int(nb.sum()), int(nr.sum())
        boff = np.concatenate(([0], np.cumsum(nb)))
        roff = np.concatenate(([0], np.cumsum(nr)))
        enc = metas[0]["union"]
        if any(m["union"] != enc for m in metas):
            raise ValueError("a batch must not mix union encodings (dense / sparse / sparse12)")
        sparse, s12 = enc in ("sparse", "sparse12"), enc == "sparse12"
        dist = metas[0]["dist"]
        if any(m["dist"] != dist for m in metas):
            raise ValueError("a batch must not mix distribution encodings")
        nnz = np.asarray([m["union_nnz_stored"] if s12 else m["union_nnz"] for m in metas], dtype=np.int64)
        voff = np.concatenate(([0], np.cumsum(nnz)))

        def buf(name, nbytes, dtype, shape):
            return st.get(name, max(int(nbytes), 16)).numpy()[:int(nbytes)].view(dtype).reshape(shape)

        out = {
            "boxes": buf("boxes", N * 20, np.float32, (N, 5)), "labels32": buf("labels32", N * 4, np.int32, (N,)),
            "scores": buf("scores", N * 4, np.float32, (N,)), "features": buf("features", N * 4096, np.uint16, (N, 2048)),
            "pair32": buf("pair32", R * 8, np.int32, (R, 2)), "im_idx": buf("im_idx", R * 4, np.int32, (R,)),
        }
        if dist == "create_dis":
            out["dist_conf"], out["dist_idx"] = buf("dist_conf", N * 4, np.float32, (N,)), buf("dist_idx", N * 4, np.int32, (N,))
            out["dist_other"] = buf("dist_other", N * 4, np.float32, (N,))
        elif dist == "full":
            out["dist_full"] = buf("dist_full", N * 144, np.float32, (N, 36))
        if sparse:
            out["union_bitmap"] = buf("union_bitmap", R * 49 * 256, np.uint64, (R * 49, 32))
            out["union_rownnz"] = buf("union_rownnz", R * 49 * 4, np.uint32, (R * 49,))
            if s12:
                out["union_lo"] = buf("union_lo", int(voff[-1]) + 32, np.uint8, (int(voff[-1]) + 32,))
                out["union_hx"] = buf("union_hx", int(voff[-1]) // 2 + 32, np.uint8, (int(voff[-1]) // 2 + 32,))
                out["union_base"] = buf("union_base", R * 49, np.uint8, (R * 49,))
                nexc = np.asarray([m.get("union_nexc", 0) for m in metas], dtype=np.int64)
                eoff = np.concatenate(([0], np.cumsum(nexc)))
                out["union_exc_pos"] = buf("union_exc_pos", int(eoff[-1]) * 4, np.uint32, (int(eoff[-1]),))
                out["union_exc_val"] = buf("union_exc_val", int(eoff[-1]) * 2, np.uint16, (int(eoff[-1]),))
            else:
                out["union_vals"] = buf("union_vals", int(voff[-1]) * 2 + 32, np.uint16, (int(voff[-1]) + 16,))
        else:
            out["union_dense"] = buf("union_dense", R * 49 * 4096, np.uint16, (R * 49, 2048))
        has_labels = all(m.get("has_labels") for m in metas)
        lab = {"att_vals": [], "att_lens": [], "spa_bits": [], "con_bits": []}
        for v, (p, m) in enumerate(zip(paths, metas)):
            b0, b1, r0, r1 = int(boff[v]), int(boff[v + 1]), int(roff[v]), int(roff[v + 1])
            dest = {"boxes": out["boxes"][b0:b1], "labels": out["labels32"][b0:b1], "scores": out["scores"][b0:b1],
                    "features": out["features"][b0:b1], "pair_idx": out["pair32"][r0:r1], "im_idx": out["im_idx"][r0:r1]}
            if dist == "create_dis":
                dest["dist_conf"], dest["dist_idx"] = out["dist_conf"][b0:b1], out["dist_idx"][b0:b1]
                dest["dist_other"] = out["dist_other"][b0:b1]
            elif dist == "full":
                dest["dist_full"] = out["dist_full"][b0:b1]
            if sparse:
                dest["union_bitmap"] = out["union_bitmap"][r0 * 49:r1 * 49]
                dest["union_rownnz"] = out["union_rownnz"][r0 * 49:r1 * 49]
                if s12:
                    dest["union_lo"] = out["union_lo"][int(voff[v]):int(voff[v + 1])]
                    dest["union_hx"] = out["union_hx"][int(voff[v]) // 2:int(voff[v + 1]) // 2]
                    dest["union_base"] = out["union_base"][r0 * 49:r1 * 49]
                    dest["union_exc_pos"] = out["union_exc_pos"][int(eoff[v]):int(eoff[v + 1])]
                    dest["union_exc_val"] = out["union_exc_val"][int(eoff[v]):int(eoff[v + 1])]
                else:
                    dest["union_vals"] = out["union_vals"][int(voff[v]):int(voff[v + 1])]
            else:
                dest["union_dense"] = out["union_dense"][r0 * 49:r1 * 49]
            with open(p, "rb", buffering=0) as f:
                for name, arr in dest.items():
                    o, nbytes, _, _ = m["sections"][name]
                    if nbytes == 0:
                        continue
                    f.seek(m["_data0"] + o)
                    got = f.readinto(memoryview(arr.reshape(-1).view(np.uint8)))
                    if got != nbytes:
                        raise IOError(f"{p}: short read of section {name}")
                if has_labels:
                    for name in lab:
                        o, nbytes, dt, shape = m["sections"][name]
                        f.seek(m["_data0"] + o)
                        lab[name].append(np.frombuffer(f.read(nbytes), dtype=dt).reshape(shape))
            out["pair32"][r0:r1] += b0                     # pair indices become batch-global box rows
            if s12 and eoff[v + 1] > eoff[v]:
                assert (r1 * 49) * 2048 < 2 ** 32
                out["union_exc_pos"][int(eoff[v]):int(eoff[v + 1])] += np.uint32(r0 * 49 * 2048)     # ... and exception positions batch-global

        hb = M.Batch()
        hb.n_boxes, hb.n_pairs = [int(x) for x in nb], [int(x) for x in nr]
        hb.frame_ids = [out["im_idx"][int(roff[v]):int(roff[v + 1])].astype(np.int64) for v in range(len(paths))]
        hb.gt_lists = [(None, None, None)] * len(paths)
        hb.lab_csr = None
        if has_labels:
            hb.lab_csr = (np.concatenate(lab["att_vals"]).astype(np.int64), np.concatenate(lab["att_lens"]).astype(np.int64),
                          np.concatenate(lab["spa_bits"]), np.concatenate(lab["con_bits"]))
        T = torch.from_numpy
        hb.features = T(out["features"].view(np.int16)).view(torch.bfloat16)
        hb.boxes, hb.scores = T(out["boxes"]), T(out["scores"])
        # int64 copies of the two index arrays the kernels take as int64 (N and 2R elements: small); staged like the rest
        lab64 = buf("labels64", N * 8, np.int64, (N,)); lab64[:] = out["labels32"]
        pair64 = buf("pair64", R * 16, np.int64, (R, 2)); pair64[:] = out["pair32"]
        hb.labels, hb.pair_idx = T(lab64), T(pair64)
        hb.distribution = T(out["dist_full"]) if dist == "full" else None
        hb.dist_conf = T(out["dist_conf"]) if dist == "create_dis" else None
        hb.dist_idx = T(out["dist_idx"]) if dist == "create_dis" else None
        hb.dist_other = T(out["dist_other"]) if dist == "create_dis" else None
        hb.spatial_masks = None
        if sparse:
            offs = buf("union_off", (R * 49 + 1) * 4, np.uint32, (R * 49 + 1,))
            offs[0] = 0
            np.cumsum(out["union_rownnz"], out=offs[1:])
            hb.union_bitmap = T(out["union_bitmap"].view(np.int64))
            hb.union_off = T(offs.view(np.int32))
            if s12:
                out["union_lo"][int(voff[-1]):] = 0
                out["union_hx"][int(voff[-1]) // 2:] = 0
                hb.union_feat, hb.union_hx, hb.union_base = T(out["union_lo"]), T(out["union_hx"]), T(out["union_base"])
                hb.union_exc_pos = T(out["union_exc_pos"].view(np.int32)) if eoff[-1] else None
                hb.union_exc_val = T(out["union_exc_val"].view(np.int16)) if eoff[-1] else None
                hb.union_rows = 3
            else:
                out["union_vals"][int(voff[-1]):] = 0
                hb.union_feat = T(out["union_vals"].view(np.int16)).view(torch.bfloat16)
                hb.union_rows = 2
        else:
            hb.union_feat = T(out["union_dense"].view(np.int16)).view(torch.bfloat16)
            hb.union_bitmap = hb.union_off = None
            hb.union_rows = 1
        return hb


def write_videos(dirpath: str, entries: List[dict], sparse: Optional[bool] = None, pack12: bool = True) -> List[str]:
    os.makedirs(dirpath, exist_ok=True)
    paths, encs = [], []
    for i, e in enumerate(entries):
        p = os.path.join(dirpath, f"video_{i:05d}.nlvf")
        encs.append(write_video(p, e, sparse, pack12)["union"])
        paths.append(p)
    if "sparse12" in encs and "sparse" in encs:      # a loader batch holds one encoding: the set falls back to 16-bit values together
        for p, e, enc in zip(paths, entries, encs):
            if enc == "sparse12":
                write_video(p, e, True, False)
    return paths
