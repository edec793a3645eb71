"""ctypes binding of the C-ABI library (include/nlv_b200.h).

The library is the product: there is no Python/torch fallback.  Importing this module on a box
with a GPU but without the built library raises; calling a kernel entry point without CUDA raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_build", "libnlv_b200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NLV_F32, NLV_BF16 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
FORCE_SIMT = 0x100

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into _build/libnlv_b200.so (in-tree)."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(INCLUDE, "nlv_b200.h")]
    if not force and os.path.exists(SO_PATH) and all(os.path.getmtime(SO_PATH) >= os.path.getmtime(d) for d in deps):
        return SO_PATH
    os.makedirs(os.path.dirname(SO_PATH), exist_ok=True)
    objs = []
    procs = []
    for s in srcs:  # one nvcc per translation unit, in parallel
        o = os.path.join(os.path.dirname(SO_PATH), os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d)
                                                   for d in [s] + deps[len(srcs):]):
            continue
        cmd = ["nvcc"] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", "-o", o, s]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    subprocess.run(["nvcc", "-Wno-deprecated-gpu-targets", "-shared", "-o", SO_PATH] + objs, check=True)
    return SO_PATH


class Dropout(ctypes.Structure):
    """nlv_dropout (include/nlv_b200.h)."""
    _fields_ = [("thr16", ctypes.c_uint), ("scale", ctypes.c_float), ("seed_lo", ctypes.c_uint), ("seed_hi", ctypes.c_uint),
                ("stream", ctypes.c_uint)]

    @classmethod
    def make(cls, p: float, seed: int, stream: int) -> "Dropout":
        d = cls()
        if p > 0:
            d.thr16 = min(65535, int(p * 65536.0 + 0.5))
            d.scale = 1.0 / (1.0 - p)
            d.seed_lo, d.seed_hi, d.stream = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, stream
        return d


class GemmArgs(ctypes.Structure):
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("d", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("residual", ctypes.c_void_p),
                ("m", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int),
                ("lda", ctypes.c_int), ("ldb", ctypes.c_int), ("ldd", ctypes.c_int), ("ldr", ctypes.c_int),
                ("a_major", ctypes.c_int), ("b_major", ctypes.c_int),
                ("ab_dtype", ctypes.c_int), ("d_dtype", ctypes.c_int), ("r_dtype", ctypes.c_int),
                ("relu", ctypes.c_int), ("gate", ctypes.c_void_p), ("ldg", ctypes.c_int), ("gate_dtype", ctypes.c_int),
                ("gate_scale", ctypes.c_float), ("drop", Dropout)]


_vp, _ip, _i, _ll, _f = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class Model(ctypes.Structure):
    """nlv_model (include/nlv_b200.h)."""
    _fields_ = [("arch", _i), ("mode", _i), ("precision", _i), ("n_enc", _i), ("n_dec", _i), ("training", _i), ("n_slots", _i),
                ("params", _vp), ("params_op", _vp), ("grad_base", _vp), ("grad_elems", _ll), ("grad_offset", _vp),
                ("dropout_p", _f), ("seed", ctypes.c_ulonglong), ("additive_mask", _i), ("pe_rows", _i), ("transformer_both", _i)]


class BatchDesc(ctypes.Structure):
    """nlv_batch (include/nlv_b200.h)."""
    _fields_ = [("nv", _i), ("n_boxes", _ll), ("n_pairs", _ll), ("n_stream", _ll),
                ("features", _vp), ("feat_dtype", _i), ("boxes", _vp), ("labels", _vp), ("distribution", _vp),
                ("union_feat", _vp), ("union_dtype", _i), ("union_rows", _i), ("spatial_masks", _vp), ("pair_idx", _vp),
                ("box_seg", _ip), ("seg196", _ip), ("seg49", _ip), ("box_row", _ip), ("pair_row", _ip),
                ("local_work", _ip), ("n_local_work", _i), ("glob_work", _ip), ("n_glob_work", _i),
                ("stream_src", _ip), ("stream_slot", _ip), ("inv", _ip), ("out_src", _ip), ("out_inv", _ip), ("passthrough", _ip),
                ("has_passthrough", _i),
                ("cls_perm", _ip), ("cls_iperm", _ip), ("cls_pos", _ip), ("cls_work", _ip), ("n_cls_work", _i),
                ("union_bitmap", _vp), ("union_off", _vp), ("dist_conf", _vp), ("dist_other", _vp), ("dist_idx", _vp),
                ("lab_att", _vp), ("w_att", _vp), ("spa_bits", _vp), ("w_spa", _vp), ("con_bits", _vp), ("w_con", _vp), ("w_obj", _vp),
                ("both_w", _vp), ("work_sorted", _i), ("n_local_long", _i), ("n_glob_long", _i), ("n_cls_long", _i),
                ("union_hx", _vp), ("union_base", _vp), ("union_exc_pos", _vp), ("union_exc_val", _vp), ("n_union_exc", _i)]


class Outputs(ctypes.Structure):
    """nlv_outputs (include/nlv_b200.h)."""
    _fields_ = [(n, _vp) for n in ("obj_logits", "logits26", "att", "spa", "con", "loss", "masks", "rel_tokens", "rel_out", "d26", "dobj")]


RUN_CTX, RUN_LOSS, RUN_BACKWARD, RUN_ACTIVATIONS, RUN_OBJECT_ONLY = 1, 2, 4, 8, 16
ARCH = {"sttran": 0, "dsg": 1}
MODE = {"predcls": 0, "sgcls": 1, "sgdet": 2}
PREC = {"bf16": 0, "bf16x3": 1, "fp32": 2}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"nlv_b200: {SO_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU fallback.")
        _lib = ctypes.CDLL(SO_PATH)
        _lib.nlv_last_error.restype = ctypes.c_char_p
        _lib.nlv_launch_count.restype = ctypes.c_longlong
        _lib.nlv_session_create.restype = ctypes.c_void_p
        _lib.nlv_session_destroy.argtypes = [ctypes.c_void_p]
        _lib.nlv_session_plan.restype = ctypes.c_longlong
        _lib.nlv_session_plan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.nlv_session_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.nlv_session_set_gradients.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int]
        _lib.nlv_session_backward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.nlv_session_transformer_forward.restype = ctypes.c_longlong
        _lib.nlv_session_transformer_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                         ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.nlv_session_transformer_backward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"nlv_b200 {what} failed ({rc}): {lib().nlv_last_error().decode()}")


def launch_count() -> int:
    return int(lib().nlv_launch_count())


def declared_symbols():
    """Entry points declared in include/nlv_b200.h (used by the CPU-side export test)."""
    import re
    txt = open(os.path.join(INCLUDE, "nlv_b200.h")).read()
    return sorted(set(re.findall(r"\b(nlv_[a-z0-9_]+)\s*\(", txt)))
