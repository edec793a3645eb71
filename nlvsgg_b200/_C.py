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


class GemmArgs(ctypes.Structure):
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("d", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("residual", ctypes.c_void_p),
                ("m", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int),
                ("lda", ctypes.c_int), ("ldb", ctypes.c_int), ("ldd", ctypes.c_int), ("ldr", ctypes.c_int),
                ("a_major", ctypes.c_int), ("b_major", ctypes.c_int),
                ("ab_dtype", ctypes.c_int), ("d_dtype", ctypes.c_int), ("r_dtype", ctypes.c_int),
                ("relu", ctypes.c_int), ("gate", ctypes.c_void_p), ("ldg", ctypes.c_int), ("gate_dtype", ctypes.c_int)]


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
