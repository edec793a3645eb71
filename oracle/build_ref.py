"""TEST INFRASTRUCTURE — compiles the reference's OWN native code from the sources where they lie under /root/reference
into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).  No reference source enters the repository.

  * lib/draw_rectangles/draw_rectangles.pyx, lib/fpn/box_intersections_cpu/bbox.pyx  (Cython, built in a /tmp scratch copy)
  * fasterRCNN/lib/model/csrc/{vision.cpp, cpu/ROIAlign_cpu.cpp, cpu/nms_cpu.cpp}    (pybind11/ATen CPU half; the CUDA half
    needs THC, which torch >= 1.11 no longer ships).  Two ATen API tokens are patched in the scratch copy
    (`.type()` -> `.scalar_type()` inside AT_DISPATCH_FLOATING_TYPES, cpu/ROIAlign_cpu.cpp:242 and cpu/nms_cpu.cpp:71).

    python oracle/build_ref.py
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
sys.path.insert(0, os.path.dirname(HERE))


def main():
    from oracle import ref_harness as H
    if not H.available():
        print("reference tree absent: nothing to build")
        return
    os.makedirs(REF, exist_ok=True)
    scratch = H.prepare_scratch()
    for pat in ("lib/draw_rectangles/draw_rectangles*.so", "lib/fpn/box_intersections_cpu/bbox*.so"):
        for f in glob.glob(os.path.join(scratch, pat)):
            shutil.copy2(f, REF)
    if not glob.glob(os.path.join(REF, "nlv_ref_C*.so")):
        csrc = os.path.join(scratch, "csrc")
        if not os.path.isdir(csrc):
            shutil.copytree(os.path.join(H.REFERENCE_ROOT, "fasterRCNN/lib/model/csrc"), csrc)
            subprocess.run(["chmod", "-R", "u+w", csrc], check=True)
            for f, old, new in (("cpu/ROIAlign_cpu.cpp", "AT_DISPATCH_FLOATING_TYPES(input.type()", "AT_DISPATCH_FLOATING_TYPES(input.scalar_type()"),
                                ("cpu/nms_cpu.cpp", "AT_DISPATCH_FLOATING_TYPES(dets.type()", "AT_DISPATCH_FLOATING_TYPES(dets.scalar_type()")):
                p = os.path.join(csrc, f)
                s = open(p).read()
                assert old in s, f
                open(p, "w").write(s.replace(old, new))
        from torch.utils import cpp_extension
        bdir = os.path.join(scratch, "build_C")
        os.makedirs(bdir, exist_ok=True)
        cpp_extension.load(name="nlv_ref_C", sources=[os.path.join(csrc, "vision.cpp")] + sorted(glob.glob(os.path.join(csrc, "cpu/*.cpp"))),
                           extra_include_paths=[csrc], build_directory=bdir, verbose=False)
        for f in glob.glob(os.path.join(bdir, "nlv_ref_C*.so")):
            shutil.copy2(f, REF)
    print("oracle/_ref:", sorted(os.listdir(REF)))


if __name__ == "__main__":
    main()
