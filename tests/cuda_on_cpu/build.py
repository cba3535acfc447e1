"""TEST INFRASTRUCTURE ONLY -- builds libchatts_shim.so: a few kernel source files of chatts_b200/csrc compiled by g++ against the
"CUDA on CPU" shim (common.cuh in this directory).  The sources are COPIED next to the shim header (so their `#include "common.cuh"`
resolves to it) into a build directory and compiled unmodified."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "chatts_b200", "csrc")
SOURCES = ["sampling.cu", "train_elementwise.cu", "allreduce_ll.cu", "attention_bwd.cu", "gemm_decode_fused.cu"]
OUT = os.path.join(HERE, "_build", "libchatts_shim.so")


def build(force=False):
    bdir = os.path.dirname(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, f) for f in ("common.cuh", "mma.h", "cooperative_groups.h", "tensormap.cuh", "shim_runtime.cpp", "build.py")]
    deps.append(os.path.join(ROOT, "include", "chatts_b200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(bdir, exist_ok=True)
    cpps = []
    for s in SOURCES:
        dst = os.path.join(bdir, s.replace(".cu", ".cpp"))
        shutil.copyfile(os.path.join(CSRC, s), dst)
        cpps.append(dst)
    for f in ("common.cuh", "mma.h", "cooperative_groups.h", "tensormap.cuh", "shim_runtime.cpp"):
        shutil.copyfile(os.path.join(HERE, f), os.path.join(bdir, f))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-w", "-I", bdir, "-I", os.path.join(ROOT, "include"), "-o", OUT,
           os.path.join(bdir, "shim_runtime.cpp")] + cpps
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("shim build failed:\n" + r.stderr[-6000:])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
