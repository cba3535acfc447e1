"""TEST INFRASTRUCTURE ONLY -- builds libchatts_shim.so: a few kernel source files of chatts_b200/csrc compiled by g++ against the
"CUDA on CPU" shim (common.cuh in this directory).  The sources are COPIED next to the shim header (so their `#include "common.cuh"`
resolves to it) into a build directory and compiled (SUBSTITUTIONS lists the only textual changes made to a copy)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "chatts_b200", "csrc")
SOURCES = ["sampling.cu", "train_elementwise.cu", "allreduce_ll.cu", "attention_bwd.cu", "gemm_decode_fused.cu", "gemm_tcgen05.cu", "attention_bwd_tc5.cu"]
OUT = os.path.join(HERE, "_build", "libchatts_shim.so")
# Files that are GPU-validated are not edited for the shim's sake (not even a spelling): their two non-portable spellings are replaced
# in the COPY.  Everything else in the copy is the product source, byte for byte.
SUBSTITUTIONS = {
    "attention_bwd_tc5.cu": [("extern __shared__ uint8_t dq_raw[];", "uint8_t* dq_raw = g_dyn_smem;"),
                             ("extern __shared__ uint8_t dkv_raw[];", "uint8_t* dkv_raw = g_dyn_smem;")],
    "gemm_tcgen05.cu": [("extern __shared__ uint8_t smem_raw[];", "uint8_t* smem_raw = g_dyn_smem;"),
                        ('asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");', "tma_store_wait_read0();")],
}


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
        text = open(os.path.join(CSRC, s)).read()
        for old, new in SUBSTITUTIONS.get(s, []):
            assert old in text, (s, old)
            text = text.replace(old, new)
        open(dst, "w").write(text)
        cpps.append(dst)
    for f in ("common.cuh", "mma.h", "cooperative_groups.h", "tensormap.cuh", "shim_runtime.cpp"):
        shutil.copyfile(os.path.join(HERE, f), os.path.join(bdir, f))
    cmd = ["g++", "-std=c++17", "-O2", "-fno-strict-aliasing", "-g", "-fPIC", "-shared", "-pthread", "-w", "-I", bdir, "-I", os.path.join(ROOT, "include"), "-o", OUT,
           os.path.join(bdir, "shim_runtime.cpp")] + cpps
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("shim build failed:\n" + r.stderr[-6000:])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
