"""Build libnsr_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the product is a plain C-ABI shared library).  Used by __graft_entry__.build()."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libnsr_b200.so")
SOURCES = ["nsr_api.cu", "nsr_simt.cu", "nsr_tc.cu", "nsr_train.cu", "nsr_comm.cu"]
HEADERS = ["nsr_internal.h", "nsr_device.cuh", "nsr_tc_ptx.cuh", "nsr_tc_mma.cuh", "nsr_jet_lut.h", os.path.join(ROOT, "include", "nsr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-diag-suppress", "177"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, extra_flags=(), lib_path: str = LIB) -> str:
    """Compile every CUDA source and link the shared library.  Returns its path.
    extra_flags/lib_path build a variant (e.g. the trace build: -DNSR_TC_TRACE=1)."""
    LIB = lib_path
    NVCC_FLAGS = [*globals()["NVCC_FLAGS"], *extra_flags]
    stamp = LIB + ".stamp"
    dig = _digest() + "|" + " ".join(extra_flags)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(ROOT, "build", os.path.basename(LIB).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build(force=True, verbose=True, extra_flags=["-DNSR_TC_TRACE=1"],
                    lib_path=os.path.join(PKG, "libnsr_b200_trace.so")))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
