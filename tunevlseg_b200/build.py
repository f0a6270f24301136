"""Build libtvs_b200.so in-tree with nvcc for sm_100a (cross-compiles on a machine without a GPU).

    python -m tunevlseg_b200.build [--force]

The shared library exports only the C ABI declared in include/tvs_b200.h; it links the static CUDA runtime and
resolves cuTensorMapEncodeTiled through cudaGetDriverEntryPoint at run time, so it loads (for symbol checks) on a
machine without libcuda.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(OUT_DIR, "libtvs_b200.so")
SOURCES = ["api.cu", "gemm_sm100.cu", "attention.cu", "attention_sm100.cu", "layernorm.cu", "loss_metrics.cu", "elementwise.cu", "conv_ops.cu", "ffn_sm100.cu", "preprocess.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "sm100_ptx.cuh"), os.path.join(ROOT, "include", "tvs_b200.h")]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--exclude-libs=ALL"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
