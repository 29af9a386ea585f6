"""Builds libpsk_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libpsk_b200.so")
SOURCES = ["psk_craft.cu", "psk_scenario.cu", "psk_light.cu", "psk_host.cu"]
HEADERS = ["psk_common.cuh", os.path.join("..", "..", "include", "psk_craft.h"),
           os.path.join("..", "..", "include", "psk_light.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-shared", "-cudart", "shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    mt = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + ["build.py"]:
        p = os.path.join(HERE, f)
        if os.path.exists(p) and os.path.getmtime(p) > mt:
            return True
    return False


def build(force=False, verbose=False, out=None, extra=()):
    """``out`` / ``extra``: experiment builds (another file name, extra -D flags) for A/B runs."""
    if not force and not needs_build() and out is None:
        return LIB
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    cmd = ([nvcc_path()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) +
           ["-o", out or LIB] + srcs)
    subprocess.check_call(cmd, cwd=HERE)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
