"""In-tree build of libdynhor_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m dynhor_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdynhor_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(HERE, "..", "include")]
# per-source extra flags.  --fmad=false: the raster / pose arithmetic is defined as un-fused IEEE fp32
# (dh_core.h); FMAs appear only where fmaf() is written.
SOURCES = {
    "dh_api.cu": [],
    "dh_jointopt.cu": ["--fmad=false"],
    "dh_dino.cu": [],
    "dh_corr.cu": [],
    "dh_roi.cu": ["--fmad=false"],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def build(force=False, verbose=False, only=None):
    """only: recompile just these sources (the other objects must exist) and relink -- kernel-variant sweeps."""
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dynhor_b200.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    objs = []
    for s in srcs:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        if only and s not in only and os.path.exists(obj):
            objs.append(obj)
            continue
        cmd = [_nvcc()] + ARCH + COMMON + SOURCES[s] + os.environ.get("DH_EXTRA_NVCC_FLAGS", "").split() + \
              (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [_nvcc()] + ARCH + ["-shared", "-o", OUT] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if a.endswith(".cu")]
    print(build(force="--force" in sys.argv or bool(only), verbose="--verbose" in sys.argv, only=only or None))
