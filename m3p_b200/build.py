"""Build libm3p_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Usage: python -m m3p_b200.build [--force] [--verbose]
The library has no torch / libcuda link dependency (static cudart), so it loads anywhere; compute
entry points obviously need a B200.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libm3p_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-I", "/usr/local/graft",  # cudaTypedefs.h fallback location on this image
] + os.environ.get("M3P_NVCC_EXTRA", "").split()  # e.g. -DM3P_ATTN_TRACE for the phase-timing printfs


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    with open(path, "rb") as f:
        h.update(f.read())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "m3p_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _compile_one(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(c for _, c in results)
    if rebuilt or not os.path.exists(LIB_PATH):
        cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                          "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
