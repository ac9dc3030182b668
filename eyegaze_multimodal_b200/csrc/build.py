"""Builds libeyegaze_b200.so (and the standalone device tests) with nvcc for sm_100a, in-tree.

Usage: python -m eyegaze_multimodal_b200.csrc.build [--tests] [--force]
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(PKG, "libeyegaze_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("EGB_NVCC_DEFS", "").split()      # e.g. -DEGB_ATT_TIMING for the in-kernel phase timers


def _sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG, "..", "include", "eyegaze_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, hm):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(HERE, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), hm):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(tests=False, force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, hm), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if tests:
        tdir = os.path.join(HERE, "tests")
        for f in sorted(os.listdir(tdir)):
            if not f.endswith(".cu"):
                continue
            exe = os.path.join(OBJ, f[:-3])
            src = os.path.join(tdir, f)
            if not force and os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
                continue
            cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", src, "-o", exe,
                   "-L" + PKG, "-leyegaze_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../.."]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("test build failed for %s:\n%s\n%s" % (f, r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    lib = build(tests="--tests" in sys.argv, force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", lib)
