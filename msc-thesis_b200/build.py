"""Builds libheli3dof.so (the C-ABI CUDA library, include/heli3dof.h) in-tree with nvcc for sm_100a.

    python msc-thesis_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
INCLUDE = os.path.join(ROOT, "include")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libheli3dof.so")

SOURCES = ("h3_api.cu", "h3_loop_f64.cu", "h3_loop_mixed.cu", "h3_loop_f32.cu", "h6_api.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", INCLUDE]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    files.append(os.path.abspath(__file__))
    return files


def up_to_date():
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(f) <= t for f in _deps())


def build(force=False, verbose=False, extra_flags=(), out=None, sources=SOURCES):
    """Compile every .cu for sm_100a and link the shared library; returns its path.  `out` / `extra_flags`
    build a tuning variant next to the default library (selected at run time with H3_LIB_PATH)."""
    global LIB_PATH, BUILD_DIR
    if out is not None:
        LIB_PATH = os.path.join(PKG_DIR, out)
        BUILD_DIR = os.path.join(PKG_DIR, "build", os.path.splitext(out)[0])
        force = True
    if not force and up_to_date():
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(sources)) as ex:
        results = list(ex.map(compile_one, sources))
    objs = []
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-Xcompiler", "-fPIC", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    if "--fast" in sys.argv:
        flags.append("-DH3_FAST_BUILD")
    out = None
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            out = a.split("=", 1)[1]
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, extra_flags=flags, out=out)
    print(path)
