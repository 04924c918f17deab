"""Build csrc/libmxprune.so for sm_100a with nvcc (in-tree, so the .so travels with the repo).

Every .cu file of csrc/ is one translation unit; they are compiled in parallel nvcc processes and
linked into one shared library (no relocatable device code: no kernel calls across units)."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(CSRC, "libmxprune.so")
OBJDIR = os.path.join(CSRC, "build")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + \
           [os.path.join(ROOT, "include", "mxprune.h")]


def _newer(deps, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in _sources()]
    if not force and not _newer(srcs + _headers(), LIB):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
             "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include")]
    os.makedirs(OBJDIR, exist_ok=True)
    procs, objs = [], []
    for src in srcs:
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer([src] + _headers(), obj):
            cmd = [nvcc] + flags + ["-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    link = [nvcc, "-shared", "-o", LIB] + objs
    if verbose:
        print(" ".join(link), file=sys.stderr)
    subprocess.run(link, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
