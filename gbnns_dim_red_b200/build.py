"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

    libgbdr.so      CUDA kernels + C ABI          nvcc -gencode arch=compute_100a,code=sm_100a
    host/bin/*      drop-in C++ drivers           g++ against libgbdr.so
    liboracle.so    CPU checker (tests only)      gcc, strict IEEE
    oracle/_ref/*   the reference's own headers   g++ (only where /root/reference exists)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gbnns_dim_red_b200", "csrc")
LIB = os.path.join(ROOT, "gbnns_dim_red_b200", "libgbdr.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
    # the image's default host compiler wrapper lacks OpenMP specs; the system one is complete
    "-ccbin", "/usr/bin/g++",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    out = _sources()
    out += sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    out.append(os.path.join(ROOT, "include", "gbdr.h"))
    return out


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libgbdr.so (incremental per translation unit)."""
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    headers = [p for p in _deps() if not p.endswith(".cu")]
    objs = []
    relink = force or not os.path.exists(LIB)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([src] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        relink = True
    log = []
    for src, stamp, dig, pr in procs:
        out, _ = pr.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
        with open(stamp, "w") as f:
            f.write(dig)
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                   "-ccbin", "/usr/bin/g++", "-cudart", "static"]
        subprocess.run(cmd, check=True)
    if log:
        with open(os.path.join(objdir, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
    return LIB


HOST = os.path.join(ROOT, "gbnns_dim_red_b200", "host")
HOST_BIN = os.path.join(HOST, "bin")


def build_host() -> None:
    """The drop-in C++ drivers (host/final_test.cpp, host/prepare_graph.cpp) against libgbdr.so."""
    os.makedirs(HOST_BIN, exist_ok=True)
    for name in ("final_test", "prepare_graph"):
        src = os.path.join(HOST, name + ".cpp")
        out = os.path.join(HOST_BIN, name)
        deps = [src, os.path.join(HOST, "search_function.h"), os.path.join(ROOT, "include", "gbdr.h"), LIB]
        if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
            continue
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++14", "-Wall", "-I", HOST, "-I", os.path.join(ROOT, "include"), src,
                        "-o", out, "-L", os.path.dirname(LIB), "-lgbdr", "-Wl,-rpath,$ORIGIN/../.."], check=True)


def build_oracle() -> None:
    """gcc/g++ the CPU checkers (tests and baselines only).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force=force, verbose=verbose)
    build_host()
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
