"""Compile rvc_maker_b200/csrc/*.cu into the in-tree C-ABI library ``rvc_maker_b200/librvc_b200.so`` (sm_100a only).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librvc_b200.so")
STAMP = LIB + ".stamp"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v", "-lcuda"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "rvc_b200.h")]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, trace=False, defines=(), suffix=""):
    """trace=True builds the debugging twin ``librvc_b200_trace.so`` (-DRVC_PAIR_TRACE: timeline records in the fused ResBlock kernel,
    tools/pair_trace.py selects it through RVC_B200_LIB); the product library never carries the trace points."""
    if trace or suffix:
        # experiment builds (tools/ only): extra -D switches, a library name of their own
        name = "librvc_b200" + ("_trace" if trace else "") + (f"_{suffix}" if suffix else "") + ".so"
        lib, stamp, tmpname = os.path.join(HERE, name), os.path.join(HERE, name + ".stamp"), "build" + ("_trace" if trace else "") + (f"_{suffix}" if suffix else "")
        extra = (["-DRVC_PAIR_TRACE"] if trace else []) + [f"-D{d}" for d in defines]
    else:
        if defines:
            raise ValueError("the product library takes no extra defines; give the experiment build a suffix")
        lib, stamp, extra, tmpname = LIB, STAMP, [], "build"
    dig = _digest() + "+" + " ".join(extra)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return lib
    objs = []
    nvcc = _nvcc()
    tmp = os.path.join(HERE, tmpname)
    os.makedirs(tmp, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"] + extra + ["-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    link = [nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcuda"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(r.stdout)
    if r.returncode != 0:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("link failed")
    with open(os.path.join(tmp, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    _defs = [sys.argv[i + 1] for i, a in enumerate(sys.argv[:-1]) if a == "--define"]
    _suf = sys.argv[sys.argv.index("--suffix") + 1] if "--suffix" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv, defines=_defs, suffix=_suf))
