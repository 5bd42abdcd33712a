"""In-tree build of libnlc_b200.so (sm_100a only) with plain nvcc.

The shared object is written next to this file so it travels with the repository snapshot to the GPU
box; nothing is JIT-compiled at import time.  `python -m nlc_b200.build` or `__graft_entry__.build()`.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnlc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libnlc_b200.so cannot be built")


def _digest(path, extra):
    h = hashlib.sha1()
    h.update(" ".join(extra).encode())
    for dep in sorted(os.listdir(CSRC)):
        if dep.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, dep), "rb") as f:
                h.update(f.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "nlc_b200.h"), "rb") as f:
        h.update(f.read())
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _compile_one(nvcc, src, verbose):
    name = os.path.basename(src)[:-3]
    obj = os.path.join(OBJDIR, name + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src, NVCC_FLAGS)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (name, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj


def build(verbose=False, force=False):
    """Compile every csrc/*.cu for sm_100a and link libnlc_b200.so. Returns the library path."""
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs + [
            "-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
