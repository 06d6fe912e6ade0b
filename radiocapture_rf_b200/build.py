"""In-tree build of libb200chan.so for sm_100a (nvcc cross-compiles without a GPU).

Every csrc/*.cu is one translation unit; they are compiled in parallel and linked into the one shared library the
C ABI (include/b200chan.h) lives in."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ_DIR = os.path.join(_HERE, "csrc", "_obj")
OUT = os.path.join(_HERE, "libb200chan.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_src():
    t = 0.0
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for f in os.listdir(root):
            p = os.path.join(root, f)
            if os.path.isfile(p):
                t = max(t, os.path.getmtime(p))
    return t


def build_library(force=False, verbose=False, extra_flags=(), out=None):
    out = out or OUT
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest_src():
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    tag = os.path.splitext(os.path.basename(out))[0]
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, "%s.%s.o" % (tag, os.path.splitext(os.path.basename(src))[0]))
        subprocess.check_call([nvcc] + flags + ["-c", "-o", obj, src])
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, _sources()))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
    return out


def build_experiments(force=True):
    """Tuning build (never shipped, never loaded by default): -DRCB_EXPERIMENTS makes the library read RCB_PFB_VARIANT /
    RCB_FFT_VARIANT so kernel variants can be A/B-ed on the GPU box in one visit.  Select it with RCB_LIBRARY=<path>."""
    return build_library(force=force, extra_flags=["-DRCB_EXPERIMENTS"], out=os.path.join(_HERE, "libb200chan_exp.so"))
