"""Build libroko_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m roko_b200.build            # or  __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libroko_b200.so")
OBJ = os.path.join(CSRC, "build")
SOURCES = ["pack.cu", "front.cu", "front_tc.cu", "proj.cu", "proj_tc3.cu", "proj_h.cu", "rec.cu", "rec_tc.cu", "rec_h.cu", "head.cu", "api.cu",
           "gemm.cu", "train.cu", "train_tc.cu", "rec_bwd.cu", "train_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "roko_b200.h"))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not _stale(obj, [os.path.join(CSRC, src)] + _deps()):
        return obj, ""
    extra = os.environ.get("ROKO_B200_EXTRA_NVCC", "").split()      # tuning experiments only
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr if verbose else ""


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if _stale(LIB, objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
               "-o", LIB] + objs + ["-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
