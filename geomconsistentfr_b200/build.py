"""Builds libgfr_b200.so (all CUDA kernels + the C ABI) in-tree for sm_100a with nvcc.

    python -m geomconsistentfr_b200.build [--force] [-v]

Every .cu is compiled to an object file in parallel (csrc/build/*.o, only the stale ones), then linked.  The .so is
git-ignored but travels to the GPU box with the repo snapshot."""
import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.environ.get("GFR_LIB_PATH") or os.path.join(CSRC, "libgfr_b200.so")      # GFR_LIB_PATH: an alternative build (A/B runs)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
EXTRA = os.environ.get("GFR_NVCC_EXTRA", "").split()       # e.g. -DGFR_CONV_OCC=3 for an A/B build (use with GFR_LIB_PATH)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libgfr_b200.so")
    tag = "" if not (EXTRA or os.environ.get("GFR_LIB_PATH")) else "_" + str(abs(hash((tuple(EXTRA), LIB))) % 10 ** 8)
    objdir = OBJ + tag
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max([os.path.getmtime(h) for h in _headers()] + [0.0])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("".join(e for _, e in results))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
