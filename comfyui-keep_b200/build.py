"""Build libkeep_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C-ABI)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkeep_b200.so")
OBJ = os.path.join(HERE, "build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


# extra -D switches for A/B experiments on the GPU box, e.g. KEEP_NVCC_EXTRA="-DKEEP_PDL_TRIGGER_MAX=0" (run build.py --force with it)
EXTRA = os.environ.get("KEEP_NVCC_EXTRA", "").split()


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (needed to build libkeep_b200.so for sm_100a)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "keep_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libkeep_b200.so next to this file."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".h"))
    hdr_t = max(hdr_t, os.path.getmtime(os.path.join(HERE, "..", "include", "keep_b200.h")))

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + EXTRA + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp"] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    # libcuda is deliberately NOT linked: the library must load on a driver-less box (CPU test tier);
    # driver entry points (cuTensorMapEncodeTiled) are fetched at run time via cudaGetDriverEntryPoint.
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
