"""In-tree build of libctcb200.so (hand-written sm_100a kernels behind a C ABI).

`nvcc` cross-compiles without a GPU, so this runs on the CPU-only build box; the resulting shared
object is git-ignored but travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "libctcb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-DCTCB200_BUILD",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libctcb200.so")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    if os.path.isdir(inc):
        hdrs += [os.path.join(inc, f) for f in os.listdir(inc)]
    return max([os.path.getmtime(h) for h in hdrs] + [0.0])


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > lib_m for s in sources()) or _deps_mtime() > lib_m


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libctcb200.so. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _deps_mtime()
    inc = os.path.join(os.path.dirname(HERE), "include")

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_m):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + ["-I", CSRC, "-I", inc, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    if "--tags" in sys.argv:   # debugging build: stalled mbarrier waits of the pipelined kernel report their role and step
        NVCC_FLAGS.append("-DCTCB200_WAIT_TAGS")
    if "--trace" in sys.argv:  # profiling build: in-kernel phase stamps of the recurrent kernels (CTCB200_LSTM_TRACE=1 at run time)
        NVCC_FLAGS.append("-DCTCB200_TRACE")
    print(build(force="--force" in sys.argv or "--tags" in sys.argv or "--trace" in sys.argv, verbose="-v" in sys.argv))
