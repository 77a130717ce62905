"""In-tree build of libvdet_b200.so (hand-written sm_100a CUDA behind a C ABI).

    python -m vdetlib_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting ``vdetlib_b200/libvdet_b200.so`` is
git-ignored but travels to the GPU box with the gpurun snapshot.  Flags that matter:

  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX fallback for other parts
  -fmad=false                               no FMA contraction: every float op rounds once,
                                            as the reference's C / NumPy arithmetic does
  -lineinfo                                 ncu source pages map to these files
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libvdet_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def sources():
    """CUDA sources + plain host sources (nvcc hands .cpp files to the host compiler)."""
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _deps_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources()) or _deps_mtime() > t


def _compile_one(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(os.path.basename(src))[0] + ".o")
    if os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), _deps_mtime()):
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libvdet_b200.so.  Returns its path."""
    if not force and not needs_build():
        return LIB
    srcs = sources()
    if not srcs:
        raise RuntimeError("no CUDA sources under %s" % CSRC)
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for o in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(o)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
