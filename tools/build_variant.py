#!/usr/bin/env python
"""Build a kernel VARIANT of libvdet_b200.so next to the product library (CPU only: nvcc cross-compiles).

    python tools/build_variant.py NAME [-DMACRO=1 ...]      ->  vdetlib_b200/variants/libvdet_b200_NAME.so
    VDET_B200_LIB=vdetlib_b200/variants/libvdet_b200_NAME.so python tools/nms_time.py     (on the GPU box)

Experiments live behind compile-time macros in csrc/ (all off in the product build); a variant goes through the
same parity tests (`VDET_B200_LIB=... pytest -m gpu`) and timing tools before its macro becomes the default."""
import concurrent.futures
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdetlib_b200 import build as B          # noqa: E402


def main():
    if len(sys.argv) < 2 or sys.argv[1].startswith("-"):
        sys.exit(__doc__)
    name, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "vdetlib_b200", "variants")
    obj_dir = os.path.join(out_dir, "_obj_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    srcs = B.sources()

    def one(src):
        obj = os.path.join(obj_dir, os.path.splitext(os.path.basename(src))[0] + ".o")
        r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + defs + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        return obj
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    lib = os.path.join(out_dir, "libvdet_b200_%s.so" % name)
    r = subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr)
    print(lib)


if __name__ == "__main__":
    main()
