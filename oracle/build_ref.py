#!/usr/bin/env python
"""Build the REAL reference NMS (utils/nms.pyx) as a CPU checker -> oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vdetlib_b200/`` may import this.

What it does
------------
* reads ``/root/reference/utils/nms.pyx`` where it lies (the reference tree is
  read-only and is never copied into this repository),
* applies the 3-token dtype-alias patch numpy>=1.24 / numpy 2.x needs
  (``np.int_t`` -> ``np.int64_t`` at nms.pyx:25,28,80,83,150,186 and
  ``dtype=np.int`` -> ``dtype=np.int64`` at nms.pyx:29,84,151).  The arithmetic
  is untouched,
* cythonizes + compiles it in a throw-away temp dir (the reference's own
  ``setup.py:2-4`` uses distutils, which Python 3.12 no longer ships),
* copies ONLY the resulting ``cython_nms*.so`` into ``oracle/_ref/``
  (git-ignored; it still travels to the GPU box with the gpurun snapshot).

If ``/root/reference`` is absent (the GPU box) this is a no-op: the prebuilt
``.so`` that travelled with the snapshot is used.
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/utils/nms.pyx"
OUT_DIR = os.path.join(HERE, "_ref")

SETUP_PY = r'''
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy as np
setup(
    name="vdetlib_ref_nms",
    ext_modules=cythonize(
        [Extension("cython_nms", ["nms.pyx"],
                   extra_compile_args=["-Wno-cpp", "-Wno-unused-function"],
                   include_dirs=[np.get_include()])],
        language_level=2),
)
'''


def ref_so_path():
    """Path of the built reference extension, or None."""
    hits = sorted(glob.glob(os.path.join(OUT_DIR, "cython_nms*.so")))
    return hits[0] if hits else None


def build(force=False, verbose=False, keep_c=False):
    if not os.path.isfile(REF_PYX):
        return ref_so_path()          # GPU box: use what travelled
    have = ref_so_path()
    if have and not force and os.path.getmtime(have) >= os.path.getmtime(REF_PYX):
        return have
    src = open(REF_PYX).read()
    patched = re.sub(r"np\.int_t", "np.int64_t", src)
    patched = re.sub(r"dtype=np\.int\)", "dtype=np.int64)", patched)
    tmp = tempfile.mkdtemp(prefix="vdet_ref_build_")
    try:
        with open(os.path.join(tmp, "nms.pyx"), "w") as f:
            f.write(patched)
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(SETUP_PY)
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"],
                           cwd=tmp, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout[-4000:] + r.stderr[-4000:])
        if r.returncode != 0:
            raise RuntimeError("reference nms.pyx failed to build")
        os.makedirs(OUT_DIR, exist_ok=True)
        for old in glob.glob(os.path.join(OUT_DIR, "cython_nms*.so")):
            os.remove(old)
        built = glob.glob(os.path.join(tmp, "cython_nms*.so"))[0]
        shutil.copy2(built, OUT_DIR)
        if keep_c:   # for inspecting the generated arithmetic; never committed
            shutil.copy2(os.path.join(tmp, "nms.c"), "/tmp/vdet_ref_nms.c")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return ref_so_path()


def load():
    """Import the built reference module (``nms``, ``vid_nms``, ``track_det_nms``)."""
    import importlib.util
    path = ref_so_path()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("cython_nms", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True, keep_c="--keep-c" in sys.argv)
    print("reference cython_nms:", p)
