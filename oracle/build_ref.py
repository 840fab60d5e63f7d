#!/usr/bin/env python3
"""Build the UNMODIFIED reference gfunction module into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE - never imported by the product path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may load what this script produces.

Recipe: the reference hot path lives in ONE Cython source,
/root/reference/newanalysis/gfunction/gfunction.pyx.  It is cythonized where it lies: the .pyx is
*read* from /root/reference, the generated C++ and all build temporaries go to a scratch directory
under /tmp, and the only thing written into this repository is the built extension module
oracle/_ref/gfunction*.so (git-ignored, NOT gpurun-ignored, so it travels to the GPU box).
No reference source text is copied into the repository.

Flags: g++ -O2 -fopenmp, no -march / -ffast-math / -ffp-contract=fast, so the binary contains no
FMA contraction and is the bit-exact definition of the reference arithmetic (gfunction.pyx:110-165).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
Exit code 0 and prints the path on success; exit code 3 if the reference tree is absent
(GPU box: the prebuilt .so that travelled with the snapshot is used as-is).
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SETUP = r"""
import numpy, sys
from setuptools import setup, Extension
from Cython.Build import cythonize
ext = Extension('gfunction', [sys.argv.pop()], language='c++',
                extra_compile_args=['-fopenmp', '-O2', '-ffp-contract=off'], extra_link_args=['-fopenmp'],
                include_dirs=[numpy.get_include()],
                define_macros=[('NPY_NO_DEPRECATED_API', 'NPY_1_7_API_VERSION')])
setup(name='gfunction_ref', ext_modules=cythonize([ext], language_level=3, quiet=True))
"""


def built_module():
    hits = sorted(glob.glob(os.path.join(OUT, "gfunction*.so")))
    return hits[0] if hits else None


def build(reference="/root/reference", force=False):
    pyx = os.path.join(reference, "newanalysis", "gfunction", "gfunction.pyx")
    have = built_module()
    if have and not force:
        return have
    if not os.path.isfile(pyx):
        return None
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="gfn_ref_build_") as tmp:
        # Cython writes its generated .cpp next to the .pyx it is given; /root/reference is
        # read-only, so the build sees the source through a scratch copy that never enters the repo.
        scratch_pyx = os.path.join(tmp, "gfunction.pyx")
        shutil.copyfile(pyx, scratch_pyx)
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP)
        env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++", LDSHARED="/usr/bin/g++ -shared")
        cmd = [sys.executable, "setup.py", "build_ext", "--build-lib", OUT, "--build-temp",
               os.path.join(tmp, "bt"), scratch_pyx]
        res = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout)
            raise RuntimeError("reference gfunction.pyx failed to build")
    return built_module()


SETUP_HELPERS = r"""
import numpy, sys
from setuptools import setup, Extension
from Cython.Build import cythonize
ext = Extension('helpers', ['helpers.pyx', 'BertholdHorn.cpp'], language='c++',
                extra_compile_args=['-fopenmp', '-O2', '-ffp-contract=off'], extra_link_args=['-fopenmp'],
                include_dirs=[numpy.get_include(), '.'],
                define_macros=[('NPY_NO_DEPRECATED_API', 'NPY_1_7_API_VERSION')])
setup(name='helpers_ref', ext_modules=cythonize([ext], language_level=3, quiet=True))
"""


def build_helpers(reference="/root/reference", force=False):
    """The reference helpers module (comByResidue / dipByResidue, next-row N1): helpers.pyx + BertholdHorn.cpp,
    read from /root/reference, built in /tmp, only the .so lands in oracle/_ref/."""
    hits = sorted(glob.glob(os.path.join(OUT, "helpers*.so")))
    if hits and not force:
        return hits[0]
    src = os.path.join(reference, "newanalysis", "helpers")
    if not os.path.isfile(os.path.join(src, "helpers.pyx")):
        return None
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="gfn_ref_helpers_") as tmp:
        for f in ("helpers.pyx", "BertholdHorn.cpp", "BertholdHorn.h"):
            shutil.copyfile(os.path.join(src, f), os.path.join(tmp, f))
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP_HELPERS)
        env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++", LDSHARED="/usr/bin/g++ -shared")
        res = subprocess.run([sys.executable, "setup.py", "build_ext", "--build-lib", OUT, "--build-temp", os.path.join(tmp, "bt")],
                             cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout[-3000:])
            raise RuntimeError("reference helpers.pyx failed to build")
    hits = sorted(glob.glob(os.path.join(OUT, "helpers*.so")))
    return hits[0] if hits else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    so = build(a.reference, a.force)
    if so is None:
        print("reference tree not present and no prebuilt oracle/_ref module", file=sys.stderr)
        sys.exit(3)
    print(so)
    print(build_helpers(a.reference, a.force))


if __name__ == "__main__":
    main()
