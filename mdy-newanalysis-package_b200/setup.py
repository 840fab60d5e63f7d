"""Builds the Cython shim newanalysis/gfunction*.so in place (links ./newanalysis/libgfn.so).

    make -C csrc && python setup.py build_ext --inplace

Mirrors the reference's Extension('newanalysis.gfunction', ...) entry (/root/reference/setup.py:73-78)
minus OpenMP: the pair loop no longer runs on the host.
"""
import os

import numpy
from Cython.Build import cythonize
from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

ext = Extension(
    "newanalysis.gfunction",
    [os.path.join("newanalysis", "gfunction.pyx")],
    language="c++",
    include_dirs=[numpy.get_include(), os.path.join(ROOT, "include")],
    library_dirs=[os.path.join(HERE, "newanalysis")],
    libraries=["gfn"],
    runtime_library_dirs=["$ORIGIN"],
    extra_compile_args=["-O2"],
    define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
)

setup(name="newanalysis_b200", packages=["newanalysis"],
      ext_modules=cythonize([ext], language_level=3, quiet=True, build_dir="build/cython"))
