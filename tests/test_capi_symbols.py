"""The C-ABI library loads on a machine without a GPU and exports every symbol include/vofx.h declares;
without a device the entry points fail loudly (no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import pytest

import dune_vof_b200 as pkg
from dune_vof_b200 import lib as vlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    if not os.path.exists(vlib.LIB_PATH):
        pkg.build()
    return vlib.load()


def test_header_symbols_exported(library):
    header = open(os.path.join(ROOT, "include", "vofx.h")).read()
    declared = set(re.findall(r"\b(vofx_[a-z_0-9]+)\s*\(", header))
    declared -= {"vofx_ctx"}
    assert declared == set(vlib.SYMBOLS), declared ^ set(vlib.SYMBOLS)
    for name in declared:
        assert hasattr(library, name), name


def test_fails_loudly_without_device(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = vlib.GridDesc()
    d.dim = 2
    for k in range(3):
        d.n_global[k] = d.n_local[k] = 8 if k < 2 else 1
        d.h[k] = 0.125
    handle = C.c_void_p()
    status = library.vofx_create(C.byref(d), -1, C.byref(handle))
    assert status == 2 and b"no CUDA device" in library.vofx_last_error()
    with pytest.raises(vlib.VofxError):
        from dune_vof_b200.operators import CartesianGrid
        CartesianGrid((8, 8))


def test_argument_validation(library):
    handle = C.c_void_p()
    assert library.vofx_create(None, -1, C.byref(handle)) == 1
    d = vlib.GridDesc()
    d.dim = 4
    assert library.vofx_create(C.byref(d), -1, C.byref(handle)) == 1
