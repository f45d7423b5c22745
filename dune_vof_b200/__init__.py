"""Importable name of the package that lives in ``dune-vof_b200/`` (a hyphen cannot be imported).

``import dune_vof_b200`` resolves its sub-modules from the directory ``dune-vof_b200/`` next to this one:
``dune_vof_b200.lib`` (ctypes binding of the C ABI, include/vofx.h), ``dune_vof_b200.operators`` (host-side mirror
of dune-vof's operator API) and ``dune_vof_b200.decomposition`` (slab decomposition over the GPUs of one node).
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
PACKAGE_DIR = _os.path.join(_os.path.dirname(_HERE), "dune-vof_b200")
__path__.append(PACKAGE_DIR)

from .lib import build, load, LIB_PATH  # noqa: E402,F401
