"""The C++ header adaptor include/dune/vofx/*.hh driven like the reference drives its operators (algorithm.hh:57-93):
tests/adaptor/test_adaptor is built on the CPU box by __graft_entry__.build() and travels to the GPU box."""
import os
import subprocess

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def test_adaptor_binary():
    exe = os.path.join(ROOT, "tests", "adaptor", "test_adaptor")
    assert os.path.exists(exe), "tests/adaptor/test_adaptor is missing: run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ADAPTOR TEST PASSED" in r.stdout
    assert "FAIL" not in r.stdout


def test_unstructured_adaptor_binary():
    """include/dune/vofx/unstructured.hh on a mixed triangle/quadrilateral grid view"""
    exe = os.path.join(ROOT, "tests", "adaptor", "test_adaptor_mesh")
    assert os.path.exists(exe), "tests/adaptor/test_adaptor_mesh is missing: run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ADAPTOR MESH TEST PASSED" in r.stdout
    assert "FAIL" not in r.stdout
    ref = exe + "_ref"
    if os.path.exists(ref):   # built only where /root/reference was mounted; travels to the GPU box as a binary
        r = subprocess.run([ref], capture_output=True, text=True, timeout=600)
        print(r.stdout)
        assert r.returncode == 0 and "ADAPTOR MESH TEST PASSED" in r.stdout, r.stdout + r.stderr


def test_unstructured_adaptor_3d_binary():
    """include/dune/vofx/unstructured.hh on a 3D grid view of tetrahedra (makePolyhedron simplex table, upwindPolygon3d prisms)"""
    exe = os.path.join(ROOT, "tests", "adaptor", "test_adaptor_mesh3")
    assert os.path.exists(exe), "tests/adaptor/test_adaptor_mesh3 is missing: run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ADAPTOR MESH3 TEST PASSED" in r.stdout
    assert "FAIL" not in r.stdout
    ref = exe + "_ref"
    if os.path.exists(ref):   # built only where /root/reference was mounted; travels to the GPU box as a binary
        r = subprocess.run([ref], capture_output=True, text=True, timeout=600)
        print(r.stdout)
        assert r.returncode == 0 and "ADAPTOR MESH3 TEST PASSED" in r.stdout, r.stdout + r.stderr


def test_adaptor_reference_containers_binary():
    """the same driver with the reference's own ColorFunction / FlagSet / ReconstructionSet (built where /root/reference is mounted)"""
    ref = os.path.join(ROOT, "tests", "adaptor", "test_adaptor_ref")
    if not os.path.exists(ref):
        pytest.skip("test_adaptor_ref is only built where the reference sources are mounted")
    r = subprocess.run([ref], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "ADAPTOR TEST PASSED" in r.stdout, r.stdout + r.stderr


def test_adaptor_parallel_binary():
    """Dune::VoFx::Algorithm on a parallel grid view: the ranks are threads with a slab grid view each, the device contexts connect
    through vofx_comm_export / vofx_comm_connect; interior values == the oracle on the whole grid, bit for bit"""
    exe = os.path.join(ROOT, "tests", "adaptor", "test_adaptor_parallel")
    assert os.path.exists(exe), "tests/adaptor/test_adaptor_parallel is missing: run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ADAPTOR PARALLEL TEST PASSED" in r.stdout
    assert "FAIL" not in r.stdout
