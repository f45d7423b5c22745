# 3D unstructured meshes (tetrahedra / hexahedra) on the device: the whole GPU suite
set -x
O=gpurun_out/r2q
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_adaptor.py -x -q -m gpu > $O/pytest_mesh.log 2>&1
tail -15 $O/pytest_mesh.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1
tail -5 $O/pytest.log
