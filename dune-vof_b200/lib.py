"""ctypes binding of the vofx C ABI (include/vofx.h) and the in-tree build of dune-vof_b200/libvofx.so.

There is no fallback: if the shared library is missing or no CUDA device is present every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libvofx.so")
CSRC = os.path.join(HERE, "csrc")

RECON_YOUNGS, RECON_HF, RECON_SWARTZ = 0, 1, 2
PROB_ROTATING_CIRCLE, PROB_SFLOW, PROB_SLOTTED_CYLINDER, PROB_LINEAR_WALL, PROB_LEVEQUE = 0, 1, 2, 3, 4
TIME_CONST, TIME_COSPI = 0, 1

# every symbol include/vofx.h declares
SYMBOLS = [
    "vofx_last_error", "vofx_create", "vofx_destroy", "vofx_set_stream", "vofx_cells", "vofx_velocity_clear",
    "vofx_velocity_set_component", "vofx_velocity_set_factor", "vofx_velocity_set_faces", "vofx_velocity_builtin", "vofx_face_velocity",
    "vofx_flag", "vofx_reconstruct", "vofx_evolve", "vofx_axpy", "vofx_curvature", "vofx_curvature_swartz", "vofx_curvature_swartz_host", "vofx_step_begin", "vofx_set_incremental_flagging", "vofx_step",
    "vofx_step_state", "vofx_band_list_fetch", "vofx_step_counters", "vofx_step_local", "vofx_step_phase", "vofx_step_finish", "vofx_dt_est_device", "vofx_flag_host",
    "vofx_reconstruct_host", "vofx_evolve_host", "vofx_curvature_host", "vofx_step_host", "vofx_step_host_sparse", "vofx_host_transfer_bytes", "vofx_track_changes", "vofx_changed_fetch", "vofx_color_upload",
    "vofx_step_resident", "vofx_color_download", "vofx_checkpoint_name", "vofx_checkpoint_write", "vofx_checkpoint_read", "vofx_comm_export", "vofx_comm_connect", "vofx_comm_connect_local", "vofx_comm_disconnect",
    "vofx_color_device", "vofx_flags_device", "vofx_halfspaces_device", "vofx_halo_push", "vofx_step_distributed", "vofx_step_distributed_phase", "vofx_band_list_fetch_in_flight",
    "vofx_velocity_set_band_faces", "vofx_diagnostics", "vofx_interface", "vofx_interface_host", "vofx_interface_fetch", "vofx_init", "vofx_init_circle", "vofx_l1error_circle", "vofx_l1error_circle_device",
    "vofx_test_fp64_rate", "vofx_test_arith", "vofx_test_locate", "vofx_test_flux", "vofx_test_interface", "vofx_launch_count", "vofx_profile_enable",
    "vofx_profile_read",
    "vofx_mesh_create", "vofx_mesh_destroy", "vofx_mesh_set_stream", "vofx_mesh_cells", "vofx_mesh_faces", "vofx_mesh_launch_count",
    "vofx_mesh_flag", "vofx_mesh_reconstruct", "vofx_mesh_evolve", "vofx_mesh_flag_host", "vofx_mesh_reconstruct_host",
    "vofx_mesh_evolve_host", "vofx_mesh_reconstruct_swartz", "vofx_mesh_reconstruct_swartz_host", "vofx_mesh_set_reconstruction", "vofx_mesh_color_upload", "vofx_mesh_velocity_upload", "vofx_mesh_run", "vofx_mesh_color_download",
]


class GridDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_global", C.c_int32 * 3), ("origin", C.c_double * 3), ("h", C.c_double * 3),
                ("lo", C.c_int32 * 3), ("n_local", C.c_int32 * 3), ("halo", C.c_int32)]


class MeshDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_cells", C.c_int64), ("corner_offsets", C.c_void_p), ("corners", C.c_void_p),
                ("centers", C.c_void_p), ("volumes", C.c_void_p), ("face_offsets", C.c_void_p), ("face_neighbor", C.c_void_p),
                ("face_corners", C.c_void_p), ("face_centers", C.c_void_p), ("face_unit_normals", C.c_void_p),
                ("face_int_normals", C.c_void_p), ("face_volumes", C.c_void_p), ("stencil_offsets", C.c_void_p),
                ("stencil", C.c_void_p), ("face_corner_offsets", C.c_void_p)]


class StepParams(C.Structure):
    _fields_ = [("reconstruction", C.c_int32), ("max_iterations", C.c_int32), ("eps", C.c_double), ("cfl", C.c_double),
                ("with_curvature", C.c_int32)]


class VofxError(RuntimeError):
    pass


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(ROOT, "include", "vofx.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false ... -> dune-vof_b200/libvofx.so (in-tree)."""
    newest = max(os.path.getmtime(s) for s in sources())
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = ["make", "-j8", "-f", os.path.join(CSRC, "Makefile")]
        if force:
            cmd.append("-B")
        subprocess.check_call(cmd, cwd=ROOT, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def load():
    """Load libvofx.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VofxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(vofx has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, bp, ip = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p   # raw addresses (device or host)
    L.vofx_last_error.restype = C.c_char_p
    L.vofx_create.argtypes = [C.POINTER(GridDesc), C.c_int, C.POINTER(C.c_void_p)]
    L.vofx_destroy.argtypes = [vp]
    L.vofx_destroy.restype = None
    L.vofx_set_stream.argtypes = [vp, vp]
    L.vofx_cells.argtypes = [vp]
    L.vofx_cells.restype = C.c_int64
    L.vofx_launch_count.argtypes = [vp]
    L.vofx_launch_count.restype = C.c_int64
    L.vofx_velocity_clear.argtypes = [vp]
    L.vofx_velocity_set_component.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    L.vofx_velocity_set_factor.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.vofx_velocity_set_faces.argtypes = [vp, dp]
    L.vofx_velocity_builtin.argtypes = [vp, C.c_int]
    L.vofx_face_velocity.argtypes = [vp, C.c_double, C.c_int64, C.c_int, dp]
    L.vofx_flag.argtypes = [vp, dp, C.c_double, bp]
    L.vofx_reconstruct.argtypes = [vp, C.c_int, dp, bp, dp, C.c_int]
    L.vofx_evolve.argtypes = [vp, dp, bp, C.c_double, C.c_double, dp, C.POINTER(C.c_double)]
    L.vofx_axpy.argtypes = [vp, dp, C.c_double, dp]
    L.vofx_curvature.argtypes = [vp, dp, dp, bp, dp, bp]
    L.vofx_step_begin.argtypes = [vp, C.c_double, C.c_double]
    L.vofx_set_incremental_flagging.argtypes = [vp, C.c_int]
    L.vofx_step.argtypes = [vp, C.POINTER(StepParams), dp, bp, dp, dp]
    L.vofx_step_local.argtypes = [vp, C.POINTER(StepParams), dp, bp, dp, dp]
    L.vofx_step_phase.argtypes = [vp, C.POINTER(StepParams), dp, bp, dp, dp, C.c_int]
    L.vofx_step_finish.argtypes = [vp, C.POINTER(StepParams)]
    L.vofx_dt_est_device.argtypes = [vp]
    L.vofx_dt_est_device.restype = C.c_void_p
    L.vofx_step_state.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.vofx_step_counters.argtypes = [vp, C.POINTER(C.c_int64)]
    L.vofx_band_list_fetch.argtypes = [vp, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.vofx_flag_host.argtypes = [vp, dp, C.c_double, bp]
    L.vofx_reconstruct_host.argtypes = [vp, C.c_int, dp, bp, dp, C.c_int]
    L.vofx_evolve_host.argtypes = [vp, dp, bp, C.c_double, C.c_double, dp, C.POINTER(C.c_double)]
    L.vofx_curvature_host.argtypes = [vp, dp, dp, bp, dp, bp]
    L.vofx_step_host.argtypes = [vp, C.POINTER(StepParams), dp, bp]
    L.vofx_step_host_sparse.argtypes = [vp, C.POINTER(StepParams), dp, bp, vp, C.c_int64]
    L.vofx_track_changes.argtypes = [vp, C.c_int]
    L.vofx_changed_fetch.argtypes = [vp, dp, C.c_int64, C.POINTER(C.c_int64)]
    L.vofx_host_transfer_bytes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.vofx_color_upload.argtypes = [vp, dp]
    L.vofx_step_resident.argtypes = [vp, C.POINTER(StepParams)]
    L.vofx_color_download.argtypes = [vp, dp, bp]
    L.vofx_checkpoint_name.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.vofx_checkpoint_write.argtypes = [vp, C.c_char_p, C.c_double]
    L.vofx_checkpoint_read.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
    L.vofx_comm_export.argtypes = [vp, C.c_void_p]
    L.vofx_comm_connect.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_void_p]
    L.vofx_comm_connect_local.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.vofx_comm_disconnect.argtypes = [vp]
    for name in ("vofx_color_device", "vofx_flags_device", "vofx_halfspaces_device"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_void_p
    L.vofx_halo_push.argtypes = [vp]
    L.vofx_step_distributed.argtypes = [vp, C.POINTER(StepParams)]
    L.vofx_step_distributed_phase.argtypes = [vp, C.POINTER(StepParams), C.c_int]
    L.vofx_band_list_fetch_in_flight.argtypes = [vp, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.vofx_velocity_set_band_faces.argtypes = [vp, C.c_int64, dp]
    L.vofx_curvature_swartz.argtypes = [vp, dp, bp, dp]
    L.vofx_curvature_swartz_host.argtypes = [vp, dp, bp, dp]
    L.vofx_interface.argtypes = [vp, dp, bp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.vofx_interface_host.argtypes = [vp, dp, bp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.vofx_interface_fetch.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vofx_init.argtypes = [vp, C.c_int, C.c_double, dp]
    L.vofx_init_circle.argtypes = [vp, C.POINTER(C.c_double), C.c_double, dp]
    L.vofx_l1error_circle.argtypes = [vp, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
    L.vofx_l1error_circle_device.argtypes = [vp, dp, bp, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
    L.vofx_diagnostics.argtypes = [vp, dp, dp, C.POINTER(C.c_double)]
    L.vofx_test_fp64_rate.argtypes = [C.POINTER(C.c_double)]
    L.vofx_test_arith.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    L.vofx_profile_enable.argtypes = [vp, C.c_int]
    L.vofx_profile_read.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.vofx_test_locate.argtypes = [C.c_int, C.c_int64, dp, dp, dp, dp, dp]
    L.vofx_test_flux.argtypes = [C.c_int, C.c_int64, dp, dp, ip, dp, dp, dp]
    L.vofx_test_interface.argtypes = [C.c_int, C.c_int64, dp, dp, dp, ip, dp, dp]
    L.vofx_mesh_create.argtypes = [C.POINTER(MeshDesc), C.c_int, C.POINTER(C.c_void_p)]
    L.vofx_mesh_destroy.argtypes = [vp]
    L.vofx_mesh_destroy.restype = None
    L.vofx_mesh_set_stream.argtypes = [vp, vp]
    for name in ("vofx_mesh_cells", "vofx_mesh_faces", "vofx_mesh_launch_count"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int64
    for suffix in ("", "_host"):
        getattr(L, "vofx_mesh_flag" + suffix).argtypes = [vp, dp, C.c_double, bp]
        getattr(L, "vofx_mesh_reconstruct" + suffix).argtypes = [vp, dp, bp, dp]
        getattr(L, "vofx_mesh_evolve" + suffix).argtypes = [vp, dp, bp, dp, C.c_double, dp, C.POINTER(C.c_double)]
    L.vofx_mesh_reconstruct_swartz.argtypes = [vp, dp, bp, dp, C.c_int]
    L.vofx_mesh_reconstruct_swartz_host.argtypes = [vp, dp, bp, dp, C.c_int]
    L.vofx_mesh_set_reconstruction.argtypes = [vp, C.c_int, C.c_int]
    L.vofx_mesh_color_upload.argtypes = [vp, dp]
    L.vofx_mesh_velocity_upload.argtypes = [vp, dp]
    L.vofx_mesh_run.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.vofx_mesh_color_download.argtypes = [vp, dp, bp]
    _lib = L
    return L


def check(status: int) -> None:
    if status != 0:
        raise VofxError(f"vofx status {status}: {load().vofx_last_error().decode()}")
