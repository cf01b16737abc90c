"""ctypes binding of librt_sssp.so (include/rt_sssp.h).  There is no fallback: if the shared library is missing
or an entry point fails, an exception is raised."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("RT_SSSP_LIB") or os.path.join(HERE, "librt_sssp.so")  # RT_SSSP_LIB: kernel-variant A/B runs

I64 = C.c_int64
VP = C.c_void_p
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

# every symbol include/rt_sssp.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "rt_last_error", "rt_version", "rt_device_count", "rt_set_device", "rt_annulus_build", "rt_mesh_sizes",
    "rt_mesh_export", "rt_mesh_from_arrays", "rt_grid3d_build", "rt_grid3d_export", "rt_mesh_free",
    "rt_interp_velocity", "rt_interp_velocity_dev", "rt_interpolate_cells", "rt_nodal_adjacency", "rt_rcm", "rt_mesh_coords_dev", "rt_closest_point", "rt_bfm_solve",
    "rt_bfm_solve_dev", "rt_bfm_solve_multi", "rt_bfm_solve_dual", "rt_dual_velocity", "rt_set_option", "rt_reconstruct_paths", "rt_reconstruct_paths_dev",
    "rt_grid3d_axes", "rt_grid3d_points", "rt_grid3d_connectivity", "rt_closest_point3d", "rt_polardistance3d",
    "rt_reconstruct_paths_guarded", "rt_travel_times", "rt_travel_times_dev", "rt_sssp_nodal", "rt_partition_grid", "rt_bfm_continue",
    "rt_comm_unique_id", "rt_comm_init", "rt_comm_shard", "rt_bfm_solve_sharded", "rt_bfm_solve_sharded_host", "rt_comm_destroy",
]


class RtStats(C.Structure):
    _fields_ = [("sweeps", I64), ("relaxed_edges", I64), ("vertex_updates", I64), ("graph_edges", I64),
                ("kernel_ms", C.c_double), ("relax_ms", C.c_double), ("relax_launches", I64),
                ("total_launches", I64), ("prev_ms", C.c_double), ("screened_edges", I64), ("exact_edges", I64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rt_sssp error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "librt_sssp.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    L.rt_last_error.restype = C.c_char_p
    L.rt_version.restype = C.c_char_p
    L.rt_device_count.argtypes = [C.POINTER(C.c_int)]
    L.rt_set_device.argtypes = [C.c_int]
    L.rt_annulus_build.argtypes = [I64, I64, C.c_double, C.POINTER(VP)]
    L.rt_mesh_sizes.argtypes = [VP, I64P]
    L.rt_mesh_export.argtypes = [VP] + [VP] * 12
    L.rt_mesh_from_arrays.argtypes = [I64, I64, I64P, I64P, I64P, I64P, VP, I64, F64P, F64P, VP, VP, C.POINTER(VP)]
    L.rt_grid3d_build.argtypes = [F64P, F64P, I64P, C.c_int, C.c_int, C.POINTER(VP)]
    L.rt_grid3d_export.argtypes = [VP, F64P, F64P, F64P]
    L.rt_mesh_free.argtypes = [VP]
    L.rt_interp_velocity.argtypes = [F64P, F64P, I64, F64P, I64, C.c_double, F64P]
    L.rt_interpolate_cells.argtypes = [VP, VP, F64P]
    L.rt_nodal_adjacency.argtypes = [VP, VP, VP, VP, I64]
    L.rt_rcm.argtypes = [VP, I64P]
    L.rt_interp_velocity_dev.argtypes = [F64P, F64P, I64, VP, I64, C.c_double, VP]
    L.rt_mesh_coords_dev.argtypes = [VP, C.POINTER(VP), C.POINTER(VP), C.POINTER(VP), C.POINTER(VP)]
    L.rt_closest_point.argtypes = [VP, F64P, F64P, I64, C.c_int, I64P]
    L.rt_bfm_solve.argtypes = [VP, F64P, I64P, I64, C.c_int, VP, VP, C.POINTER(RtStats)]
    L.rt_bfm_solve_dev.argtypes = [VP, VP, I64P, I64, C.c_int, VP, VP, C.POINTER(RtStats)]
    L.rt_bfm_solve_multi.argtypes = [C.POINTER(VP), C.c_int, F64P, I64P, I64, C.c_int, VP, VP, C.POINTER(RtStats)]
    L.rt_bfm_solve_dual.argtypes = [VP, F64P, I64P, I64, VP, VP, C.POINTER(RtStats)]
    L.rt_dual_velocity.argtypes = [F64P, F64P, I64, F64P, I64, C.c_double, F64P]
    L.rt_set_option.argtypes = [VP, C.c_char_p, C.c_double]
    L.rt_reconstruct_paths.argtypes = [I64P, I64, I64, I64P, I64, I64P, VP, I64]
    L.rt_reconstruct_paths_dev.argtypes = [VP, I64, I64, I64P, I64, I64P, VP, I64]
    L.rt_grid3d_axes.argtypes = [VP, VP, VP, VP]
    L.rt_grid3d_points.argtypes = [VP, I64P, I64, VP, VP]
    L.rt_grid3d_connectivity.argtypes = [VP, I64, I64, I64P]
    L.rt_closest_point3d.argtypes = [VP, F64P, F64P, F64P, I64, I64P]
    L.rt_polardistance3d.argtypes = [F64P, F64P, I64, F64P]
    L.rt_reconstruct_paths_guarded.argtypes = [I64P, I64, I64, I64P, I64, I64P, VP, I64]
    L.rt_travel_times.argtypes = [F64P, I64, I64, I64P, I64, F64P]
    L.rt_travel_times_dev.argtypes = [VP, I64, I64, I64P, I64, F64P]
    L.rt_comm_unique_id.argtypes = [C.c_char_p]
    L.rt_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(VP)]
    L.rt_comm_shard.argtypes = [I64, C.c_int, C.c_int, C.POINTER(I64), C.POINTER(I64)]
    L.rt_bfm_solve_sharded.argtypes = [VP, VP, VP, I64P, I64, C.c_int, VP, VP, C.POINTER(RtStats)]
    L.rt_bfm_solve_sharded_host.argtypes = [VP, VP, F64P, I64P, I64, C.c_int, VP, VP, C.POINTER(RtStats)]
    L.rt_comm_destroy.argtypes = [VP]
    L.rt_partition_grid.argtypes = [VP, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    L.rt_bfm_continue.argtypes = [VP, F64P, VP, I64P, I64, F64P, I64P, C.POINTER(RtStats)]
    L.rt_sssp_nodal.argtypes = [VP, F64P, I64, C.c_int, F64P, I64P, C.POINTER(RtStats)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RtError(rc, lib().rt_last_error().decode(errors="replace"))


def ptr(a):
    """numpy array (or None) -> void*"""
    if a is None:
        return None
    return a.ctypes.data_as(VP)
