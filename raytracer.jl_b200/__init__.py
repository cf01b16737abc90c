"""raytracer.jl_b200 -- B200-native (sm_100a) drop-in for the shortest-path-method hot path of RayTracer.jl.

The directory name contains a dot (it mirrors the reference's name), so import it through `rt_loader.load()`
at the repository root (registers the package as `raytracer_jl_b200`).
"""
from .api import (R, AdjencyList, adjacency_list, element_degree, SparseAdjencyList, nodal_degree, reorder, sparse_adjacency_list, symrcm, GridPartition, partition_grid, directions, bfm_continue, bfm_multiphase, initial_state, dijkstra, radius_stepping, AbstractSPM, BellmanFordMoore, Dijkstra, RadiusStepping, Point, connectivity, polardistance3D, Grid2D, Grid3D, LinearInterpolation, SparseMatrixCSC, VelProfile, bfm, bfm_gpu, bfm_multi, set_device, bfm3d,
                  closest_point, device_count, dual_velocity, grid, init_annulus, interpolate_inplace, interpolate_velocity, mesh_from_arrays,
                  recontruct_path, travel_times, velocity_profile)
from ._lib import RtError, RtStats, SO_PATH, SYMBOLS, lib

__all__ = [
    "R", "AdjencyList", "adjacency_list", "element_degree", "SparseAdjencyList", "nodal_degree", "reorder", "sparse_adjacency_list", "symrcm", "GridPartition", "partition_grid", "directions", "bfm_continue", "bfm_multiphase", "initial_state", "dijkstra", "radius_stepping", "AbstractSPM", "BellmanFordMoore", "Dijkstra", "RadiusStepping", "Point", "connectivity", "polardistance3D", "Grid2D", "Grid3D", "LinearInterpolation", "SparseMatrixCSC", "VelProfile", "bfm", "bfm_gpu", "bfm_multi", "set_device",
    "bfm3d", "closest_point", "device_count", "dual_velocity", "grid", "init_annulus", "interpolate_inplace", "interpolate_velocity", "mesh_from_arrays",
    "recontruct_path", "travel_times", "velocity_profile", "RtError", "RtStats", "SO_PATH", "SYMBOLS", "lib",
]
