"""ctypes front of the CPU oracle (oracle/rt_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
the product package (raytracer.jl_b200) never does.  Parity status: "parity unpinned" (no golden vectors exist
in the reference and Julia is unavailable) -- see the header of rt_oracle.cpp for what pins the oracle instead.

All ids are 1-based int64 as in the Julia reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

I64 = C.c_int64
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
I8P = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "librt_oracle.so")
    src = os.path.join(_HERE, "rt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librt_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ora_annulus_build.restype = C.c_void_p
        L.ora_annulus_build.argtypes = [I64, I64, C.c_double]
        L.ora_mesh_sizes.argtypes = [C.c_void_p, I64P]
        L.ora_mesh_export.argtypes = [C.c_void_p, F64P, F64P, F64P, F64P, I64P, I64P, I64P, I64P, I64P, I64P,
                                      I64P, I8P]
        L.ora_mesh_free.argtypes = [C.c_void_p]
        L.ora_interp_velocity.argtypes = [F64P, F64P, I64, F64P, I64, C.c_double, F64P]
        L.ora_closest_point.restype = I64
        L.ora_closest_point.argtypes = [F64P, F64P, I64, C.c_double, C.c_double]
        L.ora_bfm.argtypes = [I64, I64, I64P, I64P, I64P, I64P, I64P, I64, F64P, F64P, F64P, I64, C.c_int, I64,
                              F64P, I64P, I64P]
        L.ora_dijkstra.argtypes = [I64, I64, I64P, I64P, I64P, I64P, I64P, I64, F64P, F64P, F64P, I64, F64P]
        L.ora_reconstruct_path.restype = I64
        L.ora_reconstruct_path.argtypes = [I64P, I64, I64, I64, I64P, I64]
        L.ora_grid3d_coords.argtypes = [F64P, F64P, I64P, C.c_int, F64P, F64P, F64P]
        L.ora_bfm3d.argtypes = [I64P, C.c_int, F64P, F64P, F64P, F64P, I64, C.c_int, I64, F64P, I64P, I64P]
        L.ora_bfm_f32.argtypes = [I64, I64, I64P, I64P, I64P, I64P, I64P, I64, F64P, F64P, F64P, I64, C.c_int, F64P,
                                  I64P, I64P]
        L.ora_bfm3d_f32.argtypes = [I64P, C.c_int, F64P, F64P, F64P, F64P, I64, C.c_int, F64P, I64P, I64P]
        L.ora_dijkstra3d.argtypes = [I64P, C.c_int, F64P, F64P, F64P, F64P, I64, F64P]
        L.ora_interpolate_cells.argtypes = [I64, I64P, I64P, I8P, F64P, F64P, F64P]
        L.ora_nodal_adjacency.argtypes = [I64, I64, I64P, I64P, I64P, I64P, C.c_void_p]
        L.ora_bfm_dual.argtypes = [I64, I64, I64P, I64P, I64P, I64P, I64P, I64, F64P, F64P, F64P, F64P, F64P, I64, C.c_int,
                                   F64P, I64P, I64P]
        L.ora_dual_velocity.argtypes = [F64P, F64P, I64, F64P, I64, C.c_double, F64P]
        L.ora_dijkstra_nodal.argtypes = [I64, I64P, I64P, F64P, F64P, F64P, I64, F64P, I64P]
        L.ora_radius_stepping_nodal.restype = I64
        L.ora_radius_stepping_nodal.argtypes = [I64, I64P, I64P, F64P, F64P, F64P, I64, F64P, I64P]
        L.ora_partition_grid.argtypes = [F64P, I64, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
        L.ora_bfm_continue.argtypes = [I64, I64, I64P, I64P, I64P, I64P, I64P, I64, F64P, F64P, F64P, C.c_void_p, I64P,
                                       I64, C.c_int, F64P, I64P, I64P]
        L.ora_set_weight3d.argtypes = [C.c_int]
        L.ora_window3d.argtypes = [C.c_int]
        L.ora_nodal_incidence3d.argtypes = [I64P, C.c_int, I64P, C.c_void_p]
        L.ora_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


class Annulus:
    """Arrays of `gr, G, halo = init_annulus(ntheta, nr; spacing)` (src/GridAnnulus.jl:57-70)."""

    def __init__(self, ntheta, nr, spacing):
        L = lib()
        h = L.ora_annulus_build(int(ntheta), int(nr), float(spacing))
        sz = np.zeros(8, np.int64)
        L.ora_mesh_sizes(h, sz)
        n, nel, se, nnz, hr, sn, nth, nrr = (int(v) for v in sz)
        self.n, self.nel, self.ntheta, self.nr = n, nel, nth, nrr
        self.x, self.z, self.theta, self.r = (np.zeros(n) for _ in range(4))
        self.e2n_off = np.zeros(nel + 1, np.int64)
        self.e2n_idx = np.zeros(se, np.int64)
        self.G_colptr = np.zeros(n + 1, np.int64)
        self.G_rowval = np.zeros(nnz, np.int64)
        self.halo = np.zeros(max(2 * hr, 1), np.int64)
        self.nbr_off = np.zeros(nel + 1, np.int64)
        self.nbr_idx = np.zeros(max(sn, 1), np.int64)
        self.el_type = np.zeros(nel, np.int8)
        L.ora_mesh_export(h, self.x, self.z, self.theta, self.r, self.e2n_off, self.e2n_idx, self.G_colptr,
                          self.G_rowval, self.halo, self.nbr_off, self.nbr_idx, self.el_type)
        L.ora_mesh_free(h)
        self.halo = self.halo[:2 * hr]
        self.nbr_idx = self.nbr_idx[:sn]
        self.halo_rows = hr

    def halo_matrix(self):
        """(2H, 2) like the Julia Matrix{Int64} (column-major storage)."""
        return self.halo.reshape(2, self.halo_rows).T


def interp_velocity(knots_r, knots_v, r, buffer=-1.0):
    r = np.ascontiguousarray(r, np.float64)
    out = np.zeros_like(r)
    rc = lib().ora_interp_velocity(np.ascontiguousarray(knots_r, np.float64),
                                   np.ascontiguousarray(knots_v, np.float64), len(knots_r), r, len(r),
                                   float(buffer), out)
    if rc:
        raise ValueError("interpolation point outside the knots (BoundsError in the reference)")
    return out


def closest_point(a, b, pa, pb):
    return int(lib().ora_closest_point(np.ascontiguousarray(a), np.ascontiguousarray(b), len(a), float(pa),
                                       float(pb)))


def bfm(mesh, U, source, nthreads=None, max_sweeps=0):
    """bfm(G, halo, source, gr, U) (src/SSSP/bfm.jl:1-52). Returns dist, prev (1-based, 0 = never set), stats."""
    n = mesh.n
    dist = np.zeros(n)
    prev = np.zeros(n, np.int64)
    stats = np.zeros(4, np.int64)
    halo = mesh.halo if mesh.halo_rows else np.zeros(1, np.int64)
    rc = lib().ora_bfm(n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, mesh.G_colptr, mesh.G_rowval, halo,
                       mesh.halo_rows, mesh.x, mesh.z, np.ascontiguousarray(U, np.float64), int(source),
                       _nt(nthreads), int(max_sweeps), dist, prev, stats)
    if rc:
        raise ValueError("bad source")
    return dist, prev, dict(sweeps=int(stats[0]), relaxed_edges=int(stats[1]), vertex_updates=int(stats[2]),
                            graph_edges=int(stats[3]))


def bfm_f32(mesh, U, source, nthreads=None):
    """Float32 comparison path (src/SSSP/bfm_gpu.jl:170-205, 487-526): x, z, U cast to Float32, travel times relaxed
    in Float32.  Returns the Float32 travel times widened to float64, prev, stats."""
    n = mesh.n
    dist = np.zeros(n)
    prev = np.zeros(n, np.int64)
    stats = np.zeros(4, np.int64)
    halo = mesh.halo if mesh.halo_rows else np.zeros(1, np.int64)
    rc = lib().ora_bfm_f32(n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, mesh.G_colptr, mesh.G_rowval, halo,
                           mesh.halo_rows, mesh.x, mesh.z, np.ascontiguousarray(U, np.float64), int(source),
                           _nt(nthreads), dist, prev, stats)
    if rc:
        raise ValueError("bad source")
    return dist, prev, dict(sweeps=int(stats[0]), relaxed_edges=int(stats[1]), vertex_updates=int(stats[2]),
                            graph_edges=int(stats[3]))


def dijkstra(mesh, U, source):
    dist = np.zeros(mesh.n)
    halo = mesh.halo if mesh.halo_rows else np.zeros(1, np.int64)
    lib().ora_dijkstra(mesh.n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, mesh.G_colptr, mesh.G_rowval, halo,
                       mesh.halo_rows, mesh.x, mesh.z, np.ascontiguousarray(U, np.float64), int(source), dist)
    return dist


def reconstruct_path(prev, source, receiver):
    prev = np.ascontiguousarray(prev, np.int64)
    cap = 1024
    while True:
        out = np.zeros(cap, np.int64)
        ln = int(lib().ora_reconstruct_path(prev, len(prev), int(source), int(receiver), out, cap))
        if ln == -1:
            raise RuntimeError("path does not reach the source")
        if ln < 0:
            cap = -ln
            continue
        return out[:ln].copy()


def grid3d_coords(c0, c1, nn, coord_system=0):
    nn = np.asarray(nn, np.int64)
    n = int(np.prod(nn))
    X, Y, Z = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().ora_grid3d_coords(np.asarray(c0, np.float64), np.asarray(c1, np.float64), nn, int(coord_system), X, Y,
                            Z)
    return X, Y, Z


def bfm3d(nn, star_levels, X, Y, Z, U, source, nthreads=None, max_sweeps=0):
    nn = np.asarray(nn, np.int64)
    n = int(np.prod(nn))
    dist = np.zeros(n)
    prev = np.zeros(n, np.int64)
    stats = np.zeros(4, np.int64)
    rc = lib().ora_bfm3d(nn, int(star_levels), X, Y, Z, np.ascontiguousarray(U, np.float64), int(source),
                         _nt(nthreads), int(max_sweeps), dist, prev, stats)
    if rc:
        raise ValueError("bad source")
    return dist, prev, dict(sweeps=int(stats[0]), relaxed_edges=int(stats[1]), vertex_updates=int(stats[2]),
                            graph_edges=int(stats[3]))


def bfm3d_f32(nn, star_levels, X, Y, Z, U, source, nthreads=None):
    """3-D solve with coordinates, U and travel times in Float32 (benchmarks/cpu.jl:9-13 runs the grid in Float32)."""
    nn = np.asarray(nn, np.int64)
    n = int(np.prod(nn))
    dist = np.zeros(n)
    prev = np.zeros(n, np.int64)
    stats = np.zeros(4, np.int64)
    rc = lib().ora_bfm3d_f32(nn, int(star_levels), X, Y, Z, np.ascontiguousarray(U, np.float64), int(source),
                             _nt(nthreads), dist, prev, stats)
    if rc:
        raise ValueError("bad source")
    return dist, prev, dict(sweeps=int(stats[0]), relaxed_edges=int(stats[1]), vertex_updates=int(stats[2]),
                            graph_edges=int(stats[3]))


def dijkstra3d(nn, star_levels, X, Y, Z, U, source):
    nn = np.asarray(nn, np.int64)
    dist = np.zeros(int(np.prod(nn)))
    lib().ora_dijkstra3d(nn, int(star_levels), X, Y, Z, np.ascontiguousarray(U, np.float64), int(source), dist)
    return dist


def dijkstra_nodal(mesh, U, source, adjacency=None):
    """dijkstra(G, source, gr, U) src/SSSP/dijkstra.jl:68-136 on nodal_incidence(gr) (star-0) -> (dist, prev)."""
    deg, off, lst = adjacency if adjacency is not None else nodal_adjacency(mesh)
    dist = np.zeros(mesh.n)
    prev = np.zeros(mesh.n, np.int64)
    if lib().ora_dijkstra_nodal(mesh.n, off, lst, mesh.x, mesh.z, np.ascontiguousarray(U, np.float64), int(source), dist,
                                prev):
        raise ValueError("bad source")
    return dist, prev


def radius_stepping_nodal(mesh, U, source, adjacency=None):
    """radius_stepping(Gsp, source, gr, U) src/SSSP/radius_stepping.jl:7-46 on the same graph -> (dist, prev, it)."""
    deg, off, lst = adjacency if adjacency is not None else nodal_adjacency(mesh)
    dist = np.zeros(mesh.n)
    prev = np.zeros(mesh.n, np.int64)
    it = lib().ora_radius_stepping_nodal(mesh.n, off, lst, mesh.x, mesh.z, np.ascontiguousarray(U, np.float64),
                                         int(source), dist, prev)
    if it < 0:
        raise ValueError("bad source")
    return dist, prev, int(it)


def partition_grid(r):
    """partition_grid(gr) src/topology/topology.jl:183-206 -> int32 ids: k > 0 = "Layer_k", -k = "Boundary_k"."""
    r = np.ascontiguousarray(r, np.float64)
    out = np.zeros(len(r), np.int32)
    lib().ora_partition_grid(r, len(r), out)
    return out


def bfm_continue(mesh, U, allowed, seeds, dist, prev, nthreads=None):
    """Restricted continuation (inner loop of bfm_multiphase, src/SSSP/bfm_multiphase.jl:118-150): returns new (dist,
    prev, stats); the inputs are not modified."""
    dist = np.array(dist, np.float64)
    prev = np.array(prev, np.int64)
    seeds = np.ascontiguousarray(np.atleast_1d(seeds), np.int64)
    al = None if allowed is None else np.ascontiguousarray(allowed, np.uint8)
    stats = np.zeros(4, np.int64)
    halo = mesh.halo if mesh.halo_rows else np.zeros(1, np.int64)
    rc = lib().ora_bfm_continue(mesh.n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, mesh.G_colptr, mesh.G_rowval, halo,
                                mesh.halo_rows, mesh.x, mesh.z, np.ascontiguousarray(U, np.float64),
                                None if al is None else al.ctypes.data_as(C.c_void_p), seeds, len(seeds), _nt(nthreads),
                                dist, prev, stats)
    if rc:
        raise ValueError("bad seed")
    return dist, prev, dict(sweeps=int(stats[0]), relaxed_edges=int(stats[1]), vertex_updates=int(stats[2]))


def set_weight3d(mode):
    """0: src/SSSP/weights.jl:20 (default); 1: the expression inside BFM/foo! src/Dijsktra.jl:388."""
    lib().ora_set_weight3d(int(mode))


def window3d(star_levels):
    """Half width of the implicit window equivalent to nodal_incidence(gr; neighbour_levels) (2^L)."""
    return int(lib().ora_window3d(int(star_levels)))


def nodal_incidence3d(nn, star_levels):
    """nodal_incidence(gr; neighbour_levels) src/StructuredGrid.jl:177-223 built literally (hexes -> Dict of Sets ->
    expansion rounds).  Returns (off[n+1], list ascending per node, 1-based).  Small grids only."""
    nn = np.asarray(nn, np.int64)
    n = int(np.prod(nn))
    off = np.zeros(n + 1, np.int64)
    lib().ora_nodal_incidence3d(nn, int(star_levels), off, None)
    lst = np.zeros(int(off[-1]), np.int64)
    lib().ora_nodal_incidence3d(nn, int(star_levels), off, lst.ctypes.data_as(C.c_void_p))
    return off, lst


def grid3d_axes(c0, c1, nn):
    """collect(LinRange(c0[d], c1[d], nn[d])) src/StructuredGrid.jl:38-40 (Julia's lerpi: t = j/d; (1-t)*a + t*b)."""
    out = []
    for d in range(3):
        m = int(nn[d])
        if m == 1:
            out.append(np.array([float(c0[d])]))
            continue
        t = np.arange(m, dtype=np.float64) / float(m - 1)
        out.append((1.0 - t) * float(c0[d]) + t * float(c1[d]))
    return out


def cartesian_index3d(nn, I):
    """CartesianIndex(gr, I) src/StructuredGrid.jl:90-96 (1-based)."""
    nx, nxny = int(nn[0]), int(nn[0]) * int(nn[1])
    i = (I - 1) % nx + 1
    k = -(-I // nxny)
    j = -(-(I - nxny * (k - 1)) // nx)
    return i, j, k


def connectivity3d(nn):
    """connectivity(gr) src/StructuredGrid.jl:121-168 with cornerindex_ijk :106-112 -> (nel, 8) 1-based."""
    nx, ny, nz = (int(v) for v in nn)
    ex, ey, ez = nx - 1, ny - 1, nz - 1
    iel = np.arange(1, ex * ey * ez + 1, dtype=np.int64)
    i = (iel - 1) % ex + 1
    j = (-(-iel // ex) - 1) % ey + 1
    k = -(-iel // (ex * ey))
    idx = i + (j - 1) * nx + (k - 1) * nx * ny
    nxny = nx * ny
    return np.stack([idx, idx + 1, idx + 1 + nx, idx + nx, idx + nxny, idx + nxny + 1, idx + nxny + 1 + nx,
                     idx + nxny + nx], axis=1)


def closest_point3d(axes, p):
    """closest_point(gr, x, y, z) src/StructuredGrid.jl:257-270: first linear index (x fastest) of the strict minimum
    of distance3D(gr[i], p) on the raw axis coordinates; -1 if no distance is < Inf."""
    ax, ay, az = axes
    d = np.sqrt((((ax - p[0]) ** 2)[None, None, :] + ((ay - p[1]) ** 2)[None, :, None]) + ((az - p[2]) ** 2)[:, None, None])
    d = d.reshape(-1)
    ok = ~np.isnan(d) & (d < np.inf)
    if not ok.any():
        return -1
    return int(np.flatnonzero(ok & (d == d[ok].min()))[0]) + 1


def reconstruct_path_struct(prev, source, receiver):
    """recontruct_path(D, source, receiver) src/SSSP/ssspm.jl:14-28 (the struct method), literally."""
    path = [int(receiver)]
    ip = int(prev[receiver - 1])
    while ip not in path:
        if ip < 1:
            raise IndexError("BoundsError: prev[%d]" % ip)
        path.append(ip)
        ip = int(prev[ip - 1])
    path.append(int(source))
    return np.asarray(path, np.int64)


def interpolate_cells(mesh, V):
    """interpolate!(V, gr) src/Interpolations/interpolation.jl:5-18 (in place on a copy)."""
    V = np.array(V, np.float64)
    lib().ora_interpolate_cells(mesh.nel, mesh.e2n_off, mesh.e2n_idx, mesh.el_type, mesh.theta, mesh.r, V)
    return V


def nodal_adjacency(mesh):
    """nodal_incidence(gr::Grid2D) src/GridAnnulus.jl:763-804 -> (deg, off, list ascending per node, 1-based)."""
    deg = np.zeros(mesh.n, np.int64)
    off = np.zeros(mesh.n + 1, np.int64)
    lib().ora_nodal_adjacency(mesh.n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, deg, off, None)
    lst = np.zeros(int(off[-1]), np.int64)
    lib().ora_nodal_adjacency(mesh.n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, deg, off, lst.ctypes.data_as(C.c_void_p))
    return deg, off, lst


def dual_velocity(knots_r, knots_v, r, buffer=1.0):
    """dual_velocity(r, interpolant; buffer) src/utils.jl:51-66 -> (n, 2) array like the Julia Matrix."""
    r = np.ascontiguousarray(r, np.float64)
    out = np.zeros(2 * len(r))
    if lib().ora_dual_velocity(np.ascontiguousarray(knots_r, np.float64), np.ascontiguousarray(knots_v, np.float64),
                               len(knots_r), r, len(r), float(buffer), out):
        raise ValueError("interpolation point outside the knots")
    return out.reshape(2, len(r)).T.copy()


def bfm_dual(mesh, U2, source, nthreads=None):
    """bfm(G, halo, source, gr, U::Matrix) -> _relax!(..., U::Matrix) src/SSSP/bfm.jl:113-159."""
    n = mesh.n
    U2 = np.asarray(U2, np.float64)
    u1, u2 = np.ascontiguousarray(U2[:, 0]), np.ascontiguousarray(U2[:, 1])
    dist = np.zeros(n)
    prev = np.zeros(n, np.int64)
    stats = np.zeros(4, np.int64)
    halo = mesh.halo if mesh.halo_rows else np.zeros(1, np.int64)
    rc = lib().ora_bfm_dual(n, mesh.nel, mesh.e2n_off, mesh.e2n_idx, mesh.G_colptr, mesh.G_rowval, halo,
                            mesh.halo_rows, mesh.x, mesh.z, mesh.r, u1, u2, int(source), _nt(nthreads), dist, prev, stats)
    if rc:
        raise ValueError("bad source")
    return dist, prev, dict(sweeps=int(stats[0]), relaxed_edges=int(stats[1]), vertex_updates=int(stats[2]),
                            graph_edges=int(stats[3]))


def _nt(nthreads):
    """None = every host core: the sweeps are double-buffered, results do not depend on the thread count (tests/test_oracle.py)."""
    return num_threads() if nthreads is None else int(nthreads)


def num_threads():
    return int(lib().ora_num_threads())
