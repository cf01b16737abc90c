"""Host-side mirror of the reference's exported interface (src/RayTracer.jl:24-34) on top of the C ABI.

Julia is not installed in the build/bench image, so this Python layer plays the part of julia/RayTracerB200.jl
(which binds the same entry points with `ccall`): same function names, argument meaning, 1-based ids and error
behaviour, so that tests read like the README example of the reference (README.md:15-53):

    gr, G, halo = init_annulus(ntheta, nr, spacing=1)
    source      = closest_point(gr, 0.0, R, system="polar")
    profile     = velocity_profile()
    Vp          = interpolate_velocity(gr.r, LinearInterpolation(profile.r, profile.Vp))
    D           = bfm(G, halo, source, gr, Vp)          # D.dist, D.prev
    path        = recontruct_path(D.prev, source, receiver)

All computation happens in librt_sssp.so on the GPU; there is no CPU path here.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import RtError, RtStats, check, lib, ptr

R = 6371.0  # src/utils.jl:2

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ------------------------------------------------------------------------------------------------ handles
class _Handle:
    """Owns an rt_mesh*."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            if self.h:
                lib().rt_mesh_free(self.h)
        except Exception:
            pass
        self.h = None

    def sizes(self):
        s = np.zeros(8, np.int64)
        check(lib().rt_mesh_sizes(self.h, s))
        return s

    def set_option(self, key, value):
        check(lib().rt_set_option(self.h, key.encode(), float(value)))

    def coords_dev(self):
        p = [C.c_void_p() for _ in range(4)]
        check(lib().rt_mesh_coords_dev(self.h, *[C.byref(q) for q in p]))
        return [q.value for q in p]


class SparseMatrixCSC:
    """G::SparseMatrixCSC{Bool,Int64} (nel x n): colptr / rowval exactly as Julia stores them (1-based)."""

    def __init__(self, m, n, colptr, rowval):
        self.m, self.n, self.colptr, self.rowval = int(m), int(n), colptr, rowval

    @property
    def nnz(self):
        return len(self.rowval)


class Grid2D:
    """Grid2D of src/GridAnnulus.jl:9-21.  e2n is kept flattened (e2n_off 0-based offsets, e2n_idx 1-based ids);
    `gr.e2n[el]` (el 1-based) returns the node list of an element like the reference Dict."""

    def __init__(self, x, z, theta, r, e2n_off, e2n_idx, ntheta, nr, nel, nnods, nbr_off=None, nbr_idx=None,
                 element_type=None, handle=None):
        self.x, self.z, self.theta, self.r = x, z, theta, r
        self.e2n_off, self.e2n_idx = e2n_off, e2n_idx
        self.ntheta, self.nr, self.nel, self.nnods = int(ntheta), int(nr), int(nel), int(nnods)
        self.nbr_off, self.nbr_idx, self.element_type = nbr_off, nbr_idx, element_type
        self._handle = handle
        self._handle_key = None

    def __len__(self):
        return self.nnods

    class _E2N:
        def __init__(self, g):
            self.g = g

        def __getitem__(self, el):
            return self.g.e2n_idx[self.g.e2n_off[el - 1]:self.g.e2n_off[el]]

    @property
    def e2n(self):
        return Grid2D._E2N(self)

    def neighbours(self, el):
        return self.nbr_idx[self.nbr_off[el - 1]:self.nbr_off[el]]


class AbstractSPM:
    """abstract type AbstractSPM (src/SSSP/ssspm.jl:1): result structs hold (prev, dist); prev is int64 1-based
    (0 = never set).  Base.getindex(spm) = spm.prev (:12) is spelled spm[()]."""

    def __init__(self, prev, dist, stats=None):
        self.prev, self.dist, self.stats = prev, dist, stats

    def __getitem__(self, key):
        if key != ():
            raise IndexError("only spm[()] (Base.getindex(spm::AbstractSPM) = spm.prev) is defined")
        return self.prev


class BellmanFordMoore(AbstractSPM):
    """BellmanFordMoore(prev, dist) of src/SSSP/ssspm.jl:3-10."""


class Dijkstra(AbstractSPM):
    """Dijkstra(prev, dist) of src/SSSP/ssspm.jl:3-10 (result of `dijkstra`)."""


class RadiusStepping(AbstractSPM):
    """RadiusStepping(prev, dist) of src/SSSP/ssspm.jl:3-10 (result of `radius_stepping`)."""


class VelProfile:
    def __init__(self, r, Vp, Vs):
        self.r, self.Vp, self.Vs = r, Vp, Vs


class LinearInterpolation:
    """Stand-in for Interpolations.LinearInterpolation(knots, values) (README.md:32): holds the knots; evaluation
    happens on the device in interpolate_velocity."""

    def __init__(self, knots, values):
        self.knots = np.ascontiguousarray(knots, np.float64)
        self.values = np.ascontiguousarray(values, np.float64)
        if self.knots.shape != self.values.shape or self.knots.ndim != 1 or len(self.knots) < 2:
            raise ValueError("knots and values must be 1-D arrays of equal length >= 2")


# ----------------------------------------------------------------------------------------------- builders
def init_annulus(ntheta, nr, spacing=20, export=True):
    """gr, G, halo = init_annulus(ntheta, nr; spacing) -- src/GridAnnulus.jl:57-70, built on the device.
    With export=False the host arrays are not materialised (gr only carries the device handle and sizes)."""
    h = C.c_void_p()
    check(lib().rt_annulus_build(int(ntheta), int(nr), float(spacing), C.byref(h)))
    handle = _Handle(h)
    n, nel, se, nnz, hr, sn, nth, nrr = (int(v) for v in handle.sizes())
    if not export:
        gr = Grid2D(None, None, None, None, None, None, nth, nrr, nel, n, handle=handle)
        return gr, SparseMatrixCSC(nel, n, None, None), None
    x, z, th, r = (np.zeros(n) for _ in range(4))
    e2n_off = np.zeros(nel + 1, np.int64)
    e2n_idx = np.zeros(se, np.int64)
    colptr = np.zeros(n + 1, np.int64)
    rowval = np.zeros(nnz, np.int64)
    halo = np.zeros(2 * hr, np.int64)
    nbr_off = np.zeros(nel + 1, np.int64)
    nbr_idx = np.zeros(sn, np.int64)
    el_type = np.zeros(nel, np.int8)
    check(lib().rt_mesh_export(h, ptr(x), ptr(z), ptr(th), ptr(r), ptr(e2n_off), ptr(e2n_idx), ptr(colptr),
                               ptr(rowval), ptr(halo), ptr(nbr_off), ptr(nbr_idx), ptr(el_type)))
    gr = Grid2D(x, z, th, r, e2n_off, e2n_idx, nth, nrr, nel, n, nbr_off, nbr_idx, el_type, handle=handle)
    G = SparseMatrixCSC(nel, n, colptr, rowval)
    return gr, G, halo.reshape(2, hr).T.copy()  # (2H, 2) like Matrix{Int64}


def mesh_from_arrays(gr, G, halo):
    """Adopt arrays built elsewhere (by the reference, or by a test) -> device handle (cached on gr)."""
    key = (id(G.colptr), id(G.rowval), id(halo) if halo is not None else 0)
    if gr._handle is not None and (gr._handle_key is None or gr._handle_key == key):
        return gr._handle
    n, nel = int(G.n), int(gr.nel)
    if halo is None or len(halo) == 0:
        halo_cm, hr = None, 0
    else:
        halo = np.asarray(halo, np.int64)
        hr = halo.shape[0]
        halo_cm = np.ascontiguousarray(halo.T).reshape(-1)  # column-major (2H x 2)
    h = C.c_void_p()
    c = lambda a, t: np.ascontiguousarray(a, t)
    th = None if gr.theta is None else c(gr.theta, np.float64)
    rr = None if gr.r is None else c(gr.r, np.float64)
    check(lib().rt_mesh_from_arrays(n, nel, c(gr.e2n_off, np.int64), c(gr.e2n_idx, np.int64),
                                    c(G.colptr, np.int64), c(G.rowval, np.int64), ptr(halo_cm), hr,
                                    c(gr.x, np.float64), c(gr.z, np.float64), ptr(th), ptr(rr), C.byref(h)))
    gr._handle = _Handle(h)
    gr._handle_key = key
    return gr._handle


class Grid3D:
    """grid(c0, c1, nnods) of src/StructuredGrid.jl:35-45 with its star-L adjacency (nodal_incidence :177-223)
    kept implicit on the device."""

    def __init__(self, c0, c1, nnods, neighbour_levels=1, coord_system="cartesian"):
        self.c0 = np.asarray(c0, np.float64).copy()
        self.c1 = np.asarray(c1, np.float64).copy()
        self.nnods = tuple(int(v) for v in nnods)
        self.n = int(np.prod(self.nnods))
        self.neighbour_levels = int(neighbour_levels)
        self.coord_system = coord_system
        h = C.c_void_p()
        cs = {"cartesian": 0, "spherical": 1}[coord_system]
        check(lib().rt_grid3d_build(self.c0, self.c1, np.asarray(self.nnods, np.int64), self.neighbour_levels, cs,
                                    C.byref(h)))
        self._handle = _Handle(h)

    def coordinates(self):
        """Cartesian node coordinates (x fastest), i.e. spherical2cart(gr) for spherical grids."""
        X, Y, Z = np.zeros(self.n), np.zeros(self.n), np.zeros(self.n)
        check(lib().rt_grid3d_export(self._handle.h, X, Y, Z))
        return X, Y, Z

    # ---- the fields / indexing of the reference's Grid (src/StructuredGrid.jl:7-16, 57-104)
    @property
    def nels(self):
        return tuple(v - 1 for v in self.nnods)

    @property
    def nxny(self):
        return self.nnods[0] * self.nnods[1]

    def _axes(self):
        if getattr(self, "_ax", None) is None:
            ax = [np.zeros(v) for v in self.nnods]
            check(lib().rt_grid3d_axes(self._handle.h, ptr(ax[0]), ptr(ax[1]), ptr(ax[2])))
            self._ax = ax
        return self._ax

    x = property(lambda self: self._axes()[0])
    y = property(lambda self: self._axes()[1])
    z = property(lambda self: self._axes()[2])

    def __getitem__(self, key):
        """gr[I] (linear, 1-based, x fastest; :77-81) and gr[I, J, K] (:57-62) -> Point(x, y, z) of the RAW axes.
        An array of linear indices gives an (m, 3) array."""
        if isinstance(key, tuple):
            if len(key) != 3:
                raise IndexError("gr[I] or gr[I, J, K]")
            I, J, K = (int(v) for v in key)
            for v, nmax in zip((I, J, K), self.nnods):
                if not 1 <= v <= nmax:  # the reference @asserts only the upper bound; Julia then throws BoundsError
                    raise IndexError("index out of range")
            ax = self._axes()
            return Point(ax[0][I - 1], ax[1][J - 1], ax[2][K - 1])
        ids = np.atleast_1d(np.asarray(key, np.int64)).copy()
        xyz = np.zeros((len(ids), 3))
        check(lib().rt_grid3d_points(self._handle.h, ids, len(ids), ptr(xyz), None))
        return Point(*xyz[0]) if np.ndim(key) == 0 else xyz

    def CartesianIndex(self, I):
        """CartesianIndex(gr, I) (:90-96) -> (i, j, k), 1-based; arrays give an (m, 3) array."""
        ids = np.atleast_1d(np.asarray(I, np.int64)).copy()
        ijk = np.zeros((len(ids), 3), np.int64)
        check(lib().rt_grid3d_points(self._handle.h, ids, len(ids), None, ptr(ijk)))
        return tuple(int(v) for v in ijk[0]) if np.ndim(I) == 0 else ijk


class Point:
    """Point{T}(x, y, z) of src/StructuredGrid.jl:27-31."""

    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __eq__(self, o):
        return tuple(self) == tuple(o)

    def __repr__(self):
        return "Point(%r, %r, %r)" % (self.x, self.y, self.z)


def connectivity(gr3, iel=None):
    """connectivity(gr) -> (nel, 8) int64 array of 1-based corner ids; connectivity(gr, iel) -> the 8-tuple of one
    element (src/StructuredGrid.jl:121-168)."""
    nel = int(np.prod([max(v, 0) for v in gr3.nels]))
    if iel is None:
        out = np.zeros((nel, 8), np.int64)
        if nel:
            check(lib().rt_grid3d_connectivity(gr3._handle.h, 1, nel, out.reshape(-1)))
        return out
    out = np.zeros(8, np.int64)
    check(lib().rt_grid3d_connectivity(gr3._handle.h, int(iel), 1, out))
    return tuple(int(v) for v in out)


def polardistance3D(p1, p2):
    """polardistance3D(p1, p2) src/StructuredGrid.jl:245-255: p = (theta, phi, r) -> Euclidean distance of the
    spherical2cart images.  Accepts Points / triples, or (m, 3) arrays for a batch."""
    a = np.ascontiguousarray(np.atleast_2d(np.asarray([tuple(p1)] if isinstance(p1, Point) else p1, np.float64)))
    b = np.ascontiguousarray(np.atleast_2d(np.asarray([tuple(p2)] if isinstance(p2, Point) else p2, np.float64)))
    if a.shape != b.shape or a.shape[1] != 3:
        raise ValueError("points must be (theta, phi, r) triples")
    out = np.zeros(len(a))
    check(lib().rt_polardistance3d(a.reshape(-1), b.reshape(-1), len(a), out))
    return float(out[0]) if (isinstance(p1, Point) or np.ndim(p1) == 1) else out


def grid(c0, c1, nnods, neighbour_levels=1, coord_system="cartesian"):
    return Grid3D(c0, c1, nnods, neighbour_levels, coord_system)


# ----------------------------------------------------------------------------------------------- velocity
def velocity_profile():
    """velocity_profile() of src/utils.jl:23-30 (AK135; the reference's IASP91 file is byte-identical).
    The table is shipped as data/ak135_profile.npz (made by tools/make_ak135_fixture.py)."""
    d = np.load(os.path.join(_DATA, "ak135_profile.npz"))
    depth = d["depth_km"]
    r = depth.max() - depth
    return VelProfile(r[::-1].copy(), d["vp"][::-1].copy(), d["vs"][::-1].copy())


def interpolate_velocity(r, interpolant, buffer=None):
    """interpolate_velocity(r, interpolant) src/utils.jl:38-44; with `buffer` the variant of
    src/ShortestPath.jl:74-90.  Raises (like the reference's BoundsError) if a point is outside the knots."""
    r = np.ascontiguousarray(r, np.float64)
    out = np.empty_like(r)
    check(lib().rt_interp_velocity(interpolant.knots, interpolant.values, len(interpolant.knots), r.reshape(-1),
                                   r.size, -1.0 if buffer is None else float(buffer), out.reshape(-1)))
    return out


def interpolate_inplace(V, gr, G=None, halo=None):
    """interpolate!(V, gr) of src/Interpolations/interpolation.jl:5-18 (Python cannot spell the `!`): cell-wise
    bilinear / barycentric re-interpolation of the secondary-node velocities, in place.  Returns V."""
    handle = gr._handle if gr._handle is not None else mesh_from_arrays(gr, G, halo)
    if not (isinstance(V, np.ndarray) and V.dtype == np.float64 and V.flags.c_contiguous):
        raise ValueError("V must be a contiguous float64 array (it is updated in place)")
    et = None if gr.element_type is None else np.ascontiguousarray(gr.element_type, np.int8)
    check(lib().rt_interpolate_cells(handle.h, ptr(et), V))
    return V


# ----------------------------------------------------------------------------------------------- topology
class SparseAdjencyList:
    """SparseAdjencyList{list, deg, idx} of src/topology/topology.jl:88-92 (sic): CSR of the star-0 node adjacency;
    idx is 1-based like the reference (idx = 1 + cumsum(deg))."""

    def __init__(self, lst, deg, idx):
        self.list, self.deg, self.idx = lst, deg, idx

    def neighbours(self, node):
        return self.list[self.idx[node - 1] - 1:self.idx[node - 1] - 1 + self.deg[node - 1]]


def _handle_of(gr, G=None, halo=None):
    return gr._handle if gr._handle is not None else mesh_from_arrays(gr, G, halo)


def nodal_degree(gr, G=None, halo=None):
    """nodal_degree(nodal_incidence(gr)) -- src/topology/topology.jl:70-77 on src/GridAnnulus.jl:763-804."""
    deg = np.zeros(gr.nnods, np.int64)
    check(lib().rt_nodal_adjacency(_handle_of(gr, G, halo).h, ptr(deg), None, None, 0))
    return deg


def sparse_adjacency_list(gr, G=None, halo=None):
    """sparse_adjacency_list(nodal_incidence(gr)) -- src/topology/topology.jl:94-111."""
    h = _handle_of(gr, G, halo)
    n = gr.nnods
    deg = np.zeros(n, np.int64)
    off = np.zeros(n + 1, np.int64)
    check(lib().rt_nodal_adjacency(h.h, ptr(deg), ptr(off), None, 0))
    lst = np.zeros(int(off[-1]), np.int64)
    check(lib().rt_nodal_adjacency(h.h, None, None, ptr(lst), len(lst)))
    return SparseAdjencyList(lst, deg, off[:-1] + 1)


class AdjencyList:
    """AdjencyList{G, N} of src/topology/topology.jl:1-4 (sic): G is the zero-padded [maxdeg x nnods] neighbour matrix
    (column j = neighbours of node j, 1-based, 0 = padding; Julia column-major == this array's Fortran order), N the
    nodal degrees."""

    def __init__(self, G, N):
        self.G, self.N = G, N


def adjacency_list(gr, G=None, halo=None):
    """adjacency_list(nodal_incidence(gr)) -- src/topology/topology.jl:52-68: the dense padded form of the same
    star-0 adjacency (neighbours in ascending id; the reference's order is the iteration order of a Julia Set).
    Only sensible on coarse meshes: the matrix has maxdeg x nnods Int32 entries."""
    sp = sparse_adjacency_list(gr, G, halo)
    n = gr.nnods
    maxdeg = int(sp.deg.max()) if n else 0
    out = np.zeros((maxdeg, n), np.int32, order="F")
    col = np.repeat(np.arange(n), sp.deg)
    row = np.arange(len(sp.list)) - np.repeat(sp.idx - 1, sp.deg)
    out[row, col] = sp.list
    return AdjencyList(out, sp.deg.astype(np.int32))


def element_degree(G):
    """element_degree(IM) -- src/topology/topology.jl:79-86: number of elements in each column of G."""
    return np.diff(G.colptr).astype(np.int64)


def symrcm(gr, G=None, halo=None):
    """symrcm(nodal_incidence(gr), degrees) -- src/SSSP/rcm.jl:2-46.  Returns prm (1-based int64)."""
    prm = np.zeros(gr.nnods, np.int64)
    check(lib().rt_rcm(_handle_of(gr, G, halo).h, prm))
    return prm


def reorder(gr, G, halo, prm):
    """reorder!(gr, prm) of src/SSSP/rcm.jl:62-85 as a pure function: coordinates permuted (x .= x[prm]), e2n
    relabelled with the inverse map (rordering_map :87-94).  Unlike the reference (which forgets `halo` and leaves
    G stale) the halo matrix and the columns of G are relabelled too, so the result can be passed to bfm."""
    prm = np.asarray(prm, np.int64)
    n = gr.nnods
    inv = np.zeros(n + 1, np.int64)
    inv[prm] = np.arange(1, n + 1)
    sel = prm - 1
    col_len = np.diff(G.colptr)[sel]
    colptr = np.concatenate([[1], 1 + np.cumsum(col_len)]).astype(np.int64)
    starts = G.colptr[:-1][sel] - 1
    take = np.repeat(starts - (colptr[:-1] - 1), col_len) + np.arange(int(col_len.sum()))
    gr2 = Grid2D(gr.x[sel].copy(), gr.z[sel].copy(), gr.theta[sel].copy(), gr.r[sel].copy(), gr.e2n_off.copy(),
                 inv[gr.e2n_idx], gr.ntheta, gr.nr, gr.nel, n, gr.nbr_off, gr.nbr_idx, gr.element_type)
    G2 = SparseMatrixCSC(G.m, n, colptr, G.rowval[take].copy())
    halo2 = None if halo is None else inv[np.asarray(halo, np.int64)]
    return gr2, G2, halo2


def dual_velocity(r, interpolant, buffer=1):
    """dual_velocity(r, interpolant; buffer) src/utils.jl:51-66 -> (n, 2) array: column 0 = U[:,1] (value just
    below a discontinuity), column 1 = U[:,2] (just above); both equal away from the discontinuities."""
    r = np.ascontiguousarray(r, np.float64)
    out = np.empty(2 * r.size, np.float64)
    check(lib().rt_dual_velocity(interpolant.knots, interpolant.values, len(interpolant.knots), r.reshape(-1), r.size,
                                 float(buffer), out))
    return out.reshape(2, r.size).T.copy()


# ------------------------------------------------------------------------------------------ closest_point
def closest_point(gr, px, pz, *rest, system=None):
    """closest_point(gr, px, pz; system) src/GridAnnulus.jl:823-840 -> 1-based node id (scalar or array); `system`
    may also be given as the fourth positional argument.
    With a Grid3D: closest_point(gr, x, y, z) of src/StructuredGrid.jl:257-270 (raw axis coordinates)."""
    if isinstance(gr, Grid3D):
        if len(rest) != 1 or system is not None:
            raise TypeError("closest_point(gr::Grid, x, y, z) needs three coordinates")
        pw = rest[0]
        a, b, c = np.broadcast_arrays(*(np.atleast_1d(np.asarray(v, np.float64)) for v in (px, pz, pw)))
        a, b, c = (np.ascontiguousarray(v) for v in (a, b, c))
        out = np.zeros(len(a), np.int64)
        check(lib().rt_closest_point3d(gr._handle.h, a, b, c, len(a), out))
        return int(out[0]) if all(np.ndim(v) == 0 for v in (px, pz, pw)) else out
    if len(rest) > 1 or (rest and system is not None):
        raise TypeError("closest_point(gr, px, pz; system)")
    system = rest[0] if rest else (system or "cartesian")
    handle = gr._handle
    if handle is None:
        raise ValueError("grid has no device handle; call mesh_from_arrays(gr, G, halo) or bfm(...) first")
    pa = np.atleast_1d(np.asarray(px, np.float64)).copy()
    pb = np.atleast_1d(np.asarray(pz, np.float64)).copy()
    pa, pb = np.broadcast_arrays(pa, pb)
    pa, pb = np.ascontiguousarray(pa), np.ascontiguousarray(pb)
    out = np.zeros(len(pa), np.int64)
    sysid = {"cartesian": 0, "polar": 1}[system]
    check(lib().rt_closest_point(handle.h, pa, pb, len(pa), sysid, out))
    return int(out[0]) if np.ndim(px) == 0 and np.ndim(pz) == 0 else out


# ------------------------------------------------------------------------------------------------- solver
def _solve(handle, n, U, sources, want_prev=True, precision=64):
    U = np.ascontiguousarray(U, np.float64)
    if U.shape != (n,):
        raise ValueError("U must have one entry per node (%d), got %s" % (n, U.shape))
    src = np.atleast_1d(np.asarray(sources, np.int64)).copy()
    dist = np.empty((len(src), n), np.float64)
    prev = np.empty((len(src), n), np.int64) if want_prev else None
    st = RtStats()
    if precision not in (64, 32):
        raise ValueError("precision must be 64 or 32")
    check(lib().rt_bfm_solve(handle.h, U, src, len(src), int(precision), ptr(dist), ptr(prev), C.byref(st)))
    if precision == 32:  # the values are Float32 numbers held in float64 storage: the narrowing is exact
        dist = dist.astype(np.float32)
    return dist, prev, st.as_dict()


def _solve_dual(handle, n, U2, source):
    U2 = np.asarray(U2, np.float64)
    if U2.shape != (n, 2):
        raise ValueError("dual velocity must have shape (%d, 2), got %s" % (n, U2.shape))
    ucm = np.ascontiguousarray(U2.T).reshape(-1)  # column-major like the Julia Matrix
    src = np.atleast_1d(np.asarray(source, np.int64)).copy()
    dist = np.empty((len(src), n), np.float64)
    prev = np.empty((len(src), n), np.int64)
    st = RtStats()
    check(lib().rt_bfm_solve_dual(handle.h, ucm, src, len(src), ptr(dist), ptr(prev), C.byref(st)))
    if np.ndim(source) == 0:
        return BellmanFordMoore(prev[0], dist[0], st.as_dict())
    return BellmanFordMoore(prev, dist, st.as_dict())


SCHEDULES = {"jacobi": 0, "near-far": 1}


def bfm(G, halo, source, gr, U, schedule=None, delta=None, precision=64, canonical_prev=None):
    """D = bfm(G, halo, source, gr, U) -- src/SSSP/bfm.jl:1-52.  `source` may be an array of sources, in which
    case D.dist / D.prev are [nsrc x n] tables (the batch API); a scalar gives vectors as in the reference.

    schedule (extension): "jacobi" = the reference's sweeps (dist and prev bit-identical, ties included);
    "near-far" = work-efficient push schedule (dist bit-identical, prev identical except on exact ties).
    precision=32: the Float32 arithmetic of bfm_gpu (src/SSSP/bfm_gpu.jl:170-205): x, z, U cast to Float32, travel
    times relaxed in Float32; D.dist is a float32 array.
    canonical_prev=True (near-far): a post-pass rebuilds the reference's predecessors exactly, ties included
    (rt_set_option "canonical_prev"; costs about one sweep over the graph)."""
    handle = mesh_from_arrays(gr, G, halo)
    if schedule is not None:
        handle.set_option("schedule", SCHEDULES[schedule])
    if canonical_prev is not None:
        handle.set_option("canonical_prev", 1 if canonical_prev else 0)
    if delta is not None:
        handle.set_option("delta", delta)
    if np.ndim(U) == 2:  # U::Matrix -> dual-velocity relax (bfm.jl:113-159)
        if precision != 64:
            raise ValueError("the dual-velocity relax is Float64 only")
        return _solve_dual(handle, int(G.n), U, source)
    dist, prev, st = _solve(handle, int(G.n), U, source, precision=precision)
    if np.ndim(source) == 0:
        return BellmanFordMoore(prev[0], dist[0], st)
    return BellmanFordMoore(prev, dist, st)


def bfm_gpu(G, halo, source, gr, U, schedule=None, canonical_prev=None):
    """bfm_gpu(G, halo, source, gr, U) src/SSSP/bfm_gpu.jl:212-250: the reference's Float32 device path."""
    return bfm(G, halo, source, gr, U, schedule=schedule, precision=32, canonical_prev=canonical_prev)


def bfm3d(gr3, source, U, schedule=None, delta=None, precision=64, canonical_prev=None):
    """BFM(G, source, gr, U, fw) of src/Dijsktra.jl:294-343 on the implicit star-L graph of a Grid3D with the
    edge weight of src/SSSP/weights.jl:20 (option "weight3d" = 1: the expression inside foo!, Dijsktra.jl:388).
    canonical_prev=True (near-far): predecessors of the reference schedule, ties included.
    schedule="near-far" runs tile-pull rounds (option "tile_pull" = 0: push units); a batch of sources keeps up to four
    of them in flight in per-source slots (option "batch" = 1: one after the other)."""
    if schedule is not None:
        gr3._handle.set_option("schedule", SCHEDULES[schedule])
    if canonical_prev is not None:
        gr3._handle.set_option("canonical_prev", 1 if canonical_prev else 0)
    if delta is not None:
        gr3._handle.set_option("delta", delta)
    dist, prev, st = _solve(gr3._handle, gr3.n, U, source, precision=precision)
    if np.ndim(source) == 0:
        return BellmanFordMoore(prev[0], dist[0], st)
    return BellmanFordMoore(prev, dist, st)


# ------------------------------------------------------------------------------------------------ multiphase
RLAYER = tuple(R - d for d in (20.0, 35.0, 210.0, 410.0, 660.0, 2740.0, 2891.5))  # topology.jl:184-192


class GridPartition:
    """GridPartition of src/topology/topology.jl:150-181.  `code` holds the node ids as integers (k > 0 = "Layer_k",
    -k = "Boundary_k"); `id` materialises the reference's Vector{String} on demand.  layers / boundaries / iterator are
    the reference's name tuples and its LayerIterations dictionary (keys 1 .. 2 nlayers - 1)."""

    def __init__(self, code, rboundaries=RLAYER):
        self.code = np.ascontiguousarray(code, np.int32)
        self.rboundaries = tuple(rboundaries)
        self.nboundaries = len(self.rboundaries)
        self.nlayers = self.nboundaries + 1
        self.layers = tuple("Layer_%d" % (i + 1) for i in range(self.nlayers))
        self.boundaries = tuple("Boundary_%d" % (i + 1) for i in range(self.nboundaries))
        nl, nmax = self.nlayers, 2 * self.nlayers - 1
        it = {1: (self.layers[0], self.boundaries[0]), nmax: (self.layers[0], self.boundaries[0])}
        for i in range(2, nl):
            it[i] = (self.layers[i - 1], self.boundaries[i - 2], self.boundaries[i - 1])
            it[nmax - i + 1] = (self.layers[i - 1], self.boundaries[i - 2], self.boundaries[i - 1])
        it[nl] = (self.layers[-1], self.boundaries[-1])
        self.iterator = it

    @property
    def id(self):
        return np.array([("Layer_%d" % c) if c > 0 else ("Boundary_%d" % -c) for c in self.code], dtype=object)

    def code_of(self, name):
        kind, k = name.split("_")
        return int(k) if kind == "Layer" else -int(k)

    def mask(self, names):
        """allowed[i] = ID[i] in names (`ID[Gi] ∉ current_level && continue`, bfm_multiphase.jl:121)."""
        return np.isin(self.code, [self.code_of(nm) for nm in names]).astype(np.uint8)


def partition_grid(gr):
    """partition_grid(gr) src/topology/topology.jl:183-206 (computed on the device from gr.r)."""
    code = np.zeros(gr.nnods, np.int32)
    check(lib().rt_partition_grid(_handle_of(gr).h, code))
    return GridPartition(code)


def directions(nlayers):
    """directions(nlayers) src/SSSP/bfm_multiphase.jl:2-14."""
    nmax = 2 * nlayers - 1
    d = {1: ("above", "above"), nmax: ("above", "above")}
    for i in range(2, nlayers):
        d[i] = d[nmax - i + 1] = ("below", "above")
    d[nlayers] = ("below", "below")
    return d


def bfm_continue(G, halo, gr, U, allowed, seeds, dist, prev):
    """Restricted continuation (rt_bfm_continue): reference sweeps from the state (dist, prev) in which only nodes with
    allowed != 0 change; the frontier starts as the allowed nodes of the seeds' star patches.  Returns a new
    BellmanFordMoore; the inputs are not modified."""
    handle = mesh_from_arrays(gr, G, halo)
    n = int(G.n)
    U = np.ascontiguousarray(U, np.float64)
    d = np.array(dist, np.float64)
    p = np.array(prev, np.int64)
    if U.shape != (n,) or d.shape != (n,) or p.shape != (n,):
        raise ValueError("U, dist and prev must have one entry per node (%d)" % n)
    al = None if allowed is None else np.ascontiguousarray(allowed, np.uint8)
    if al is not None and al.shape != (n,):
        raise ValueError("allowed must have one entry per node")
    sd = np.ascontiguousarray(np.atleast_1d(seeds), np.int64)
    st = RtStats()
    check(lib().rt_bfm_continue(handle.h, U, ptr(al), sd, len(sd), d, p, C.byref(st)))
    return BellmanFordMoore(p, d, st.as_dict())


def initial_state(halo, source, n):
    """dist = fill(Inf, n); dist[source] = 0 and p after init_halo_path! (src/SSSP/bfm.jl:12-13, 21-22, 64-70)."""
    dist = np.full(n, np.inf)
    dist[source - 1] = 0.0
    prev = np.zeros(n, np.int64)
    if halo is not None:
        for a, b in np.asarray(halo, np.int64):
            prev[b - 1] = a
            prev[a - 1] = b
    return dist, prev


def bfm_multiphase(G, halo, source, gr, U, partition, interpolant, nphases=3, buffer_zone=1.0):
    """bfm_multiphase(Gsp, source, gr, U, partition, interpolant) -- src/SSSP/bfm_multiphase.jl:30-156, on the graph of
    bfm.  The reference routine is an unfinished draft (undefined _relax_bfm! / fillfalse!, `for i in 1:3`, the restart
    code commented out); this follows its structure and completes it where it says what it intends:
      for phase i = 1 .. nphases: current_level = partition.iterator[i] (a layer and its one or two boundaries);
        boundary_velocity! (:16-28): U at the nodes of each current boundary = interpolant(r_b -/+ buffer_zone) for the
          ray direction :above / :below of directions(nlayers)[i] (the draft passes the whole tuple, so its comparison
          with :above is never true; the per-boundary entry is what it evidently means);
        frontier: phase 1 = the allowed nodes around `source` (:120-123); later phases restart from the nodes of the
          first current boundary that have been reached (the commented-out find_new_source_min / restart block);
        sweeps restricted to ID in current_level (:118-150), continuing from the running (dist, prev) state.
    Returns BellmanFordMoore(prev, dist) like the reference; U is not modified (a copy is)."""
    n = int(G.n)
    U = np.array(U, np.float64)
    rdir = directions(partition.nlayers)
    rb = dict(zip(partition.boundaries, partition.rboundaries))
    bnodes = {nm: np.flatnonzero(partition.code == partition.code_of(nm)) + 1 for nm in partition.boundaries}
    kn, kv = interpolant.knots, interpolant.values
    dist, prev = initial_state(halo, int(source), n)
    D = BellmanFordMoore(prev, dist, {})
    sweeps = []
    for i in range(1, int(nphases) + 1):
        level = partition.iterator[i]
        for k, b in enumerate(level[1:]):
            rq = rb[b] - buffer_zone if rdir[i][k] == "above" else rb[b] + buffer_zone
            U[bnodes[b] - 1] = np.interp(rq, kn, kv)
        if i == 1:
            seeds = np.array([int(source)], np.int64)
        else:
            cand = bnodes[level[1]]
            seeds = cand[np.isfinite(D.dist[cand - 1])]
        D = bfm_continue(G, halo, gr, U, partition.mask(level), seeds, D.dist, D.prev)
        sweeps.append(D.stats["sweeps"])
    D.stats["phase_sweeps"] = sweeps
    return D


def _sssp_nodal(gr, source, U, algorithm, G=None, halo=None):
    h = _handle_of(gr, G, halo)
    n = gr.nnods
    U = np.ascontiguousarray(U, np.float64)
    if U.shape != (n,):
        raise ValueError("U must have one entry per node (%d), got %s" % (n, U.shape))
    dist = np.empty(n)
    prev = np.empty(n, np.int64)
    st = RtStats()
    check(lib().rt_sssp_nodal(h.h, U, int(source), algorithm, dist, prev, C.byref(st)))
    return prev, dist, st.as_dict()


def dijkstra(G, source, gr, U):
    """D = dijkstra(G, source, gr, U) -- src/SSSP/dijkstra.jl:68-136: G is the node graph nodal_incidence(gr) (star-0; the
    reference passes it as a Dict, here it is implied by `gr`: pass None, or the SparseAdjencyList for symmetry).
    Returns Dijkstra(prev, dist); unreachable nodes keep dist = Inf, prev = 0 (no halo coupling on this graph)."""
    return Dijkstra(*_sssp_nodal(gr, source, U, 0))


def radius_stepping(Gsp, source, gr, U):
    """D = radius_stepping(Gsp, source, gr, U) -- src/SSSP/radius_stepping.jl:7-46 on the same node graph.
    Returns RadiusStepping(prev, dist)."""
    return RadiusStepping(*_sssp_nodal(gr, source, U, 1))


# -------------------------------------------------------------------------------------------------- paths
def recontruct_path(prev, source, receiver):
    """recontruct_path(prev::Vector, source, receiver) src/SSSP/ssspm.jl:30-40 -> [receiver, ..., source].
    (The misspelling is the reference's API.)  `receiver` may be an array -> list of paths.
    Passing a result struct (BellmanFordMoore / Dijkstra / RadiusStepping) dispatches, as in Julia, to the struct
    method recontruct_path(D, source, receiver) :14-28: the chase runs until a node repeats, then `source` is
    appended; it raises RtError (the reference: BoundsError) when it reads an unset predecessor."""
    fn = lib().rt_reconstruct_paths
    if isinstance(prev, AbstractSPM):
        prev = prev.prev
        fn = lib().rt_reconstruct_paths_guarded
    prev = np.ascontiguousarray(prev, np.int64)
    rec = np.atleast_1d(np.asarray(receiver, np.int64)).copy()
    off = np.zeros(len(rec) + 1, np.int64)
    check(fn(prev, len(prev), int(source), rec, len(rec), off, None, 0))
    idx = np.zeros(int(off[-1]), np.int64)
    check(fn(prev, len(prev), int(source), rec, len(rec), off, ptr(idx), len(idx)))
    paths = [idx[off[k]:off[k + 1]].copy() for k in range(len(rec))]
    return paths[0] if np.ndim(receiver) == 0 else paths


def travel_times(D, gr, receivers, isave=False, flname=""):
    """travel_times(D, gr, receivers; isave, flname) src/utils.jl:4-15: D.dist at the receiver nodes; with isave the
    (degree, travel_time) table is written as CSV like the reference's DataFrame (degree = rad2deg(gr.theta))."""
    rec = np.ascontiguousarray(np.atleast_1d(np.asarray(receivers, np.int64)))
    d = np.ascontiguousarray(D.dist, np.float64)
    nsrc = 1 if d.ndim == 1 else d.shape[0]
    tt = np.zeros(nsrc * len(rec))
    check(lib().rt_travel_times(d.reshape(-1), d.shape[-1], nsrc, rec, len(rec), tt))  # device gather
    tt = tt if d.ndim == 1 else tt.reshape(nsrc, len(rec))
    if isave:
        if d.ndim != 1:
            raise ValueError("isave writes the (degree, travel_time) table of ONE source")
        deg = np.rad2deg(np.asarray(gr.theta, np.float64)[rec - 1])
        with open(os.path.join(os.getcwd(), flname), "w") as f:
            f.write("degree,travel_time\n")
            for a, b in zip(deg, tt):
                f.write("%r,%r\n" % (float(a), float(b)))
    return tt


def set_device(device):
    """Selects the CUDA device for the meshes built next (rt_set_device)."""
    check(lib().rt_set_device(int(device)))


def bfm_multi(grids, sources, U, schedule=None, precision=64):
    """Batch of earthquakes over several GPUs from ONE process (rt_bfm_solve_multi): `grids` are replicas of the same
    mesh, one per device (build each after set_device(d): Grid2D objects from init_annulus / mesh_from_arrays, or
    Grid3D objects); replica d solves a contiguous block of `sources` on its own host thread.  Returns
    BellmanFordMoore with [nsrc x n] tables in the order of `sources`."""
    handles = [g._handle for g in grids]
    if any(h is None for h in handles):
        raise ValueError("every replica needs a device handle (init_annulus / mesh_from_arrays / grid)")
    n = grids[0].n if isinstance(grids[0], Grid3D) else int(grids[0].nnods)
    U = np.ascontiguousarray(U, np.float64)
    if U.shape != (n,):
        raise ValueError("U must have one entry per node (%d), got %s" % (n, U.shape))
    if schedule is not None:
        for h in handles:
            h.set_option("schedule", SCHEDULES[schedule])
    src = np.atleast_1d(np.asarray(sources, np.int64)).copy()
    dist = np.empty((len(src), n), np.float64)
    prev = np.empty((len(src), n), np.int64)
    st = RtStats()
    arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
    check(lib().rt_bfm_solve_multi(arr, len(handles), U, src, len(src), int(precision), ptr(dist), ptr(prev),
                                   C.byref(st)))
    if precision == 32:
        dist = dist.astype(np.float32)
    return BellmanFordMoore(prev, dist, st.as_dict())


def device_count():
    c = C.c_int(0)
    check(lib().rt_device_count(C.byref(c)))
    return c.value
