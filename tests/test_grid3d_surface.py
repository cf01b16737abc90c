"""The rest of the 3-D grid surface: nodal_incidence window, weights, getindex / CartesianIndex,
connectivity, closest_point, polardistance3D -- src/StructuredGrid.jl:57-104, 121-168, 177-270, src/SSSP/weights.jl,
src/Dijsktra.jl:388.  CPU tests pin the oracle restatements; the gpu-marked tests compare the CUDA path with them."""
import numpy as np
import pytest

from conftest import splitmix64
from helpers import ulp_diff

R = 6371.0


# ------------------------------------------------------------------------------------------------ CPU (oracle)
@pytest.mark.parametrize("nn", [(7, 6, 5), (3, 3, 3), (2, 5, 10), (12, 11, 10)])
@pytest.mark.parametrize("lv", [0, 1, 2, 3])
def test_window_equals_literal_nodal_incidence(O, nn, lv):
    """nodal_incidence(gr; neighbour_levels) built literally with Dict/Set-union semantics (StructuredGrid.jl:177-212)
    equals the clipped window of half width 2^L that the solvers (oracle and CUDA) use: the radius DOUBLES per level."""
    off, lst = O.nodal_incidence3d(nn, lv)
    w = O.window3d(lv)
    assert w == 2 ** lv
    nx, ny, nz = nn
    for I in range(nx * ny * nz):
        i, j, k = I % nx, (I // nx) % ny, I // (nx * ny)
        win = [ii + nx * (jj + ny * kk) + 1
               for kk in range(max(0, k - w), min(nz - 1, k + w) + 1)
               for jj in range(max(0, j - w), min(ny - 1, j + w) + 1)
               for ii in range(max(0, i - w), min(nx - 1, i + w) + 1)
               if lv >= 1 or ii + nx * (jj + ny * kk) != I]
        assert list(lst[off[I]:off[I + 1]]) == win


def test_weight_mode_foo_oracle(O):
    """weight3d = 1 is the expression inside BFM/foo! (src/Dijsktra.jl:388): fw/abs(U+U)*0.5.  Same graph, so the
    label-correcting result equals an independent heap Dijkstra bit for bit, and is ~1/4 of the weights.jl:20 result."""
    nn = (9, 8, 7)
    X, Y, Z = O.grid3d_coords((0.0, 0.0, 0.0), (90.0, 70.0, 60.0), nn, 0)
    n = len(X)
    U = 4.0 + 6.0 * splitmix64(11, n)
    d0, p0, _ = O.bfm3d(nn, 1, X, Y, Z, U, 5)
    O.set_weight3d(1)
    try:
        d1, p1, _ = O.bfm3d(nn, 1, X, Y, Z, U, 5)
        dj = O.dijkstra3d(nn, 1, X, Y, Z, U, 5)
    finally:
        O.set_weight3d(0)
    assert np.array_equal(d1, dj)
    assert np.allclose(d1 * 4.0, d0, rtol=1e-13)
    i = np.nonzero(p1 > 0)[0]
    j = p1[i] - 1
    d = np.sqrt((X[j] - X[i]) ** 2 + (Y[j] - Y[i]) ** 2 + (Z[j] - Z[i]) ** 2)
    assert np.array_equal(d1[j] + d / np.abs(U[j] + U[i]) * 0.5, d1[i])  # numpy restatement of Dijsktra.jl:388


def test_cartesian_index_and_connectivity_oracle(O):
    nn = (5, 4, 3)
    nx, ny, nz = nn
    for I in range(1, nx * ny * nz + 1):
        i, j, k = O.cartesian_index3d(nn, I)
        assert I == i + nx * (j - 1) + nx * ny * (k - 1) and 1 <= i <= nx and 1 <= j <= ny and 1 <= k <= nz
    e2n = O.connectivity3d(nn)
    assert e2n.shape == (4 * 3 * 2, 8)
    assert list(e2n[0]) == [1, 2, 2 + nx, 1 + nx, 1 + nx * ny, 2 + nx * ny, 2 + nx * ny + nx, 1 + nx * ny + nx]
    # every hex: 8 distinct nodes spanning exactly one cell in each direction
    ijk = np.array([[O.cartesian_index3d(nn, int(v)) for v in row] for row in e2n])
    assert np.all(ijk.max(axis=1) - ijk.min(axis=1) == 1)
    assert len({tuple(sorted(r)) for r in e2n.tolist()}) == len(e2n)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_example3dgrid_script(rt, O):
    """example3Dgrid.jl line for line: grid, gr[1], gr[1,1,1], connectivity(gr), connectivity(gr, 1)."""
    c0, c1, nnods = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (11, 11, 11)
    gr = rt.grid(c0, c1, nnods)
    assert gr[1] == rt.Point(0.0, 0.0, 0.0)
    assert gr[1, 1, 1] == gr[1]
    ax = O.grid3d_axes(c0, c1, nnods)
    assert all(np.array_equal(a, b) for a, b in zip((gr.x, gr.y, gr.z), ax))
    e2n = rt.connectivity(gr)
    assert np.array_equal(e2n, O.connectivity3d(nnods))
    assert rt.connectivity(gr, 1) == tuple(int(v) for v in e2n[0])
    assert rt.connectivity(gr, 1000) == tuple(int(v) for v in e2n[999])
    with pytest.raises(rt.RtError):
        rt.connectivity(gr, 1001)
    # getindex over all linear ids == the Cartesian form, CartesianIndex == the oracle's
    ids = np.arange(1, gr.n + 1)
    pts = gr[ids]
    ijk = gr.CartesianIndex(ids)
    for I in (1, 11, 12, 121, 122, 1331, 700):
        i, j, k = O.cartesian_index3d(nnods, I)
        assert tuple(ijk[I - 1]) == (i, j, k) == gr.CartesianIndex(I)
        assert gr[i, j, k] == rt.Point(*pts[I - 1]) == rt.Point(ax[0][i - 1], ax[1][j - 1], ax[2][k - 1])
    assert gr.nels == (10, 10, 10) and gr.nxny == 121
    with pytest.raises(rt.RtError):
        gr[1332]
    with pytest.raises(IndexError):
        gr[12, 1, 1]


@pytest.mark.gpu
def test_closest_point3d(rt, O):
    c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
    c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    nn = (21, 17, 13)
    g = rt.grid(c0, c1, nn, neighbour_levels=1, coord_system="spherical")
    ax = O.grid3d_axes(c0, c1, nn)
    u = splitmix64(5, 3 * 40).reshape(40, 3)
    q = np.array(c0) + (np.array(c1) - np.array(c0)) * (1.3 * u - 0.15)  # some queries outside the box
    # exact ties: a grid node, the midpoint of two nodes (first index wins), and a point so far away in z that every
    # (x, y) offset is absorbed by rounding -> the first node of the closest z-plane wins
    extra = np.array([[ax[0][3], ax[1][4], ax[2][5]],
                      [0.5 * (ax[0][3] + ax[0][4]), ax[1][2], ax[2][2]],
                      [ax[0][7], ax[1][7], 1e20], [ax[0][7], ax[1][7], -1e20]])
    q = np.vstack([q, extra])
    got = rt.closest_point(g, q[:, 0], q[:, 1], q[:, 2])
    want = np.array([O.closest_point3d(ax, p) for p in q])
    assert np.array_equal(got, want)
    assert got[40] == 1 + 3 + 21 * (4 + 17 * 5)
    assert rt.closest_point(g, *extra[0]) == got[40]
    assert rt.closest_point(g, np.nan, 0.0, 0.0) == -1  # `di < dist` is never true: the reference returns -1
    # more than 65 535 queries in one call (round 1 refused them)
    big = np.tile(q[:41], (1700, 1))
    gb = rt.closest_point(g, big[:, 0], big[:, 1], big[:, 2])
    assert len(gb) == 69700 and np.array_equal(gb.reshape(1700, 41), np.tile(want[:41], (1700, 1)))


@pytest.mark.gpu
def test_closest_point2d_many_queries(rt, O, annulus):
    m = annulus(24, 6, 300.0)
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    rt.mesh_from_arrays(gr, G, m.halo_matrix())
    th = 2 * np.pi * splitmix64(3, 70000)
    got = rt.closest_point(gr, th, np.full(len(th), R), system="polar")
    for k in (0, 1, 33000, 65535, 65536, 69999):
        assert got[k] == O.closest_point(m.theta, m.r, float(th[k]), R)


@pytest.mark.gpu
def test_polardistance3d(rt):
    u = splitmix64(9, 600).reshape(100, 6)
    a = np.stack([np.pi * u[:, 0], 2 * np.pi * u[:, 1], R * u[:, 2]], axis=1)
    b = np.stack([np.pi * u[:, 3], 2 * np.pi * u[:, 4], R * u[:, 5]], axis=1)

    def s2c(p):
        return p[:, 2] * np.cos(p[:, 1]) * np.sin(p[:, 0]), p[:, 2] * np.sin(p[:, 1]) * np.sin(p[:, 0]), p[:, 2] * np.cos(p[:, 0])

    A, B = s2c(a), s2c(b)
    want = np.sqrt((A[0] - B[0]) ** 2 + (A[1] - B[1]) ** 2 + (A[2] - B[2]) ** 2)
    got = rt.polardistance3D(a, b)
    assert np.allclose(got, want, rtol=1e-13, atol=1e-9)  # device sin/cos vs glibc: last-ulp differences, amplified by cancellation
    assert rt.polardistance3D(rt.Point(*a[0]), rt.Point(*b[0])) == got[0]
    assert rt.polardistance3D(a[:1], a[:1])[0] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("lv", [0, 1, 2])
def test_bfm3d_weight_mode_foo(rt, O, lv):
    """The BFM/foo! expression (Dijsktra.jl:388) as the edge weight: both schedules bit-identical to the oracle."""
    nn = (19, 14, 11)
    g = rt.grid((0.0, 0.0, 0.0), (90.0, 70.0, 60.0), nn, neighbour_levels=lv)
    X, Y, Z = g.coordinates()
    n = g.n
    U = 4.0 + 6.0 * splitmix64(21 + lv, n)
    g._handle.set_option("weight3d", 1)
    O.set_weight3d(1)
    try:
        for src in (1, n // 2):
            dist, prev, st = O.bfm3d(nn, lv, X, Y, Z, U, src)
            D = rt.bfm3d(g, src, U, schedule="jacobi")
            assert np.array_equal(D.dist, dist) and np.array_equal(D.prev, prev)
            assert D.stats["sweeps"] == st["sweeps"]
            Dn = rt.bfm3d(g, src, U, schedule="near-far")
            assert np.array_equal(Dn.dist, dist)
            D32 = rt.bfm3d(g, src, U, schedule="near-far", precision=32)
            assert np.array_equal(D32.dist.astype(np.float64), O.bfm3d_f32(nn, lv, X, Y, Z, U, src)[0])
    finally:
        O.set_weight3d(0)
    g._handle.set_option("weight3d", 0)
    assert np.array_equal(rt.bfm3d(g, 1, U, schedule="jacobi").dist, O.bfm3d(nn, lv, X, Y, Z, U, 1)[0])


@pytest.mark.gpu
def test_neighbour_levels_3_is_refused(rt):
    with pytest.raises(rt.RtError) as e:
        rt.grid((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (20, 20, 20), neighbour_levels=3)
    assert e.value.code == 5


@pytest.mark.gpu
def test_struct_recontruct_path_and_result_structs(rt, O):
    """recontruct_path(D, source, receiver) (ssspm.jl:14-28) dispatches on the result struct: chase until a node
    repeats, then append the source.  Checked against the literal transliteration on hand-made tables."""
    #        1  2  3  4  5  6  7  8
    prev = np.array([2, 3, 1, 3, 4, 6, 0, 7], np.int64)  # 1->2->3->1 cycle, 4 and 5 feed it, 6 self loop, 7 unset
    for cls in (rt.BellmanFordMoore, rt.Dijkstra, rt.RadiusStepping):
        D = cls(prev, np.zeros(8))
        assert D[()] is prev
        for rcv in (1, 2, 3, 4, 5, 6):
            assert np.array_equal(rt.recontruct_path(D, 1, rcv), O.reconstruct_path_struct(prev, 1, rcv))
        got = rt.recontruct_path(D, 3, [5, 6, 4])
        assert [list(p) for p in got] == [list(O.reconstruct_path_struct(prev, 3, r)) for r in (5, 6, 4)]
        for rcv in (7, 8):
            with pytest.raises(IndexError):
                O.reconstruct_path_struct(prev, 1, rcv)
            with pytest.raises(rt.RtError) as e:
                rt.recontruct_path(D, 1, rcv)
            assert e.value.code == 4
    assert list(rt.recontruct_path(rt.Dijkstra(prev, None), 1, 5)) == [5, 4, 3, 1, 2, 1]
    # the vector method is unchanged
    assert list(rt.recontruct_path(prev, 1, 5)) == [5, 4, 3, 1]


@pytest.mark.gpu
def test_travel_times_batch_gather(rt, O, annulus, ak135, tmp_path, monkeypatch):
    m = annulus(36, 10, 100.0)
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    halo = m.halo_matrix()
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(*ak135))
    rt.mesh_from_arrays(gr, G, halo)
    srcs = rt.closest_point(gr, np.deg2rad([0.0, 90.0, 200.0]), np.full(3, R), system="polar")
    recv = rt.closest_point(gr, np.deg2rad(np.arange(10.0, 360.0, 25.0)), np.full(14, R), system="polar")
    D = rt.bfm(G, halo, srcs, gr, Vp)
    tt = rt.travel_times(D, gr, recv)
    assert tt.shape == (3, 14) and np.array_equal(tt, D.dist[:, recv - 1])
    D0 = rt.BellmanFordMoore(D.prev[0], D.dist[0])
    monkeypatch.chdir(tmp_path)
    t0 = rt.travel_times(D0, gr, recv, isave=True, flname="tt.csv")
    assert np.array_equal(t0, O.bfm(m, Vp, int(srcs[0]))[0][recv - 1])
    rows = open(tmp_path / "tt.csv").read().strip().split("\n")
    assert rows[0] == "degree,travel_time" and len(rows) == 15
    assert float(rows[1].split(",")[1]) == t0[0]
    with pytest.raises(rt.RtError):
        rt.travel_times(D0, gr, [m.n + 1])
