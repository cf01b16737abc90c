"""Parity tests proper: the CUDA path, called through the C ABI (ctypes front), against the CPU oracle on the
same seeded inputs.  Bar (north star): travel times bit-exact in Float64 (stronger than the 1e-6 relative the
north star allows), predecessors and reconstructed paths bit-exact in the reference (Jacobi) schedule."""
import os

import numpy as np
import pytest

from conftest import splitmix64
from helpers import scan_pairs, ulp_diff, weight2d, weight3d

pytestmark = pytest.mark.gpu
R = 6371.0
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def adopt(rt, m):
    """Oracle-built arrays -> Grid2D / G / halo objects of the product API (== what Julia would hand over)."""
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    halo = m.halo_matrix() if m.halo_rows else None
    return gr, G, halo


def test_library_loaded_and_device_present(rt):
    assert rt.device_count() >= 1


@pytest.mark.parametrize("nt,nr,sp", [(24, 6, 300.0), (36, 10, 100.0), (180, 50, 50.0), (90, 20, 20.0)])
def test_bfm2d_ak135_bit_exact(rt, O, annulus, ak135, nt, nr, sp):
    m = annulus(nt, nr, sp)
    gr, G, halo = adopt(rt, m)
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(*ak135))
    assert np.array_equal(Vp, O.interp_velocity(ak135[0], ak135[1], m.r))  # velocity kernel is bit-exact
    rt.mesh_from_arrays(gr, G, halo)
    src = rt.closest_point(gr, 0.0, R, system="polar")
    assert src == O.closest_point(m.theta, m.r, 0.0, R)
    D = rt.bfm(G, halo, src, gr, Vp)
    dist, prev, st = O.bfm(m, Vp, src)
    assert np.array_equal(D.dist, dist)
    assert np.array_equal(D.prev, prev)
    assert D.stats["sweeps"] == st["sweeps"]
    assert D.stats["graph_edges"] == st["graph_edges"]
    assert D.stats["relaxed_edges"] >= st["relaxed_edges"]  # items relax supersets of the reference frontier
    # README receivers -> identical paths
    degs = np.concatenate([np.arange(10, 151, 10), 360 - np.arange(150, 9, -10)]).astype(np.float32)
    recv = rt.closest_point(gr, np.deg2rad(degs).astype(np.float64), np.full(len(degs), R), system="polar")
    paths = rt.recontruct_path(D.prev, src, recv)
    for k, (rc, p) in enumerate(zip(recv, paths)):
        assert rc == O.closest_point(m.theta, m.r, float(np.deg2rad(degs[k])), R)
        assert np.array_equal(p, O.reconstruct_path(prev, src, int(rc)))


def test_bfm2d_golden_vectors(rt):
    g = np.load(os.path.join(GOLD, "annulus_24_6_300.npz"))
    n, nel, hr = (int(v) for v in g["sizes"])
    gr = rt.Grid2D(g["x"], g["z"], g["theta"], g["r"], g["e2n_off"], g["e2n_idx"], 24, 13, nel, n)
    G = rt.SparseMatrixCSC(nel, n, g["G_colptr"], g["G_rowval"])
    halo = g["halo"].reshape(2, hr).T
    D = rt.bfm(G, halo, int(g["source"]), gr, g["U"])
    assert np.array_equal(D.dist, g["dist"]) and np.array_equal(D.prev, g["prev"])
    assert D.stats["sweeps"] == int(g["sweeps"])
    gm = np.load(os.path.join(GOLD, "annulus_24_6_300_modes.npz"))  # Float32 and dual-velocity modes
    D32 = rt.bfm(G, halo, int(g["source"]), gr, g["U"], precision=32)
    assert np.array_equal(D32.dist, gm["dist_f32"]) and np.array_equal(D32.prev, gm["prev_f32"])
    assert np.array_equal(rt.bfm_gpu(G, halo, int(g["source"]), gr, g["U"], schedule="near-far").dist, gm["dist_f32"])
    rt.bfm(G, halo, 1, gr, g["U"], schedule="jacobi")
    Dd = rt.bfm(G, halo, int(g["source"]), gr, gm["V2"])
    assert np.array_equal(Dd.dist, gm["dist_dual"]) and np.array_equal(Dd.prev, gm["prev_dual"])


def test_bfm2d_random_velocity_and_interior_sources(rt, O, annulus):
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    U = 4.0 + 6.0 * splitmix64(20261018, m.n)  # benchmarks/cpu.jl:19 style stress test
    srcs = np.array([1, m.n, m.n // 2, int(m.halo_matrix()[0, 0]), int(m.halo_matrix()[0, 1]),
                     m.nr * m.ntheta + 1], np.int64)  # first, last (a twin), middle, orig, twin, centre node
    D = rt.bfm(G, halo, srcs, gr, U)
    for k, s in enumerate(srcs):
        dist, prev, st = O.bfm(m, U, int(s))
        assert np.array_equal(D.dist[k], dist), "source %d" % s
        assert np.array_equal(D.prev[k], prev), "source %d" % s
    # size-independent properties on the GPU result itself
    tgt, cand = scan_pairs(m)
    w = weight2d(m.x, m.z, U, tgt, cand)
    d = D.dist[2]
    assert np.all(d[tgt] <= d[cand] + w) and d[srcs[2] - 1] == 0.0
    hm = m.halo_matrix() - 1
    assert np.array_equal(d[hm[:, 0]], d[hm[:, 1]])


def test_bfm2d_no_halo_and_generic_halo(rt, O, annulus):
    m = annulus(24, 6, 300.0)
    # (a) no halo at all: layers stay disconnected below the first discontinuity -> unreached nodes stay Inf
    import copy
    m0 = copy.copy(m)
    m0.halo = np.zeros(0, np.int64)
    m0.halo_rows = 0
    gr, G, _ = adopt(rt, m0)
    U = np.full(m.n, 5.0)
    D = rt.bfm(G, None, 1, gr, U)
    dist, prev, st = O.bfm(m0, U, 1)
    assert np.array_equal(D.dist, dist) and np.array_equal(D.prev, prev) and np.isinf(dist).any()
    # (b) halo rows in a non-standard order -> generic serial halo path, still the reference's serial semantics
    m1 = copy.copy(m)
    hm = m.halo_matrix()[::-1].copy()
    m1.halo = np.ascontiguousarray(hm.T).reshape(-1)
    gr1, G1, halo1 = adopt(rt, m1)
    D1 = rt.bfm(G1, halo1, 7, gr1, U)
    d1, p1, s1 = O.bfm(m1, U, 7)
    assert np.array_equal(D1.dist, d1) and np.array_equal(D1.prev, p1)


def test_bfm_bad_arguments(rt, annulus):
    m = annulus(24, 6, 300.0)
    gr, G, halo = adopt(rt, m)
    U = np.full(m.n, 5.0)
    with pytest.raises(rt.RtError):
        rt.bfm(G, halo, 0, gr, U)
    with pytest.raises(rt.RtError):
        rt.bfm(G, halo, m.n + 1, gr, U)
    with pytest.raises(ValueError):
        rt.bfm(G, halo, 1, gr, U[:-1])
    bad = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval + 10 ** 6)
    gr2 = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    with pytest.raises(rt.RtError):
        rt.bfm(bad, halo, 1, gr2, U)


def test_interp_and_paths_errors(rt, ak135):
    itp = rt.LinearInterpolation(*ak135)
    v = rt.interpolate_velocity(np.array([0.0, 0.25, 3479.5, 6371.0]), itp)
    assert v[0] == ak135[1][0] and v[3] == ak135[1][-1]
    vb = rt.interpolate_velocity(np.array([3479.5]), itp, buffer=1)
    assert vb[0] == rt.interpolate_velocity(np.array([3480.5]), itp)[0]
    with pytest.raises(rt.RtError) as e:
        rt.interpolate_velocity(np.array([6371.0001]), itp)
    assert e.value.code == 3
    prev = np.array([0, 1, 2, 3, 0], np.int64)  # chain 4->3->2->1, node 5 unreachable
    assert list(rt.recontruct_path(prev, 1, 4)) == [4, 3, 2, 1]
    assert list(rt.recontruct_path(prev, 1, 1)) == [1, 1]
    with pytest.raises(rt.RtError) as e:
        rt.recontruct_path(prev, 1, 5)
    assert e.value.code == 4
    assert [list(p) for p in rt.recontruct_path(prev, 1, [2, 4])] == [[2, 1], [4, 3, 2, 1]]


# ------------------------------------------------------------------------------------------------- 3-D
@pytest.mark.parametrize("nn,lv,cs", [((7, 6, 5), 1, "spherical"), ((11, 11, 11), 1, "cartesian"),
                                      ((20, 9, 13), 0, "spherical"), ((33, 18, 10), 2, "cartesian"),
                                      ((40, 40, 24), 1, "spherical")])
def test_bfm3d_bit_exact(rt, O, nn, lv, cs):
    if cs == "spherical":
        c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
        c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    else:
        c0, c1 = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    g = rt.grid(c0, c1, nn, neighbour_levels=lv, coord_system=cs)
    X, Y, Z = g.coordinates()
    Xo, Yo, Zo = O.grid3d_coords(c0, c1, nn, 0 if cs == "cartesian" else 1)
    if cs == "cartesian":
        assert np.array_equal(X, Xo) and np.array_equal(Y, Yo) and np.array_equal(Z, Zo)
    else:  # device sin/cos vs glibc differ by <= 1 ulp each, x = r*cos(phi)*sin(theta) compounds them
        # (builder parity is structural; the solver's bit-exact contract is on identical arrays)
        assert max(ulp_diff(X, Xo), ulp_diff(Y, Yo), ulp_diff(Z, Zo)) <= 4
    n = int(np.prod(nn))
    U = 4.0 + 6.0 * splitmix64(7 + n, n)
    srcs = np.array([1, n, n // 2 + 3], np.int64)
    D = rt.bfm3d(g, srcs, U)
    for k, s in enumerate(srcs):
        dist, prev, st = O.bfm3d(nn, lv, X, Y, Z, U, int(s))
        assert np.array_equal(D.dist[k], dist)
        assert np.array_equal(D.prev[k], prev)
    assert D.stats["graph_edges"] == st["graph_edges"]
    # tightness of every predecessor, independent numpy arithmetic
    d, p = D.dist[0], D.prev[0]
    i = np.nonzero(p > 0)[0]
    assert len(i) == n - 1
    assert np.array_equal(d[p[i] - 1] + weight3d(X, Y, Z, U, i, p[i] - 1), d[i])
    path = rt.recontruct_path(p, 1, n)
    assert path[0] == n and path[-1] == 1 and np.all(np.diff(d[path - 1]) <= 0)


def test_bfm3d_golden(rt):
    g3 = np.load(os.path.join(GOLD, "grid3d_7_6_5.npz"))
    g = rt.grid(g3["c0"], g3["c1"], (7, 6, 5), neighbour_levels=1, coord_system="spherical")
    X, Y, Z = g.coordinates()
    D = rt.bfm3d(g, int(g3["source"]), g3["U"])
    if np.array_equal(X, g3["X"]) and np.array_equal(Y, g3["Y"]) and np.array_equal(Z, g3["Z"]):
        assert np.array_equal(D.dist, g3["dist"]) and np.array_equal(D.prev, g3["prev"])
    else:  # coordinates differ in the last ulp (device vs glibc sin/cos): travel times agree to ~1e-15
        assert np.allclose(D.dist, g3["dist"], rtol=1e-13, atol=0)
    g3f = np.load(os.path.join(GOLD, "grid3d_7_6_5_f32.npz"))
    D32 = rt.bfm3d(g, int(g3["source"]), g3["U"], precision=32)
    f = lambda a: a.astype(np.float32)
    if all(np.array_equal(f(a), f(b)) for a, b in ((X, g3["X"]), (Y, g3["Y"]), (Z, g3["Z"]))):
        assert np.array_equal(D32.dist, g3f["dist_f32"]) and np.array_equal(D32.prev, g3f["prev_f32"])
    else:
        assert np.allclose(D32.dist, g3f["dist_f32"], rtol=1e-6, atol=0)


def test_bfm3d_full_size_properties(rt):
    """At a BASELINE-like size the oracle is too slow: check size-independent properties instead."""
    nn = (96, 96, 64)
    c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
    c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    g = rt.grid(c0, c1, nn, neighbour_levels=1, coord_system="spherical")
    X, Y, Z = g.coordinates()
    n = g.n
    rr = np.sqrt(X * X + Y * Y + Z * Z)
    prof = rt.velocity_profile()
    U = rt.interpolate_velocity(np.minimum(rr, R), rt.LinearInterpolation(prof.r, prof.Vp))
    src = 1 + 48 + 96 * (48 + 96 * 63)
    D = rt.bfm3d(g, src, U)
    d, p = D.dist, D.prev
    assert d[src - 1] == 0.0 and not np.isinf(d).any()
    i = np.nonzero(p > 0)[0]
    assert len(i) == n - 1
    assert np.array_equal(d[p[i] - 1] + weight3d(X, Y, Z, U, i, p[i] - 1), d[i])  # bit-exact tightness
    # fixed point along the three axes (+-1, +-2 neighbours)
    D3 = d.reshape(64, 96, 96)
    idx = np.arange(n).reshape(64, 96, 96)
    for ax in range(3):
        for sft in (1, 2):
            a = np.take(idx, np.arange(sft, idx.shape[ax]), axis=ax).ravel()
            b = np.take(idx, np.arange(0, idx.shape[ax] - sft), axis=ax).ravel()
            w = weight3d(X, Y, Z, U, a, b)
            assert np.all(d[a] <= d[b] + w) and np.all(d[b] <= d[a] + w)
    assert D.stats["relaxed_edges"] > D.stats["graph_edges"]


# --------------------------------------------------------------------------------- device annulus builder
@pytest.mark.parametrize("nt,nr,sp", [(8, 2, 500.0), (24, 6, 300.0), (37, 11, 77.7), (180, 50, 20), (64, 3, 0.9)])
def test_init_annulus_device_builder(rt, O, annulus, nt, nr, sp):
    m = annulus(nt, nr, sp)
    gr, G, halo = rt.init_annulus(nt, nr, spacing=sp)
    assert (gr.nnods, gr.nel, gr.ntheta, gr.nr) == (m.n, m.nel, m.ntheta, m.nr)
    assert np.array_equal(gr.theta, m.theta) and np.array_equal(gr.r, m.r)  # plain arithmetic: bit-identical
    assert max(ulp_diff(gr.x, m.x), ulp_diff(gr.z, m.z)) <= 2 or np.allclose(gr.x, m.x, rtol=0, atol=2e-12)
    assert np.array_equal(gr.e2n_off, m.e2n_off) and np.array_equal(gr.e2n_idx, m.e2n_idx)
    assert np.array_equal(G.colptr, m.G_colptr) and np.array_equal(G.rowval, m.G_rowval)
    assert np.array_equal(halo, m.halo_matrix())
    assert np.array_equal(gr.nbr_off, m.nbr_off) and np.array_equal(gr.element_type, m.el_type)
    for e in range(1, m.nel + 1, max(1, m.nel // 300)):
        assert sorted(m.nbr_idx[m.nbr_off[e - 1]:m.nbr_off[e]]) == list(gr.neighbours(e))
    assert np.array_equal(gr.e2n[m.nel], m.e2n_idx[m.e2n_off[m.nel - 1]:m.e2n_off[m.nel]])


def test_readme_example_on_device_built_mesh(rt, O, ak135):
    """README.md:15-53 end to end through the drop-in API, checked against the oracle on identical arrays."""
    nt, nr, sp = 180, 50, 50
    gr, G, halo = rt.init_annulus(nt, nr, spacing=sp)
    source = rt.closest_point(gr, 0.0, R, system="polar")
    profile = rt.velocity_profile()
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(profile.r, profile.Vp))
    D = rt.bfm(G, halo, source, gr, Vp)
    # oracle on the arrays the device builder produced (x, z come from device sin/cos)
    m = O.Annulus(nt, nr, sp)
    m.x, m.z = gr.x, gr.z
    dist, prev, st = O.bfm(m, Vp, source)
    assert np.array_equal(D.dist, dist) and np.array_equal(D.prev, prev) and D.stats["sweeps"] == st["sweeps"]
    degs = np.concatenate([np.arange(10, 151, 10), 360 - np.arange(150, 9, -10)]).astype(np.float32)
    recv = rt.closest_point(gr, np.deg2rad(degs).astype(np.float64), np.full(len(degs), R), system="polar")
    paths = rt.recontruct_path(D.prev, source, recv)
    assert len(paths) == 30
    for rc, p in zip(recv, paths):
        assert np.array_equal(p, O.reconstruct_path(prev, source, int(rc)))
    for deg, t in ((30, 376.49), (90, 785.97), (180, 1189.47)):  # SURVEY 8c-4 sanity values
        rc = rt.closest_point(gr, float(np.deg2rad(np.float32(deg))), R, system="polar")
        assert abs(D.dist[rc - 1] - t) < 0.006


def test_init_annulus_rejects_bad_arguments(rt):
    for args in ((4, 10, 20.0), (36, 1, 20.0), (36, 10, 0.0)):
        with pytest.raises(rt.RtError):
            rt.init_annulus(args[0], args[1], spacing=args[2])


# ------------------------------------------------------------------------- near-far (work-efficient) schedule
def check_prev_tie_aware(m, U, src, dist, prev_new, prev_ref, weight=None):
    """dist must be bit-identical; where prev differs from the reference it must still be a
    bit-exactly tight predecessor (an exact tie), or a zero-weight coupling (halo twin / coincident node)."""
    n = m.n
    idx = np.arange(n)
    p = prev_new - 1
    reached = np.isfinite(dist) & (idx != src - 1)
    assert np.all(p[reached] >= 0), "reached node without predecessor"
    i = idx[reached]
    j = p[reached]
    w = weight(i, j) if weight else weight2d(m.x, m.z, U, i, j)
    tight = dist[j] + w == dist[i]
    hm = m.halo_matrix() - 1 if m.halo_rows else np.zeros((0, 2), np.int64)
    partner_of = {}
    for a, b in hm:
        partner_of.setdefault(int(b), []).append(int(a))
    bad = []
    for q in np.nonzero(~tight)[0]:
        node = int(i[q])
        # halo copy: inherits the predecessor of an equal-time twin (update_halo!)
        ok = any(dist[t] == dist[node] and (prev_new[t] == prev_new[node] or t == src - 1 and prev_new[node] == src)
                 for t in partner_of.get(node, []))
        if not ok:
            bad.append(node)
    assert not bad, "non-tight predecessors at nodes %s" % bad[:10]
    assert np.all(dist[j] <= dist[i])
    same = int((prev_new == prev_ref).sum())
    return same / n


@pytest.mark.parametrize("nt,nr,sp", [(24, 6, 300.0), (36, 10, 100.0), (180, 50, 50.0)])
def test_bfm2d_near_far_schedule(rt, O, annulus, ak135, nt, nr, sp):
    m = annulus(nt, nr, sp)
    gr, G, halo = adopt(rt, m)
    Vp = O.interp_velocity(ak135[0], ak135[1], m.r)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    dist, prev, st = O.bfm(m, Vp, src)
    for delta in (None, 0.5, 50.0):
        D = rt.bfm(G, halo, src, gr, Vp, schedule="near-far", delta=delta if delta else 0.0)
        assert np.array_equal(D.dist, dist), "travel times must be bit-identical in any schedule"
        frac = check_prev_tie_aware(m, Vp, src, dist, D.prev, prev)
        assert frac > 0.3
        # every reconstructed path reaches the source with non-increasing travel time
        degs = np.arange(10, 351, 20).astype(np.float32)
        for deg in degs:
            rc = O.closest_point(m.theta, m.r, float(np.deg2rad(deg)), R)
            path = rt.recontruct_path(D.prev, src, rc)
            assert path[-1] == src and np.all(np.diff(dist[path - 1]) <= 0)
        assert D.stats["relaxed_edges"] < st["relaxed_edges"]  # work-efficient
    rt.bfm(G, halo, src, gr, Vp, schedule="jacobi")  # leave the cached handle in the default schedule


def test_bfm2d_near_far_random_velocity_sources(rt, O, annulus):
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    U = 4.0 + 6.0 * splitmix64(99, m.n)
    srcs = np.array([1, m.n, m.n // 2, int(m.halo_matrix()[0, 0]), m.nr * m.ntheta + 1], np.int64)
    D = rt.bfm(G, halo, srcs, gr, U, schedule="near-far")
    for k, s in enumerate(srcs):
        dist, prev, st = O.bfm(m, U, int(s))
        assert np.array_equal(D.dist[k], dist), "source %d" % s
        check_prev_tie_aware(m, U, int(s), dist, D.prev[k], prev)
    rt.bfm(G, halo, 1, gr, U, schedule="jacobi")


def test_interpolate_cells_matches_oracle(rt, O, annulus, ak135):
    """interpolate!(V, gr) (src/Interpolations/*.jl) through the ABI, on adopted and on device-built meshes."""
    m = annulus(36, 10, 100.0)
    U = O.interp_velocity(ak135[0], ak135[1], m.r)
    want = O.interpolate_cells(m, U)
    gr, G, halo = adopt(rt, m)
    gr.element_type = m.el_type
    got = rt.interpolate_inplace(U.copy(), gr, G, halo)
    assert np.array_equal(got, want)
    gr2, G2, halo2 = rt.init_annulus(36, 10, spacing=100.0)
    V = rt.interpolate_velocity(gr2.r, rt.LinearInterpolation(*ak135))
    assert np.array_equal(rt.interpolate_inplace(V, gr2), want)
    # benchmarks/gpu.jl:54-58 flow: interpolant -> interpolate! -> bfm still solves and matches the oracle
    src = rt.closest_point(gr2, 0.0, R, system="polar")
    D = rt.bfm(G2, halo2, src, gr2, V)
    mm = O.Annulus(36, 10, 100.0)
    mm.x, mm.z = gr2.x, gr2.z
    dist, prev, st = O.bfm(mm, V, src)
    assert np.array_equal(D.dist, dist) and np.array_equal(D.prev, prev)


@pytest.mark.parametrize("nn,lv,cs", [((7, 6, 5), 1, "spherical"), ((40, 40, 24), 1, "spherical"),
                                      ((33, 18, 10), 2, "cartesian"), ((70, 9, 13), 0, "spherical")])
def test_bfm3d_near_far_schedule(rt, O, nn, lv, cs):
    if cs == "spherical":
        c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
        c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    else:
        c0, c1 = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    g = rt.grid(c0, c1, nn, neighbour_levels=lv, coord_system=cs)
    X, Y, Z = g.coordinates()
    n = int(np.prod(nn))
    U = 4.0 + 6.0 * splitmix64(11 + n, n)
    srcs = np.array([1, n, n // 2 + 3], np.int64)
    ref = [O.bfm3d(nn, lv, X, Y, Z, U, int(s)) for s in srcs]
    # tile_pull = 1 (default): targets pull from the released sources tile by tile; 0: push units per released x-line
    # early_advance > 0: the threshold moves on while stragglers of the bucket are still being released
    # batch: 0 = the sources of the call run concurrently in per-source slots (tile-pull), 1 = one after the other
    for tile_pull, delta, early, batch in ((1, 0.0, 0.0, 0), (1, 1e-3, 0.0, 2), (1, 1e9, 0.0, 0), (1, 0.0, 2.0, 1),
                                           (1, 0.0, 0.0, 1), (0, 0.0, 0.0, 0), (0, 1e9, 0.0, 0)):
        g._handle.set_option("tile_pull", tile_pull)
        g._handle.set_option("early_advance", early)
        g._handle.set_option("batch", batch)
        D = rt.bfm3d(g, srcs, U, schedule="near-far", delta=delta)
        for k, s in enumerate(srcs):
            dist, prev, st = ref[k]
            assert np.array_equal(D.dist[k], dist), "travel times must be bit-identical in any schedule"
            p = D.prev[k]
            i = np.nonzero(p > 0)[0]
            assert len(i) == n - 1 and p[s - 1] == 0
            assert np.array_equal(dist[p[i] - 1] + weight3d(X, Y, Z, U, i, p[i] - 1), dist[i])  # tight
            assert np.all(dist[p[i] - 1] < dist[i])
            path = rt.recontruct_path(p, int(s), 1 if s != 1 else n)
            assert path[-1] == s and np.all(np.diff(dist[path - 1]) <= 0)
    assert D.stats["relaxed_edges"] > 0
    g._handle.set_option("tile_pull", 1)
    g._handle.set_option("early_advance", -1)
    g._handle.set_option("batch", 0)
    rt.bfm3d(g, 1, U, schedule="jacobi")


def test_annulus_schedules_agree_at_scale(rt):
    """Size-independent properties on a mesh the oracle is too slow for (180x50 @5 km, 0.76 M nodes, E_graph
    1.2e9): both schedules give bit-identical travel times; near-far predecessors are bit-exactly tight."""
    gr, G, halo = rt.init_annulus(180, 50, spacing=5.0)
    prof = rt.velocity_profile()
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(prof.r, prof.Vp))
    src = rt.closest_point(gr, 0.0, R, system="polar")
    Dj = rt.bfm(G, halo, src, gr, Vp, schedule="jacobi")
    Dn = rt.bfm(G, halo, src, gr, Vp, schedule="near-far")
    assert np.array_equal(Dj.dist, Dn.dist)
    assert Dn.stats["relaxed_edges"] * 5 < Dj.stats["relaxed_edges"]
    d, p = Dn.dist, Dn.prev
    assert d[src - 1] == 0.0 and np.isfinite(d).all()
    assert np.array_equal(d[halo[:, 0] - 1], d[halo[:, 1] - 1])  # twins
    i = np.nonzero((p > 0))[0]
    tight = d[p[i] - 1] + weight2d(gr.x, gr.z, Vp, i, p[i] - 1) == d[i]
    halo_nodes = np.zeros(len(d), bool)
    halo_nodes[halo.ravel() - 1] = True
    assert np.all(tight | halo_nodes[i])  # non-tight only where the value came through a zero-weight twin coupling
    assert tight.mean() > 0.85  # ~13 % of the nodes of this mesh are halo twins
    recv = rt.closest_point(gr, np.deg2rad(np.arange(10.0, 351.0, 10.0)), np.full(35, R), system="polar")
    for path_j, path_n in zip(rt.recontruct_path(Dj.prev, src, recv), rt.recontruct_path(Dn.prev, src, recv)):
        assert path_j[-1] == src and path_n[-1] == src
        assert np.all(np.diff(d[path_n - 1]) <= 0)
        # the two paths may differ in node ids on exact ties, never in travel time along them
        assert d[path_j[0] - 1] == d[path_n[0] - 1]


def test_config2_full_size_properties(rt):
    """BASELINE config[1] at full size (annulus 1440x400 @0.25 km, 106.6 M nodes; the reference cannot build it):
    properties that need no oracle."""
    import ctypes as C
    import torch
    gr, G, halo = rt.init_annulus(1440, 400, spacing=0.25, export=False)
    n = gr.nnods
    assert n == 106619041  # ~106.6 M
    h = gr._handle
    prof = rt.velocity_profile()
    itp = rt.LinearInterpolation(prof.r, prof.Vp)
    x_d, z_d, th_d, r_d = h.coords_dev()
    U = torch.empty(n, dtype=torch.float64, device="cuda")
    rt.api.check(rt.lib().rt_interp_velocity_dev(itp.knots, itp.values, len(itp.knots), r_d, n, -1.0, U.data_ptr()))
    src = rt.closest_point(gr, 0.0, R, system="polar")
    h.set_option("schedule", 1)
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    p = torch.empty(n, dtype=torch.int32, device="cuda")
    st = rt.RtStats()
    rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U.data_ptr(), np.array([src], np.int64), 1, 64, d.data_ptr(),
                                           p.data_ptr(), C.byref(st)))
    assert float(d[src - 1]) == 0.0 and bool(torch.isfinite(d).all())
    assert int((p < 0).sum()) == 1  # only the source has no predecessor
    nrp, nt = gr.nr, gr.ntheta
    ring = d[:nrp * nt].reshape(nt, nrp)
    mirror = torch.flip(ring[1:], dims=[0])
    assert float(((ring[1:] - mirror).abs() / ring[1:].clamp_min(1e-9)).max()) < 1e-9  # theta <-> 2 pi - theta
    t180 = float(ring[nt // 2, nrp - 1])
    assert 1150.0 < t180 < 1212.08  # CMB-diffracted first arrival, below the straight-through bound
    assert st.relaxed_edges < 3 * st.graph_edges
    rec = rt.closest_point(gr, np.deg2rad(np.array([30.0, 90.0, 180.0])), np.full(3, R), system="polar")
    off = np.zeros(4, np.int64)
    rt.api.check(rt.lib().rt_reconstruct_paths_dev(p.data_ptr(), n, src, rec, 3, off, None, 0))
    assert np.all(np.diff(off) > 10)


def test_bfm2d_near_far_batched_sources(rt, O, annulus, ak135):
    """Many earthquakes on one mesh: sources advance in lock step inside the kernels (batch API); every table must
    equal the single-source solve bit for bit."""
    m = annulus(90, 20, 20.0)
    gr, G, halo = adopt(rt, m)
    Vp = O.interp_velocity(ak135[0], ak135[1], m.r)
    k = np.arange(37)
    srcs = np.array([O.closest_point(m.theta, m.r, 2 * np.pi * kk / 37.0, R) for kk in k] + [1, m.n, 5], np.int64)
    D = rt.bfm(G, halo, srcs, gr, Vp, schedule="near-far")
    assert D.dist.shape == (40, m.n)
    for idx in (0, 17, 31, 32, 39):  # both chunks (32 + 8), first/last entries
        d1, p1, _ = O.bfm(m, Vp, int(srcs[idx]))
        assert np.array_equal(D.dist[idx], d1), "source #%d" % idx
        check_prev_tie_aware(m, Vp, int(srcs[idx]), d1, D.prev[idx], p1)
    single = rt.bfm(G, halo, int(srcs[17]), gr, Vp, schedule="near-far")
    assert np.array_equal(single.dist, D.dist[17])
    # predecessors are schedule independent wherever a regular (positive-weight) tight predecessor exists; only
    # zero-weight alternatives (coincident duplicates, halo twins) may resolve differently between runs
    diff = np.nonzero(single.prev != D.prev[17])[0]
    halo_nodes = set(int(v) - 1 for v in halo.ravel())
    d = single.dist
    for i in diff:
        a, b = single.prev[i] - 1, D.prev[17][i] - 1
        assert int(i) in halo_nodes or d[a] == d[i] or d[b] == d[i], "regular predecessor differs at node %d" % (i + 1)
    assert len(diff) < 0.1 * m.n
    rt.bfm(G, halo, 1, gr, Vp, schedule="jacobi")


def test_edge_cases_empty_and_tiny(rt, O, annulus):
    """Empty source batch, a single-node-per-axis 3-D grid, many receivers at once."""
    import ctypes as C
    m = annulus(24, 6, 300.0)
    gr, G, halo = adopt(rt, m)
    U = np.full(m.n, 5.0)
    h = rt.mesh_from_arrays(gr, G, halo)
    st = rt.RtStats()
    rt.api.check(rt.lib().rt_bfm_solve(h.h, U, np.zeros(1, np.int64), 0, 64, None, None, C.byref(st)))  # nsrc = 0
    assert st.sweeps == 0
    D = rt.bfm(G, halo, 1, gr, U)
    recv = np.arange(2, 400, dtype=np.int64)
    paths = rt.recontruct_path(D.prev, 1, recv)
    assert len(paths) == len(recv) and all(p[0] == r and p[-1] == 1 for p, r in zip(paths, recv))
    g = rt.grid((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (5, 1, 1), neighbour_levels=1)  # degenerate axes
    X, Y, Z = g.coordinates()
    D3 = rt.bfm3d(g, 1, np.ones(5))
    d3, p3, _ = O.bfm3d((5, 1, 1), 1, X, Y, Z, np.ones(5), 1)
    assert np.array_equal(D3.dist, d3) and np.array_equal(D3.prev, p3)
    assert np.array_equal(rt.bfm3d(g, 1, np.ones(5), schedule="near-far").dist, d3)


def test_topology_containers_and_rcm(rt, O, annulus, ak135):
    """nodal_incidence / nodal_degree / sparse_adjacency_list (src/topology/topology.jl) and symrcm + reorder!
    (src/SSSP/rcm.jl) on the device."""
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    rt.mesh_from_arrays(gr, G, halo)
    deg_o, off_o, lst_o = O.nodal_adjacency(m)
    assert np.array_equal(rt.nodal_degree(gr), deg_o)
    A = rt.sparse_adjacency_list(gr)
    assert np.array_equal(A.deg, deg_o) and np.array_equal(A.idx, off_o[:-1] + 1)
    for v in (1, 7, m.nr * m.ntheta + 1, m.n // 2, m.n):  # same neighbour SETS (the reference's order is a Set's)
        assert sorted(A.neighbours(v)) == list(lst_o[off_o[v - 1]:off_o[v]])
    assert np.array_equal(np.sort(A.list[:off_o[50]].reshape(-1)), np.sort(lst_o[:off_o[50]]))
    # symrcm: a permutation; every node but the component seeds has a neighbour placed before it in BFS order
    prm = rt.symrcm(gr)
    assert sorted(prm) == list(range(1, m.n + 1))
    F = prm[::-1]  # Cuthill-McKee order before the reversal
    posF = np.zeros(m.n + 1, np.int64)
    posF[F] = np.arange(m.n)
    seeds = 0
    for v in F:
        nb = lst_o[off_o[v - 1]:off_o[v]]
        if len(nb) == 0 or posF[nb].min() > posF[v]:
            seeds += 1
    assert 1 <= seeds <= 16  # one seed per connected component (8 velocity layers, twins are separate nodes)
    assert deg_o[F[0] - 1] == deg_o.min()  # starts from a minimum-degree node (rcm.jl:4,13-22)
    # dense padded container (topology.jl:1-4, 52-68) and element_degree (:79-86)
    sp = rt.sparse_adjacency_list(gr)
    A = rt.adjacency_list(gr, G, halo)
    assert A.G.shape == (int(deg_o.max()), m.n) and np.array_equal(A.N, deg_o)
    for node in (1, 7, m.n // 2, m.n):
        col = A.G[:, node - 1]
        assert np.array_equal(col[:deg_o[node - 1]], sp.neighbours(node)) and not col[deg_o[node - 1]:].any()
    assert np.array_equal(rt.element_degree(G), np.diff(m.G_colptr))
    # reorder! + solve: travel times are the same function of the (relabelled) nodes
    Vp = O.interp_velocity(ak135[0], ak135[1], m.r)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    D = rt.bfm(G, halo, src, gr, Vp)
    gr2, G2, halo2 = rt.reorder(gr, G, halo, prm)
    inv = np.zeros(m.n + 1, np.int64)
    inv[prm] = np.arange(1, m.n + 1)
    D2 = rt.bfm(G2, halo2, int(inv[src]), gr2, Vp[prm - 1])
    assert np.array_equal(D2.dist, D.dist[prm - 1])
    # bandwidth of the element lists (max id spread inside a cell) does not explode under RCM
    spread = lambda g: max(int(g.e2n[e].max() - g.e2n[e].min()) for e in range(1, m.nel + 1, 7))
    assert spread(gr2) <= 2 * spread(gr)


def test_dual_velocity_relax_bit_exact(rt, O, annulus, ak135):
    """bfm with U::Matrix (dual_velocity, discontinuity-aware relax) in the reference schedule."""
    for nt, nr, sp in ((24, 6, 300.0), (36, 10, 100.0), (90, 20, 20.0)):
        m = annulus(nt, nr, sp)
        gr, G, halo = adopt(rt, m)
        itp = rt.LinearInterpolation(*ak135)
        V2 = rt.dual_velocity(gr.r, itp, buffer=1)
        assert np.array_equal(V2, O.dual_velocity(ak135[0], ak135[1], m.r, 1.0))
        src = O.closest_point(m.theta, m.r, 0.0, R)
        D = rt.bfm(G, halo, src, gr, V2)
        dist, prev, st = O.bfm_dual(m, V2, src)
        assert np.array_equal(D.dist, dist) and np.array_equal(D.prev, prev) and D.stats["sweeps"] == st["sweeps"]
    V2r = np.stack([4.0 + 6.0 * splitmix64(5, m.n), 4.0 + 6.0 * splitmix64(6, m.n)], 1)  # arbitrary pairs
    D = rt.bfm(G, halo, [1, m.n // 3], gr, V2r)
    for k, s in enumerate((1, m.n // 3)):
        dist, prev, st = O.bfm_dual(m, V2r, s)
        assert np.array_equal(D.dist[k], dist) and np.array_equal(D.prev[k], prev)
    with pytest.raises(ValueError):
        rt.bfm(G, halo, 1, gr, V2r[:-1])
    # near-far schedule with the dual relax: same travel times bit for bit, tight predecessors
    def wdual(i, j):  # weight seen by target i pulling from j (bfm.jl:137-145)
        down = m.r[i] > m.r[j]
        us = np.where(down, V2r[i, 0], V2r[i, 1]) + np.where(down, V2r[j, 1], V2r[j, 0])
        dx, dz = m.x[i] - m.x[j], m.z[i] - m.z[j]
        return 2.0 * np.sqrt(dx * dx + dz * dz) / us
    for s in (1, m.n // 3):
        dist, prev, st = O.bfm_dual(m, V2r, s)
        Dn = rt.bfm(G, halo, s, gr, V2r, schedule="near-far")
        assert np.array_equal(Dn.dist, dist)
        check_prev_tie_aware(m, None, s, dist, Dn.prev, prev, weight=wdual)
    rt.bfm(G, halo, 1, gr, V2r, schedule="jacobi")


# ------------------------------------------------- BASELINE configs[2], [3], [4] at full size (device-side checks)
def _solve_dev(rt, h, n, U_dev, srcs):
    import ctypes as C
    import torch
    srcs = np.ascontiguousarray(srcs, np.int64)
    d = torch.empty(len(srcs) * n, dtype=torch.float64, device="cuda")
    p = torch.empty(len(srcs) * n, dtype=torch.int32, device="cuda")
    st = rt.RtStats()
    rt.api.check(rt.lib().rt_bfm_solve_dev(h.h, U_dev.data_ptr(), srcs, len(srcs), 64, d.data_ptr(), p.data_ptr(),
                                           C.byref(st)))
    return d.view(len(srcs), n), p.view(len(srcs), n), st


def test_config3_batch_512_sources_properties(rt):
    """BASELINE config[2]: annulus 720x200 (591 121 nodes), 512 earthquakes, full travel-time tables resident on
    the device.  Checked with torch fp64 elementwise ops (no contraction, IEEE sqrt/div): source rows, twin
    equality, bit-exact tightness of every regular predecessor, reciprocity T_a(b) ~ T_b(a), and one table against
    the reference schedule."""
    import torch
    gr, G, halo = rt.init_annulus(720, 200)
    n = gr.nnods
    assert n == 591121
    h = gr._handle
    prof = rt.velocity_profile()
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(prof.r, prof.Vp))
    U = torch.from_numpy(Vp).cuda()
    k = np.arange(512)
    srcs = rt.closest_point(gr, 2 * np.pi * k / 512.0, np.full(512, R), "polar")
    h.set_option("schedule", 1)
    d, p, st = _solve_dev(rt, h, n, U, srcs)
    torch.cuda.synchronize()
    s_t = torch.from_numpy(srcs - 1).cuda()
    rows = torch.arange(512, device="cuda")
    assert bool((d[rows, s_t] == 0).all()) and bool(torch.isfinite(d).all())
    assert int((p < 0).sum()) == 512  # only the sources have no predecessor
    hl = torch.from_numpy(halo - 1).cuda()
    assert bool((d[:, hl[:, 0]] == d[:, hl[:, 1]]).all())
    # reciprocity: the graph is undirected, so T_a(b) and T_b(a) are the same path summed in opposite order
    T = d[:, s_t]
    assert float(((T - T.T).abs() / T.clamp_min(1.0)).max()) < 1e-12
    x, z = torch.from_numpy(gr.x).cuda(), torch.from_numpy(gr.z).cuda()
    is_halo = torch.zeros(n, dtype=torch.bool, device="cuda")
    is_halo[hl.reshape(-1)] = True
    idx = torch.arange(n, device="cuda")
    for row in (0, 129, 511):
        pr = p[row].long()
        ok = pr >= 0
        i, j = idx[ok], pr[ok]
        dx, dz = x[i] - x[j], z[i] - z[j]
        w = 2.0 * torch.sqrt(dx * dx + dz * dz) / (U[i] + U[j])
        tight = d[row, j] + w == d[row, i]
        assert bool((tight | is_halo[i]).all()) and float(tight.double().mean()) > 0.9
    h.set_option("schedule", 0)
    dj, pj, stj = _solve_dev(rt, h, n, U, srcs[129:130])
    assert bool((dj[0] == d[129]).all())
    assert st.relaxed_edges < 4 * 512 * st.graph_edges  # graph_edges is per source


def _shell(rt, nn):
    import torch
    c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
    c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    g = rt.grid(c0, c1, nn, neighbour_levels=1, coord_system="spherical")
    X, Y, Z = g.coordinates()
    rr = np.minimum(np.sqrt(X * X + Y * Y + Z * Z), R)
    prof = rt.velocity_profile()
    U = rt.interpolate_velocity(rr, rt.LinearInterpolation(prof.r, prof.Vp))
    return g, [torch.from_numpy(a).cuda() for a in (X, Y, Z, U)]


def _check_tight3d(d, p, X, Y, Z, U, chunk=1 << 24):
    import torch
    n = d.numel()
    for a in range(0, n, chunk):
        i = torch.arange(a, min(n, a + chunk), device="cuda")
        j = p[i].long()
        ok = j >= 0
        i, j = i[ok], j[ok]
        dx, dy, dz = X[i] - X[j], Y[i] - Y[j], Z[i] - Z[j]
        w = torch.sqrt(dx * dx + dy * dy + dz * dz) * (1.0 / (U[i] + U[j]).abs()) * 2.0
        assert bool((d[j] + w == d[i]).all())


def test_config4_grid3d_216_properties(rt):
    """BASELINE config[3]: 3-D spherical shell 216^3 (10 077 696 nodes), star-1, single source.  Both schedules
    give the same table bit for bit; every predecessor is bit-exactly tight (weights.jl:20 operation order)."""
    g, (X, Y, Z, U) = _shell(rt, (216, 216, 216))
    n = g.n
    assert n == 10077696
    src = 1 + 108 + 216 * (108 + 216 * 215)
    h = g._handle
    h.set_option("schedule", 0)
    dj, pj, stj = _solve_dev(rt, h, n, U, [src])
    h.set_option("schedule", 1)
    dn, pn, stn = _solve_dev(rt, h, n, U, [src])
    h.set_option("schedule", 0)
    assert bool((dj == dn).all()) and float(dj[0, src - 1]) == 0.0 and bool(torch_isfinite_all(dj))
    for d, p in ((dj, pj), (dn, pn)):
        assert int((p < 0).sum()) == 1
        _check_tight3d(d[0], p[0], X, Y, Z, U)
    assert stn.relaxed_edges < stj.relaxed_edges


def torch_isfinite_all(t):
    import torch
    return torch.isfinite(t).all()


def test_config5_grid3d_368_batch_and_receiver_sweep(rt):
    """BASELINE config[4]: 3-D shell 368^3 (49 836 032 nodes), a batch of surface sources (8 of the 64: the
    per-source work is identical) and a 1000-receiver recontruct_path sweep on device-resident predecessors."""
    import torch
    nx = 368
    g, (X, Y, Z, U) = _shell(rt, (nx, nx, nx))
    n = g.n
    assert n == 49836032
    srcs = np.array([1 + (nx * (2 * a + 1)) // 8 + nx * ((nx * (2 * b + 1)) // 4 + nx * (nx - 1))
                     for a in range(4) for b in range(2)], np.int64)
    h = g._handle
    h.set_option("schedule", 1)
    d, p, st = _solve_dev(rt, h, n, U, srcs)
    h.set_option("schedule", 0)
    rows = torch.arange(len(srcs), device="cuda")
    assert bool((d[rows, torch.from_numpy(srcs - 1).cuda()] == 0).all()) and bool(torch.isfinite(d).all())
    assert int((p < 0).sum()) == len(srcs)
    _check_tight3d(d[0], p[0], X, Y, Z, U)
    _check_tight3d(d[7], p[7], X, Y, Z, U)
    # 1000 receivers on the bottom and top faces
    rng = np.random.default_rng(7)
    rec = np.concatenate([1 + rng.integers(0, nx * nx, 500), 1 + n - nx * nx + rng.integers(0, nx * nx, 500)])
    rec = rec[rec != srcs[3]].astype(np.int64)
    off = np.zeros(len(rec) + 1, np.int64)
    lib = rt.lib()
    rt.api.check(lib.rt_reconstruct_paths_dev(p[3].data_ptr(), n, int(srcs[3]), rec, len(rec), off, None, 0))
    idx = np.zeros(int(off[-1]), np.int64)
    rt.api.check(lib.rt_reconstruct_paths_dev(p[3].data_ptr(), n, int(srcs[3]), rec, len(rec), off,
                                              rt.api.ptr(idx), len(idx)))
    d3 = d[3].cpu().numpy()
    for kk in range(0, len(rec), 37):
        path = idx[off[kk]:off[kk + 1]]
        assert path[0] == rec[kk] and path[-1] == srcs[3] and np.all(np.diff(d3[path - 1]) < 0)
    assert np.all(idx[off[1:] - 1] == srcs[3]) and np.all(idx[off[:-1]] == rec)


# ------------------------------------------------------------------ precision = 32 (Float32 path of bfm_gpu.jl)
@pytest.mark.parametrize("nt,nr,sp", [(24, 6, 300.0), (36, 10, 100.0), (90, 20, 20.0)])
def test_bfm2d_float32_mode(rt, O, annulus, ak135, nt, nr, sp):
    """Float32 comparison path (src/SSSP/bfm_gpu.jl:170-205, 487-526).  x, z, U are cast to Float32 and
    every operation is rounded to Float32; the result must equal genuine Float32 arithmetic (the oracle computes
    in `float`) bit for bit in both schedules, predecessors included in the reference schedule, and stay within
    the Float32 accuracy of the Float64 tables."""
    m = annulus(nt, nr, sp)
    gr, G, halo = adopt(rt, m)
    Vp = O.interp_velocity(ak135[0], ak135[1], m.r)
    srcs = [O.closest_point(m.theta, m.r, 0.0, R), m.n // 2]
    for src in srcs:
        d32, p32, st32 = O.bfm_f32(m, Vp, src)
        D = rt.bfm(G, halo, src, gr, Vp, schedule="jacobi", precision=32)
        assert D.dist.dtype == np.float32
        assert np.array_equal(D.dist.astype(np.float64), d32)
        assert np.array_equal(D.prev, p32) and D.stats["sweeps"] == st32["sweeps"]
        Dn = rt.bfm_gpu(G, halo, src, gr, Vp, schedule="near-far")
        assert np.array_equal(Dn.dist.astype(np.float64), d32)
        f32 = lambda a: a.astype(np.float32)

        def w32(i, j):  # Float32 weight, numpy float32 ops round after every operation
            dx, dz = f32(m.x)[i] - f32(m.x)[j], f32(m.z)[i] - f32(m.z)[j]
            return (np.float32(2) * np.sqrt(dx * dx + dz * dz) / (f32(Vp)[i] + f32(Vp)[j])).astype(np.float64)

        def wsum(i, j):  # tightness in Float32: fl32(d[j] + w)
            return w32(i, j)
        dref = d32.astype(np.float32)
        p = Dn.prev - 1
        idx = np.arange(m.n)
        ok = np.isfinite(d32) & (idx != src - 1)
        i, j = idx[ok], p[ok]
        tight = (dref[j] + w32(i, j).astype(np.float32)) == dref[i]
        halo_nodes = np.zeros(m.n, bool)
        halo_nodes[halo.ravel() - 1] = True
        assert np.all(tight | halo_nodes[i])
        d64, _, _ = O.bfm(m, Vp, src)
        assert np.max(np.abs(d32 - d64) / np.maximum(d64, 1e-9)) < 5e-5  # Float32 accumulates ~1e-5 on these meshes
    # batch of sources in lock step + back to Float64 on the same handle (the rounded copies must not leak)
    D = rt.bfm(G, halo, srcs, gr, Vp, schedule="near-far", precision=32)
    for k, src in enumerate(srcs):
        assert np.array_equal(D.dist[k].astype(np.float64), O.bfm_f32(m, Vp, src)[0])
    D64 = rt.bfm(G, halo, srcs[0], gr, Vp, schedule="jacobi")
    assert np.array_equal(D64.dist, O.bfm(m, Vp, srcs[0])[0])
    with pytest.raises(ValueError):
        rt.bfm(G, halo, 1, gr, Vp, precision=16)


@pytest.mark.parametrize("nn,cs", [((9, 8, 7), "cartesian"), ((24, 20, 16), "spherical")])
def test_bfm3d_float32_mode(rt, O, nn, cs):
    if cs == "spherical":
        c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
        c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    else:
        c0, c1 = (0.0, 0.0, 0.0), (90.0, 70.0, 60.0)
    g = rt.grid(c0, c1, nn, neighbour_levels=1, coord_system=cs)
    X, Y, Z = g.coordinates()
    U = 4.0 + 6.0 * splitmix64(11, g.n)
    for src in (1, g.n // 2):
        d32, p32, st = O.bfm3d_f32(nn, 1, X, Y, Z, U, src)
        D = rt.bfm3d(g, src, U, schedule="jacobi", precision=32)
        assert np.array_equal(D.dist.astype(np.float64), d32) and np.array_equal(D.prev, p32)
        Dn = rt.bfm3d(g, src, U, schedule="near-far", precision=32)
        assert np.array_equal(Dn.dist.astype(np.float64), d32)
        # near-far predecessors: bit-exactly tight in Float32
        f = lambda a: a.astype(np.float32)
        i = np.nonzero(Dn.prev > 0)[0]
        j = Dn.prev[i] - 1
        assert len(i) == g.n - 1
        dx, dy, dz = f(X)[i] - f(X)[j], f(Y)[i] - f(Y)[j], f(Z)[i] - f(Z)[j]
        w = np.sqrt(dx * dx + dy * dy + dz * dz) * (np.float32(1) / np.abs(f(U)[i] + f(U)[j])) * np.float32(2)
        assert w.dtype == np.float32 and np.array_equal(f(d32)[j] + w, f(d32)[i])
    rt.bfm3d(g, 1, U, schedule="jacobi")


# ----------------------------------------------------------- single-process multi-GPU batch (rt_bfm_solve_multi)
def test_bfm_multi_replicas(rt, O, annulus, ak135):
    """sources shard over replicas of the mesh, one host thread per replica, tables land in the caller's
    host buffers in source order.  Replicas go to distinct devices when the box has several, otherwise two replicas
    share device 0 (same code path: separate handles, streams and host threads)."""
    m = annulus(36, 10, 100.0)
    Vp = O.interp_velocity(ak135[0], ak135[1], m.r)
    ndev = max(2, min(rt.device_count(), 4))
    reps = []
    for d in range(ndev):
        rt.set_device(d % rt.device_count())
        gr, G, halo = adopt(rt, m)
        rt.mesh_from_arrays(gr, G, halo)
        reps.append(gr)
    rt.set_device(0)
    srcs = np.array([1, m.n, m.n // 2, 7, 1234, m.n // 3, 99], np.int64)  # 7 sources: ragged shards
    for sched in ("jacobi", "near-far"):
        D = rt.bfm_multi(reps, srcs, Vp, schedule=sched)
        assert D.dist.shape == (len(srcs), m.n)
        for k, s in enumerate(srcs):
            d1, p1, _ = O.bfm(m, Vp, int(s))
            assert np.array_equal(D.dist[k], d1), "source #%d (%s)" % (k, sched)
            if sched == "jacobi":
                assert np.array_equal(D.prev[k], p1)
    D32 = rt.bfm_multi(reps, srcs[:3], Vp, schedule="jacobi", precision=32)
    assert np.array_equal(D32.dist[2].astype(np.float64), O.bfm_f32(m, Vp, int(srcs[2]))[0])
    one = rt.bfm_multi(reps, srcs[:1], Vp)  # fewer sources than replicas
    assert np.array_equal(one.dist[0], O.bfm(m, Vp, 1)[0])
    with pytest.raises(rt.RtError):
        rt.bfm_multi(reps, [0], Vp)  # bad source id reported from the worker thread
    with pytest.raises(rt.RtError):
        rt.bfm_multi([reps[0], reps[0]], srcs, Vp)  # the same handle twice
    # 3-D replicas
    nn = (12, 10, 8)
    gs = []
    for d in range(2):
        rt.set_device(d % rt.device_count())
        gs.append(rt.grid((0.0, 0.0, 0.0), (90.0, 70.0, 60.0), nn, neighbour_levels=1, coord_system="cartesian"))
    rt.set_device(0)
    X, Y, Z = gs[0].coordinates()
    U3 = 4.0 + 6.0 * splitmix64(5, gs[0].n)
    D3 = rt.bfm_multi(gs, [1, gs[0].n, 17], U3, schedule="near-far")
    for k, s in enumerate((1, gs[0].n, 17)):
        assert np.array_equal(D3.dist[k], O.bfm3d(nn, 1, X, Y, Z, U3, s)[0])


# ------------------------------------------------------------------ every solver option of the near-far schedule
def test_near_far_option_matrix(rt, O, annulus, ak135):
    """rt_set_option switches between equivalent execution strategies of the near-far schedule (persistent kernel /
    launch sequence / host-driven rounds with timers, warp-per-item / warp-per-(item, element) / CTA units, packed
    128-bit pairs / separate tightness pass, batch width, bucket width).  Travel times must be bit-identical to the
    oracle under every combination, predecessors tie-aware tight."""
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    h = rt.mesh_from_arrays(gr, G, halo)
    Vp = O.interp_velocity(ak135[0], ak135[1], m.r)
    srcs = [O.closest_point(m.theta, m.r, 0.0, R), m.n // 2, 5]
    want = [O.bfm(m, Vp, s) for s in srcs]
    defaults = dict(profile_timers=0, check_every=0, cta_units=0, warp_units=-1, batch=0, packed_prev=1, persistent=-1,
                    delta_factor=0.0, fuse_begin=0, use_graph=1, target_lists=1)
    combos = [dict(persistent=1, warp_units=1), dict(persistent=1, warp_units=0), dict(persistent=0, warp_units=1),
              dict(persistent=0, warp_units=0), dict(persistent=0, warp_units=0, check_every=4),
              dict(persistent=0, warp_units=0, cta_units=1), dict(profile_timers=1, warp_units=0),
              dict(profile_timers=1, warp_units=1), dict(packed_prev=0), dict(packed_prev=0, persistent=0),
              dict(batch=2), dict(batch=1), dict(delta_factor=0.25), dict(delta_factor=16.0),
              # round control as a kernel of its own / in the tail of the preceding kernel, with and without graph replay
              dict(persistent=0, fuse_begin=1), dict(persistent=0, fuse_begin=1, use_graph=0),
              dict(persistent=0, warp_units=0, fuse_begin=1, check_every=5),
              dict(persistent=0, target_lists=0, fuse_begin=1), dict(persistent=0, target_lists=0, use_graph=0)]
    try:
        for combo in combos:
            for k, v in {**defaults, **combo}.items():
                h.set_option(k, v)
            D = rt.bfm(G, halo, srcs, gr, Vp, schedule="near-far")
            for k, s in enumerate(srcs):
                assert np.array_equal(D.dist[k], want[k][0]), "options %s, source %d" % (combo, s)
                check_prev_tie_aware(m, Vp, int(s), want[k][0], D.prev[k], want[k][1])
    finally:
        for k, v in defaults.items():
            h.set_option(k, v)
        h.set_option("schedule", 0)
