"""layer-restricted / multiphase propagation -- partition_grid / GridPartition
(src/topology/topology.jl:137-206) and the restricted continuation that bfm_multiphase
(src/SSSP/bfm_multiphase.jl:30-156) loops over.  The reference routine is an unfinished draft, so the CPU oracle
restates its inner loop on the graph of bfm (oracle.bfm_continue) and the device path must reproduce it bit for bit."""
import numpy as np
import pytest

from conftest import splitmix64

R = 6371.0


def test_partition_grid_oracle(O, annulus):
    m = annulus(24, 6, 300.0)
    ids = O.partition_grid(m.r)
    rl = [R - d for d in (20.0, 35.0, 210.0, 410.0, 660.0, 2740.0, 2891.5)]
    for k, rb in enumerate(rl):
        assert np.array_equal(ids == -(k + 1), np.round(m.r, 2) == rb)
    lay = ids > 0
    edges = np.array([np.inf] + rl + [-np.inf])
    for i in np.flatnonzero(lay)[::37]:
        k = ids[i]
        assert edges[k - 1] > round(m.r[i], 2) > edges[k]
    # twins sit 0.05 km below their discontinuity: they belong to the layer below it (:925-940 of GridAnnulus.jl)
    hm = m.halo_matrix()
    H = m.halo_rows // 2
    assert np.all(ids[hm[:H, 0] - 1] < 0) and np.all(ids[hm[:H, 1] - 1] == -ids[hm[:H, 0] - 1] + 1)


def test_continuation_oracle_semantics(O, annulus, ak135):
    """Unrestricted continuation from the initial state is bfm itself; a restricted one leaves every other node alone."""
    m = annulus(24, 6, 300.0)
    U = O.interp_velocity(ak135[0], ak135[1], m.r)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    dist, prev, st = O.bfm(m, U, src)
    d0 = np.full(m.n, np.inf)
    d0[src - 1] = 0.0
    p0 = np.zeros(m.n, np.int64)
    for a, b in m.halo_matrix():
        p0[b - 1] = a
        p0[a - 1] = b
    d1, p1, s1 = O.bfm_continue(m, U, None, [src], d0, p0)
    assert np.array_equal(d1, dist) and np.array_equal(p1, prev) and s1["sweeps"] == st["sweeps"]
    ids = O.partition_grid(m.r)
    allowed = np.isin(ids, [1, -1]).astype(np.uint8)
    d2, p2, s2 = O.bfm_continue(m, U, allowed, [src], d0, p0)
    out = allowed == 0
    assert np.array_equal(d2[out], d0[out]) and np.array_equal(p2[out], p0[out])
    inside = allowed == 1
    assert np.all(np.isfinite(d2[inside])) and np.all(d2[inside] >= dist[inside])  # fewer paths: never earlier


@pytest.mark.gpu
def test_partition_and_continuation_on_device(rt, O, annulus, ak135):
    m = annulus(36, 10, 100.0)
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    halo = m.halo_matrix()
    rt.mesh_from_arrays(gr, G, halo)
    part = rt.partition_grid(gr)
    ids = O.partition_grid(m.r)
    assert np.array_equal(part.code, ids)
    assert part.nlayers == 8 and part.nboundaries == 7 and len(part.iterator) == 15
    assert part.iterator[1] == part.iterator[15] == ("Layer_1", "Boundary_1")
    assert part.iterator[3] == part.iterator[13] == ("Layer_3", "Boundary_2", "Boundary_3")
    assert part.iterator[8] == ("Layer_8", "Boundary_7")
    assert part.id[0] in ("Layer_%d" % ids[0], "Boundary_%d" % -ids[0])
    src = O.closest_point(m.theta, m.r, 0.0, R)
    d0, p0 = rt.initial_state(halo, src, m.n)
    for U in (O.interp_velocity(ak135[0], ak135[1], m.r), 4.0 + 6.0 * splitmix64(77, m.n)):
        # unrestricted == bfm
        D = rt.bfm_continue(G, halo, gr, U, None, [src], d0, p0)
        dist, prev, st = O.bfm(m, U, src)
        assert np.array_equal(D.dist, dist) and np.array_equal(D.prev, prev) and D.stats["sweeps"] == st["sweeps"]
        # three restricted legs chained like bfm_multiphase: crust, then the next two layers
        d, p = d0, p0
        dd, pp = d0, p0
        for leg, names in enumerate([(1, -1), (2, -1, -2), (3, -2, -3)]):
            allowed = np.isin(ids, names).astype(np.uint8)
            if leg == 0:
                seeds = np.array([src])
            else:
                cand = np.flatnonzero(ids == names[1]) + 1
                seeds = cand[np.isfinite(d[cand - 1])]
            dd, pp, so = O.bfm_continue(m, U, allowed, seeds, dd, pp)
            Dg = rt.bfm_continue(G, halo, gr, U, allowed, seeds, d, p)
            assert np.array_equal(Dg.dist, dd) and np.array_equal(Dg.prev, pp), leg
            assert Dg.stats["sweeps"] == so["sweeps"]
            d, p = Dg.dist, Dg.prev
        assert np.isfinite(d[ids == 3]).all() and np.isinf(d[ids == 5]).all()
    rt.bfm(G, halo, src, gr, U, schedule="jacobi")


@pytest.mark.gpu
def test_bfm_multiphase_driver(rt, O, annulus, ak135):
    m = annulus(36, 10, 100.0)
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    halo = m.halo_matrix()
    rt.mesh_from_arrays(gr, G, halo)
    itp = rt.LinearInterpolation(*ak135)
    U = rt.interpolate_velocity(gr.r, itp)
    src = rt.closest_point(gr, 0.0, R, system="polar")
    part = rt.partition_grid(gr)
    D = rt.bfm_multiphase(G, halo, src, gr, U, part, itp, nphases=3)
    assert len(D.stats["phase_sweeps"]) == 3 and all(s > 0 for s in D.stats["phase_sweeps"])
    ids = part.code
    reached = np.isfinite(D.dist)
    assert reached[np.isin(ids, [1, 2, 3, -1, -2, -3])].all() and not reached[np.isin(ids, [5, 6, 7, 8])].any()
    # the same three legs with the oracle (boundary velocities replaced exactly as the driver does)
    Uo = np.array(U)
    rdir = rt.directions(8)
    d, p = rt.initial_state(halo, src, m.n)
    for i in (1, 2, 3):
        level = part.iterator[i]
        for k, b in enumerate(level[1:]):
            rb = dict(zip(part.boundaries, part.rboundaries))[b]
            Uo[ids == part.code_of(b)] = np.interp(rb - 1.0 if rdir[i][k] == "above" else rb + 1.0, *ak135)
        if i == 1:
            seeds = np.array([src])
        else:
            cand = np.flatnonzero(ids == part.code_of(level[1])) + 1
            seeds = cand[np.isfinite(d[cand - 1])]
        d, p, _ = O.bfm_continue(m, Uo, part.mask(level), seeds, d, p)
    assert np.array_equal(D.dist, d) and np.array_equal(D.prev, p)
    # every path of the multiphase table ends at the source with non-increasing travel time
    rcv = int(np.flatnonzero(ids == 3)[5]) + 1
    path = rt.recontruct_path(D.prev, src, rcv)
    assert path[-1] == src and np.all(np.diff(D.dist[path - 1]) <= 0)
