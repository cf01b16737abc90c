"""Near-far schedule + canonical-predecessor pass (raytracer.jl_b200/csrc/canonical_prev.cu): the
predecessor table of the work-efficient schedule must equal the reference's (oracle Jacobi sweeps) bit for bit, exact
ties included -- not merely be tie-aware valid.  Also the larger oracle comparisons VERDICT r1 asked for (180x50 at
20 km and 5 km) and the full-size pin of BASELINE configs[0] (README example at 1 km)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import splitmix64

pytestmark = pytest.mark.gpu
R = 6371.0
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def adopt(rt, m):
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    halo = m.halo_matrix() if m.halo_rows else None
    return gr, G, halo


@pytest.mark.parametrize("nt,nr,sp", [(24, 6, 300.0), (36, 10, 100.0), (180, 50, 50.0), (90, 20, 20.0)])
def test_near_far_canonical_prev_equals_reference(rt, O, annulus, ak135, nt, nr, sp):
    m = annulus(nt, nr, sp)
    gr, G, halo = adopt(rt, m)
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(*ak135))
    src = O.closest_point(m.theta, m.r, 0.0, R)
    dist, prev, st = O.bfm(m, Vp, src)
    D = rt.bfm(G, halo, src, gr, Vp, schedule="near-far", canonical_prev=True)
    assert np.array_equal(D.dist, dist)
    assert np.array_equal(D.prev, prev)  # ties included
    D2 = rt.bfm(G, halo, src, gr, Vp, schedule="near-far", canonical_prev=True)
    assert np.array_equal(D2.prev, D.prev)  # run-to-run deterministic
    # without the pass the table is only tie-aware valid (it differs somewhere on these meshes)
    D0 = rt.bfm(G, halo, src, gr, Vp, schedule="near-far", canonical_prev=False)
    assert np.array_equal(D0.dist, dist)
    # README receivers -> identical paths
    degs = np.concatenate([np.arange(10, 151, 10), 360 - np.arange(150, 9, -10)]).astype(np.float32)
    recv = rt.closest_point(gr, np.deg2rad(degs).astype(np.float64), np.full(len(degs), R), system="polar")
    for rc, pth in zip(recv, rt.recontruct_path(D.prev, src, recv)):
        assert np.array_equal(pth, O.reconstruct_path(prev, src, int(rc)))
    rt.bfm(G, halo, src, gr, Vp, schedule="jacobi", canonical_prev=False)


def test_canonical_prev_random_velocity_special_sources_and_batch(rt, O, annulus):
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    hm = m.halo_matrix()
    H = m.halo_rows // 2
    srcs = [1, m.n, int(hm[0, 0]), int(hm[0, 1]), int(hm[H - 1, 1]), m.n // 2, int(hm[H // 2, 0])]
    # random velocities: no ties at all; a CONSTANT velocity: dozens of exact ties per node (collinear nodes, duplicated
    # radial edges: up to 29 distinct bit-exactly tight predecessors on this mesh), final values first reached through
    # non-final values of a neighbour -- the case in which a breadth-first levelling of the tight edges is wrong
    for U in (4.0 + 6.0 * splitmix64(20261018, m.n), np.full(m.n, 6.0)):
        D = rt.bfm(G, halo, np.array(srcs), gr, U, schedule="near-far", canonical_prev=True)  # one lock-step batch
        for k, s in enumerate(srcs):
            dist, prev, _ = O.bfm(m, U, s)
            assert np.array_equal(D.dist[k], dist), s
            assert np.array_equal(D.prev[k], prev), s
            Ds = rt.bfm(G, halo, s, gr, U, schedule="near-far", canonical_prev=True)
            assert np.array_equal(Ds.prev, prev), s
    m2 = annulus(24, 6, 300.0)
    gr2, G2, halo2 = adopt(rt, m2)
    U2 = np.full(m2.n, 8.0)
    for s in (1, 17, m2.n // 2):
        dist, prev, _ = O.bfm(m2, U2, s)
        D2 = rt.bfm(G2, halo2, s, gr2, U2, schedule="near-far", canonical_prev=True)
        assert np.array_equal(D2.dist, dist) and np.array_equal(D2.prev, prev), s
    rt.bfm(G, halo, 1, gr, U, schedule="jacobi", canonical_prev=False)


def test_canonical_prev_without_halo_float32_and_dual(rt, O, annulus, ak135):
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(*ak135))
    src = O.closest_point(m.theta, m.r, 0.0, R)
    # Float32 arithmetic of bfm_gpu
    d32, p32, _ = O.bfm_f32(m, Vp, src)
    D32 = rt.bfm_gpu(G, halo, src, gr, Vp, schedule="near-far", canonical_prev=True)
    assert np.array_equal(D32.dist.astype(np.float64), d32) and np.array_equal(D32.prev, p32)
    # dual-velocity relax (U::Matrix)
    V2 = O.dual_velocity(ak135[0], ak135[1], m.r, 1.0)
    dd, pd, _ = O.bfm_dual(m, V2, src)
    Dd = rt.bfm(G, halo, src, gr, V2, schedule="near-far", canonical_prev=True)
    assert np.array_equal(Dd.dist, dd) and np.array_equal(Dd.prev, pd)
    # the same graph without its halo matrix (discontinuities disconnect the layers: unreached nodes keep prev 0)
    gr2 = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    import copy
    m2 = copy.copy(m)
    m2.halo_rows, m2.halo = 0, np.zeros(0, np.int64)
    dn, pn, _ = O.bfm(m2, Vp, src)
    Dn = rt.bfm(G, None, src, gr2, Vp, schedule="near-far", canonical_prev=True)
    assert np.array_equal(Dn.dist, dn) and np.array_equal(Dn.prev, pn)
    rt.bfm(G, halo, src, gr, Vp, schedule="jacobi", canonical_prev=False)


@pytest.mark.parametrize("sp", [20.0, 5.0])
def test_annulus_180_50_against_oracle(rt, O, ak135, sp):
    """185 401 nodes (the default spacing) and 758 701 nodes against the oracle, both schedules."""
    m = O.Annulus(180, 50, sp)
    gr, G, halo = adopt(rt, m)
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(*ak135))
    src = O.closest_point(m.theta, m.r, 0.0, R)
    dist, prev, st = O.bfm(m, Vp, src, nthreads=O.num_threads())
    Dn = rt.bfm(G, halo, src, gr, Vp, schedule="near-far", canonical_prev=True)
    assert np.array_equal(Dn.dist, dist) and np.array_equal(Dn.prev, prev)
    if sp == 20.0:
        Dj = rt.bfm(G, halo, src, gr, Vp, schedule="jacobi", canonical_prev=False)
        assert np.array_equal(Dj.dist, dist) and np.array_equal(Dj.prev, prev) and Dj.stats["sweeps"] == st["sweeps"]
        gr2, G2, halo2 = rt.init_annulus(180, 50, spacing=sp)  # device builder: same integer topology
        assert np.array_equal(G2.rowval, m.G_rowval) and np.array_equal(gr2.e2n_idx, m.e2n_idx)
        assert np.array_equal(halo2, halo)


def test_config0_readme_example_full_size_pinned(rt, O):
    """BASELINE configs[0] at its full 1 km spacing (3 772 801 nodes) against the committed oracle fixture
    (tests/golden/make_config0_fixture.py: one oracle solve at this size is ~1.3e12 candidate evaluations).  The mesh
    arrays come from the oracle builder (seconds), so the inputs are the fixture's bytes; the benchmarked schedule
    (near-far) with the canonical-predecessor pass must then reproduce sha256(dist), sha256(prev), the travel times at
    the 30 README receivers and the 30 paths."""
    f = os.path.join(GOLD, "config0_180_50_1km.npz")
    if not os.path.exists(f):
        pytest.skip("fixture not generated")
    g = np.load(f)
    m = O.Annulus(180, 50, 1.0)
    assert m.n == int(g["n"]) and m.nel == int(g["nel"]) and m.halo_rows == int(g["halo_rows"])
    assert hashlib.sha256(m.e2n_idx.tobytes() + m.G_rowval.tobytes() + m.halo.tobytes()).hexdigest() == str(g["sha256_topology"])
    gr, G, halo = adopt(rt, m)
    prof = rt.velocity_profile()
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(prof.r, prof.Vp))
    assert hashlib.sha256(Vp.tobytes()).hexdigest() == str(g["sha256_U"])
    rt.mesh_from_arrays(gr, G, halo)
    src = rt.closest_point(gr, 0.0, R, system="polar")
    assert src == int(g["source"])
    D = rt.bfm(G, halo, src, gr, Vp, schedule="near-far", canonical_prev=True)
    recv = g["receivers"]
    assert np.array_equal(D.dist[::997], g["dist_sample"])
    assert np.array_equal(D.dist[recv - 1], g["T_receivers"])
    assert hashlib.sha256(D.dist.tobytes()).hexdigest() == str(g["sha256_dist"])
    assert hashlib.sha256(np.ascontiguousarray(D.prev, np.int64).tobytes()).hexdigest() == str(g["sha256_prev"])
    degs = np.concatenate([np.arange(10, 151, 10), 360 - np.arange(150, 9, -10)]).astype(np.float32)
    got_recv = rt.closest_point(gr, np.deg2rad(degs).astype(np.float64), np.full(len(degs), R), system="polar")
    assert np.array_equal(got_recv, recv)
    for k, pth in enumerate(rt.recontruct_path(D.prev, src, recv)):
        assert np.array_equal(pth, g["path_idx"][g["path_off"][k]:g["path_off"][k + 1]])
    # the device builder reproduces the same mesh (integer topology identical, coordinates to the last ulps)
    gr2, G2, halo2 = rt.init_annulus(180, 50, spacing=1.0)
    assert np.array_equal(G2.rowval, m.G_rowval) and np.array_equal(gr2.e2n_idx, m.e2n_idx) and np.array_equal(halo2, halo)
    assert np.array_equal(gr2.r, m.r) and np.array_equal(gr2.theta, m.theta)
    D2 = rt.bfm(G2, halo2, src, gr2, Vp, schedule="near-far", canonical_prev=False)
    assert np.allclose(D2.dist, D.dist, rtol=1e-12, atol=0)


def test_library_sharded_solve_single_rank(rt, O, annulus, ak135):
    """rt_comm_* / rt_bfm_solve_sharded (NCCL bound at run time inside the library) with a world of one rank: the gathered
    tables equal the tables of rt_bfm_solve.  (The N > 1 exchange itself runs in bench.py's batch_cfg3 block.)"""
    import ctypes as C
    import torch
    m = annulus(36, 10, 100.0)
    gr, G, halo = adopt(rt, m)
    Vp = rt.interpolate_velocity(gr.r, rt.LinearInterpolation(*ak135))
    h = rt.mesh_from_arrays(gr, G, halo)
    srcs = np.array([1, 57, 400, m.n, 9000], np.int64)
    raw = C.create_string_buffer(128)
    rt.api.check(rt.lib().rt_comm_unique_id(raw))
    comm = C.c_void_p()
    rt.api.check(rt.lib().rt_comm_init(raw.raw, 0, 1, C.byref(comm)))
    first, count = C.c_int64(), C.c_int64()
    for nsrc, world in ((5, 1), (10, 4), (3, 8)):
        tot = 0
        for r in range(world):
            rt.api.check(rt.lib().rt_comm_shard(nsrc, r, world, C.byref(first), C.byref(count)))
            assert first.value == tot and 0 <= count.value <= -(-nsrc // world)
            tot += count.value
        assert tot == nsrc
    U = torch.from_numpy(Vp).cuda()
    d = torch.empty((len(srcs), m.n), dtype=torch.float64, device="cuda")
    p = torch.empty((len(srcs), m.n), dtype=torch.int32, device="cuda")
    h.set_option("schedule", 1)
    st = rt.RtStats()
    rt.api.check(rt.lib().rt_bfm_solve_sharded(comm, h.h, U.data_ptr(), srcs, len(srcs), 64, d.data_ptr(), p.data_ptr(),
                                               C.byref(st)))
    # the host-buffer front (what a plain Julia process calls): 1-based int64 predecessors like rt_bfm_solve
    dh = np.empty((len(srcs), m.n), np.float64)
    ph = np.empty((len(srcs), m.n), np.int64)
    rt.api.check(rt.lib().rt_bfm_solve_sharded_host(comm, h.h, np.ascontiguousarray(Vp), srcs, len(srcs), 64,
                                                    dh.ctypes.data, ph.ctypes.data, C.byref(st)))
    rt.api.check(rt.lib().rt_comm_destroy(comm))
    D = rt.bfm(G, halo, srcs, gr, Vp, schedule="near-far")
    assert np.array_equal(d.cpu().numpy(), D.dist)
    assert np.array_equal(dh, D.dist) and np.array_equal(ph, p.cpu().numpy().astype(np.int64) + 1)
    for k, s in enumerate(srcs):
        assert np.array_equal(D.dist[k], O.bfm(m, Vp, int(s))[0])
    h.set_option("schedule", 0)


@pytest.mark.parametrize("nn,lv,cs,wm", [((7, 6, 5), 1, "spherical", 0), ((11, 11, 11), 1, "cartesian", 0),
                                         ((20, 9, 13), 0, "spherical", 1), ((33, 18, 10), 2, "cartesian", 0),
                                         ((40, 40, 24), 1, "spherical", 0)])
def test_near_far_3d_canonical_prev_equals_reference_schedule(rt, O, nn, lv, cs, wm):
    """3-D: near-far + canonical pass == the Jacobi schedule's predecessors bit for bit (canonical scan order =
    ascending linear id), including the homogeneous Cartesian grid where collinear exact ties are everywhere."""
    if cs == "spherical":
        c0 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0)
        c1 = (np.deg2rad(110.0), np.deg2rad(110.0), R)
    else:
        c0, c1 = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    g = rt.grid(c0, c1, nn, neighbour_levels=lv, coord_system=cs)
    X, Y, Z = g.coordinates()
    n = g.n
    g._handle.set_option("weight3d", wm)
    O.set_weight3d(wm)
    try:
        for U in (4.0 + 6.0 * splitmix64(7 + n, n), np.ones(n)):
            for src in (1, n // 2 + 3):
                dist, prev, st = O.bfm3d(nn, lv, X, Y, Z, U, src)
                D = rt.bfm3d(g, src, U, schedule="near-far", canonical_prev=True)
                assert np.array_equal(D.dist, dist)
                assert np.array_equal(D.prev, prev)
        d32, p32, _ = O.bfm3d_f32(nn, lv, X, Y, Z, U, 1)
        D32 = rt.bfm3d(g, 1, U, schedule="near-far", precision=32, canonical_prev=True)
        assert np.array_equal(D32.dist.astype(np.float64), d32) and np.array_equal(D32.prev, p32)
        # a batch keeps its sources in flight in per-source slots: same tables, canonical predecessors included
        srcs = np.array([1, n // 2 + 3, n, 2, n // 3 + 1], np.int64)
        DB = rt.bfm3d(g, srcs, U, schedule="near-far", canonical_prev=True)
        for k, sk in enumerate(srcs):
            dist, prev, st = O.bfm3d(nn, lv, X, Y, Z, U, int(sk))
            assert np.array_equal(DB.dist[k], dist) and np.array_equal(DB.prev[k], prev), "batch slot %d" % k
    finally:
        O.set_weight3d(0)
        g._handle.set_option("weight3d", 0)
        g._handle.set_option("canonical_prev", 0)
