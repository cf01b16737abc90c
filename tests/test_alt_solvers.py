"""the reference's alternative solvers behind the Dijkstra / RadiusStepping result structs --
dijkstra(G, source, gr, U) src/SSSP/dijkstra.jl:68-136 and radius_stepping(Gsp, source, gr, U)
src/SSSP/radius_stepping.jl:7-46 -- on the star-0 node graph nodal_incidence(gr).  The oracle holds literal
transliterations of both loops (set-scan priority queue, settle-by-radius); the CUDA path reaches the same least fixed
point with a label-correcting frontier relaxation and derives the predecessors from the settle order."""
import numpy as np
import pytest

from conftest import splitmix64

R = 6371.0


def test_oracle_dijkstra_and_radius_stepping_agree(O, annulus, ak135):
    """CPU: the two literal loops give the same tables; on this graph (no halo coupling) a surface source only reaches
    the crustal layer above the first discontinuity; tightness of every predecessor with independent numpy arithmetic."""
    m = annulus(24, 6, 300.0)
    adj = O.nodal_adjacency(m)
    for U in (O.interp_velocity(ak135[0], ak135[1], m.r), 4.0 + 6.0 * splitmix64(3, m.n)):
        src = O.closest_point(m.theta, m.r, 0.0, R)
        d, p = O.dijkstra_nodal(m, U, src, adj)
        d2, p2, it = O.radius_stepping_nodal(m, U, src, adj)
        assert np.array_equal(d, d2) and np.array_equal(p, p2)
        reached = np.isfinite(d)
        assert 1 < reached.sum() < m.n and np.all(m.r[reached] >= R - 20.0 - 1e-9)
        i = np.flatnonzero(reached & (p > 0))
        j = p[i] - 1
        w = 2 * np.sqrt((m.x[j] - m.x[i]) ** 2 + (m.z[j] - m.z[i]) ** 2) / np.abs(U[i] + U[j])
        assert np.array_equal(d[j] + w, d[i]) and len(i) == reached.sum() - 1


@pytest.mark.gpu
@pytest.mark.parametrize("mesh", [(24, 6, 300.0), (36, 10, 100.0)])
def test_dijkstra_and_radius_stepping_on_device(rt, O, annulus, ak135, mesh):
    m = annulus(*mesh)
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    rt.mesh_from_arrays(gr, G, m.halo_matrix())
    adj = O.nodal_adjacency(m)
    cases = [(O.interp_velocity(ak135[0], ak135[1], m.r), O.closest_point(m.theta, m.r, 0.0, R)),
             (4.0 + 6.0 * splitmix64(3, m.n), 1), (4.0 + 6.0 * splitmix64(8, m.n), m.n),
             (np.full(m.n, 6.0), m.n // 3)]
    for U, src in cases:
        d, p = O.dijkstra_nodal(m, U, src, adj)
        D = rt.dijkstra(None, src, gr, U)
        assert isinstance(D, rt.Dijkstra) and np.array_equal(D.dist, d) and np.array_equal(D.prev, p)
        d2, p2, it = O.radius_stepping_nodal(m, U, src, adj)
        Dr = rt.radius_stepping(None, src, gr, U)
        assert isinstance(Dr, rt.RadiusStepping) and np.array_equal(Dr.dist, d2) and np.array_equal(Dr.prev, p2)
        reached = np.flatnonzero(np.isfinite(d))
        rcv = int(reached[len(reached) // 2]) + 1
        if rcv != src:
            assert np.array_equal(rt.recontruct_path(D.prev, src, rcv), O.reconstruct_path(p, src, rcv))
    with pytest.raises(rt.RtError):
        rt.dijkstra(None, m.n + 1, gr, cases[0][0])
