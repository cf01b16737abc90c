"""Host-side logic of the Julia-API mirror (raytracer.jl_b200/api.py) that needs no GPU: containers, reorder!,
travel_times, element_degree, argument checks.  Semantics are checked with the CPU oracle where a solve is needed."""
import os
import types

import numpy as np
import pytest

from conftest import splitmix64

R = 6371.0


def as_grid(rt, m):
    gr = rt.Grid2D(m.x, m.z, m.theta, m.r, m.e2n_off, m.e2n_idx, m.ntheta, m.nr, m.nel, m.n)
    G = rt.SparseMatrixCSC(m.nel, m.n, m.G_colptr, m.G_rowval)
    return gr, G, m.halo_matrix()


def oracle_mesh(gr, G, halo):
    """The attribute set the oracle's bfm wrapper reads, from the mirror's containers."""
    rows = 0 if halo is None else halo.shape[0]
    flat = np.zeros(1, np.int64) if rows == 0 else np.ascontiguousarray(halo.T).reshape(-1)
    return types.SimpleNamespace(n=gr.nnods, nel=gr.nel, e2n_off=gr.e2n_off, e2n_idx=np.ascontiguousarray(gr.e2n_idx),
                                 G_colptr=G.colptr, G_rowval=G.rowval, halo=flat, halo_rows=rows,
                                 x=np.ascontiguousarray(gr.x), z=np.ascontiguousarray(gr.z))


def test_reorder_is_a_relabelling(rt, O, annulus, ak135):
    """reorder!(gr, prm) src/SSSP/rcm.jl:62-94: after relabelling with ANY permutation the travel times are the same
    function of the nodes (checked with the oracle's bfm on the relabelled arrays) and the containers stay valid."""
    m = annulus(24, 6, 300.0)
    gr, G, halo = as_grid(rt, m)
    prm = np.argsort(splitmix64(42, m.n), kind="stable").astype(np.int64) + 1  # a random permutation, 1-based
    gr2, G2, halo2 = rt.reorder(gr, G, halo, prm)
    assert np.array_equal(gr2.x, m.x[prm - 1]) and np.array_equal(gr2.r, m.r[prm - 1])
    inv = np.zeros(m.n + 1, np.int64)
    inv[prm] = np.arange(1, m.n + 1)
    for el in (1, 7, m.nel):
        assert np.array_equal(gr2.e2n[el], inv[gr.e2n[el]])  # rordering_map :87-94
    for node in (1, m.n // 2, m.n):  # column of the new node = column of the node it came from
        a = G2.rowval[G2.colptr[node - 1] - 1:G2.colptr[node] - 1]
        old = prm[node - 1]
        b = G.rowval[G.colptr[old - 1] - 1:G.colptr[old] - 1]
        assert np.array_equal(a, b)
    assert G2.colptr[0] == 1 and G2.colptr[-1] == G.colptr[-1]
    U = O.interp_velocity(ak135[0], ak135[1], m.r)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    d, _, _ = O.bfm(m, U, src)
    d2, p2, _ = O.bfm(oracle_mesh(gr2, G2, halo2), U[prm - 1], int(inv[src]))
    assert np.array_equal(d2, d[prm - 1])
    # identity permutation is a no-op
    gr3, G3, halo3 = rt.reorder(gr, G, halo, np.arange(1, m.n + 1))
    assert np.array_equal(gr3.e2n_idx, gr.e2n_idx) and np.array_equal(G3.rowval, G.rowval)
    assert np.array_equal(halo3, halo)


@pytest.mark.gpu
def test_travel_times_and_csv(rt, tmp_path, monkeypatch):
    """travel_times(D, gr, receivers; isave, flname) src/utils.jl:4-15 (the gather runs on the device since round 2:
    rt_travel_times, so this needs a GPU like every other compute call)."""
    n = 50
    theta = np.linspace(0.0, np.pi, n)
    gr = rt.Grid2D(np.zeros(n), np.zeros(n), theta, np.full(n, R), np.zeros(1, np.int64), np.zeros(0, np.int64), 1, 1,
                   0, n)
    D = rt.BellmanFordMoore(np.zeros(n, np.int64), np.arange(n, dtype=np.float64) * 1.5)
    rec = np.array([1, 10, 50])
    assert np.array_equal(rt.travel_times(D, gr, rec), [0.0, 13.5, 73.5])
    monkeypatch.chdir(tmp_path)
    rt.travel_times(D, gr, rec, isave=True, flname="tt.csv")
    rows = open(os.path.join(tmp_path, "tt.csv")).read().strip().split("\n")
    assert rows[0] == "degree,travel_time" and len(rows) == 4
    deg, tt = (float(v) for v in rows[3].split(","))
    assert deg == np.rad2deg(theta[49]) and tt == 73.5


def test_containers_and_argument_checks(rt, annulus):
    m = annulus(24, 6, 300.0)
    gr, G, halo = as_grid(rt, m)
    assert len(gr) == m.n and np.array_equal(gr.e2n[1], m.e2n_idx[m.e2n_off[0]:m.e2n_off[1]])
    assert np.array_equal(rt.element_degree(G), np.diff(m.G_colptr))
    assert rt.R == 6371.0
    prof = rt.velocity_profile()
    assert prof.r[0] == 0.0 and prof.r[-1] == R and len(prof.r) == len(prof.Vp) == 6372  # utils.jl:23-30
    assert np.all(np.diff(prof.r) > 0)
    itp = rt.LinearInterpolation(prof.r, prof.Vp)
    assert len(itp.knots) == len(itp.values)
    with pytest.raises(ValueError):  # no device handle yet: the mirror refuses before touching the library
        rt.closest_point(gr, 0.0, R)
    with pytest.raises(ValueError):
        rt.bfm_multi([gr], [1], np.ones(m.n))
    with pytest.raises(KeyError):
        rt.api.SCHEDULES["dijkstra"]
    assert rt.api.SCHEDULES == {"jacobi": 0, "near-far": 1}
