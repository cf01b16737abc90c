"""A second, deliberately naive transliteration of the reference solver -- bfm / relax! / _relax!(U::Vector) /
update_halo! / init_halo_path! / init_Q! / update_Q! of src/SSSP/bfm.jl:1-111, 161-210 and distance of
src/GridAnnulus.jl:808-815 -- written statement by statement in plain Python (dicts, lists, Python floats = IEEE
doubles, no vectorisation) with NO code shared with oracle/rt_oracle.cpp.  It cross-pins the oracle's Jacobi tie rule
(strict `>`, scan order, halo rows in serial order) on tiny meshes: two independent restatements of the same Julia
text must produce the same bits.  Test infrastructure only."""
import math

import numpy as np
import pytest

from conftest import splitmix64

R = 6371.0
INF = float("inf")


def naive_bfm(colptr, rowval, halo, source, e2n, x, z, U):
    """Everything 1-based like the Julia text (index 0 of the lists is unused)."""
    n = len(colptr) - 2

    def sp_column(i):  # @views A.rowval[nzrange(A, I)]
        return rowval[colptr[i]:colptr[i + 1]]  # rowval / colptr are 1-based lists with a dummy at [0]

    p = [0] * (n + 1)  # Vector{M}(undef, n): 0 stands for "undefined"
    # init_halo_path!: n = length(halo) ÷ 2 -- halo is a (rows x 2) Matrix, so length ÷ 2 == rows: every row is visited
    for i in range(1, (len(halo) * 2) // 2 + 1):
        p[halo[i - 1][1]] = halo[i - 1][0]
        p[halo[i - 1][0]] = halo[i - 1][1]
    Q = [False] * (n + 1)
    for element in sp_column(source):  # init_Q!
        for i in e2n[element]:
            Q[i] = True
    dist = [INF] * (n + 1)
    dist[source] = 0.0
    dist0 = list(dist)
    it = 1
    while sum(Q[1:]) != 0:
        # relax!
        for i in [q for q in range(1, n + 1) if Q[q]]:  # findall(Q)
            di = dist0[i]
            point = (x[i], z[i])
            Ui = U[i]
            for j in sp_column(i):
                for Gi in e2n[j]:
                    dGi = dist0[Gi]
                    if dGi == INF:
                        delta = INF
                    else:
                        d = 0.0
                        d += (point[0] - x[Gi]) ** 2  # Base.Cartesian.@nexprs: d += (a[i] - b[i])^2
                        d += (point[1] - z[Gi]) ** 2
                        delta = dGi + 2.0 * math.sqrt(d) / (Ui + U[Gi])
                    if di > delta:
                        di = delta
                        p[i] = Gi
            dist[i] = di
        # update_halo!
        for row in halo:
            if dist[row[0]] < dist0[row[0]] and dist[row[1]] > dist[row[0]]:
                dist[row[1]] = dist[row[0]]
                p[row[1]] = p[row[0]]
        Q = [False] * (n + 1)
        # update_Q!
        for i in range(1, n + 1):
            if dist[i] < INF and dist[i] < dist0[i]:
                for element in sp_column(i):
                    for k in e2n[element]:
                        if Q[k] is True:
                            continue
                        Q[k] = True
        dist0 = list(dist)
        it += 1
    return p, dist, it


def as_julia(m):
    colptr = [0] + [int(v) for v in m.G_colptr]      # colptr[i] 1-based offsets into rowval (1-based)
    rowval = [0] + [int(v) for v in m.G_rowval]
    # python slicing rowval[colptr[i]:colptr[i+1]] on the dummy-shifted list == rowval[colptr[i] : colptr[i+1]-1] in Julia
    e2n = {el + 1: [int(v) for v in m.e2n_idx[m.e2n_off[el]:m.e2n_off[el + 1]]] for el in range(m.nel)}
    hm = m.halo_matrix() if m.halo_rows else np.zeros((0, 2), np.int64)
    halo = [(int(a), int(b)) for a, b in hm]
    x = [0.0] + [float(v) for v in m.x]
    z = [0.0] + [float(v) for v in m.z]
    return colptr, rowval, halo, e2n, x, z


@pytest.mark.parametrize("mesh,seed,source", [((8, 2, 600.0), 0, "surface"), ((12, 3, 500.0), 7, "surface"),
                                             ((8, 2, 600.0), 3, "twin"), ((12, 3, 500.0), 11, "last")])
def test_two_independent_restatements_agree(O, ak135, mesh, seed, source):
    m = O.Annulus(*mesh)
    if seed == 0:
        U = O.interp_velocity(ak135[0], ak135[1], m.r)
    else:
        U = 4.0 + 6.0 * splitmix64(seed, m.n)
    hm = m.halo_matrix()
    src = {"surface": O.closest_point(m.theta, m.r, 0.0, R), "twin": int(hm[3, 1]), "last": m.n}[source]
    colptr, rowval, halo, e2n, x, z = as_julia(m)
    p, dist, it = naive_bfm(colptr, rowval, halo, src, e2n, x, z, [0.0] + [float(v) for v in U])
    d_o, p_o, st = O.bfm(m, U, src)
    assert it == st["sweeps"] + 1  # println("Converged in $it iterations"): it starts at 1
    assert np.array_equal(np.array(dist[1:]), d_o)
    assert np.array_equal(np.array(p[1:], np.int64), p_o)  # ties included; 0 = never set in both


def test_naive_halo_rule_is_order_dependent_as_in_the_reference(O):
    """The serial row order of update_halo! matters for chains (a corner duplicated twice): reversing the rows changes
    the predecessors somewhere -- both restatements must follow the ROW ORDER, not a set semantics."""
    m = O.Annulus(8, 2, 600.0)
    U = 4.0 + 6.0 * splitmix64(5, m.n)
    colptr, rowval, halo, e2n, x, z = as_julia(m)
    src = O.closest_point(m.theta, m.r, 0.0, R)
    Ul = [0.0] + [float(v) for v in U]
    p1, d1, _ = naive_bfm(colptr, rowval, halo, src, e2n, x, z, Ul)
    p2, d2, _ = naive_bfm(colptr, rowval, halo[::-1], src, e2n, x, z, Ul)
    assert d1 == d2  # the travel times are the least fixed point either way
    import copy
    m2 = copy.copy(m)
    m2.halo = np.ascontiguousarray(m.halo_matrix()[::-1].T).reshape(-1)
    d_o, p_o, _ = O.bfm(m2, U, src)
    assert np.array_equal(np.array(p2[1:], np.int64), p_o) and np.array_equal(np.array(d2[1:]), d_o)
