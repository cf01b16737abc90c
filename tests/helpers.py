"""Numpy helpers shared by the parity tests (independent of the oracle's C++ arithmetic)."""
import numpy as np


def scan_pairs(mesh):
    """All (target i, candidate j, scan position) triples of the reference scan lists, 0-based numpy arrays:
    for el in G[:, i] (ascending), for j in e2n[el] (list order) -- src/SSSP/bfm.jl:172."""
    colptr, rowval = mesh.G_colptr, mesh.G_rowval
    e2n_off, e2n_idx = mesh.e2n_off, mesh.e2n_idx
    n = mesh.n
    col_len = np.diff(colptr)
    tgt_of_entry = np.repeat(np.arange(n), col_len)          # per nnz entry: its column (node)
    el = rowval - 1                                          # per nnz entry: element (0-based)
    el_len = (e2n_off[1:] - e2n_off[:-1])[el]
    tgt = np.repeat(tgt_of_entry, el_len)
    start = np.repeat(e2n_off[el], el_len)
    # position inside each element list
    tot = int(el_len.sum())
    first = np.cumsum(el_len) - el_len
    inner = np.arange(tot) - np.repeat(first, el_len)
    cand = e2n_idx[start + inner] - 1
    return tgt, cand


def weight2d(x, z, U, i, j):
    """2.0 * sqrt(dx^2 + dz^2) / (Ui + Uj) with the reference's operation order (numpy never contracts)."""
    dx = x[i] - x[j]
    dz = z[i] - z[j]
    return 2.0 * np.sqrt(dx * dx + dz * dz) / (U[i] + U[j])


def weight3d(X, Y, Z, U, i, j):
    dx, dy, dz = X[i] - X[j], Y[i] - Y[j], Z[i] - Z[j]
    d = np.sqrt(dx * dx + dy * dy + dz * dz)
    return d * (1.0 / np.abs(U[i] + U[j])) * 2.0


def ulp_diff(a, b):
    """Max distance in units in the last place between two float64 arrays (same sign assumed)."""
    ai = np.ascontiguousarray(a, np.float64).view(np.int64)
    bi = np.ascontiguousarray(b, np.float64).view(np.int64)
    return int(np.max(np.abs(ai - bi))) if len(ai) else 0
