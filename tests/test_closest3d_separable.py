"""CPU property test of the argument behind closest3d_kernel (raytracer.jl_b200/csrc/grid3d.cu): the rounded distance
fl(sqrt(fl(fl(a_i + b_j) + c_k))) is monotone in each squared axis offset, so the FIRST linear index that attains the
minimum of closest_point(gr, x, y, z) (src/StructuredGrid.jl:257-270) is found axis by axis.  The kernel's search is
restated in numpy (same operation order) and compared with the brute-force restatement of the reference on queries built
to produce ties after rounding (midpoints, far-away points whose offsets are absorbed, points on nodes)."""
import numpy as np
import pytest


def separable(axes, p):
    ax, ay, az = axes
    if not np.all(np.isfinite(np.asarray(p, np.float64))):
        return -1
    a, b, c = (ax - p[0]) ** 2, (ay - p[1]) ** 2, (az - p[2]) ** 2
    val = lambda x, y, z: np.sqrt((x + y) + z)  # noqa: E731
    am, bm, cm = a.min(), b.min(), c.min()
    D = val(am, bm, cm)
    if not D < np.inf:
        return -1
    k0 = int(np.flatnonzero(val(am, bm, c) == D)[0])
    j0 = int(np.flatnonzero(val(am, b, c[k0]) == D)[0])
    i0 = int(np.flatnonzero(val(a, b[j0], c[k0]) == D)[0])
    return i0 + len(ax) * (j0 + len(ay) * k0) + 1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_separable_search_equals_brute_force(O, seed):
    rng = np.random.default_rng(seed)
    for trial in range(6):
        nn = rng.integers(1, 9, 3)
        lo = rng.uniform(-3.0, 3.0, 3)
        hi = lo + rng.uniform(0.0, 4.0, 3) * (rng.random(3) > 0.2)  # some degenerate axes (all nodes equal)
        axes = [np.linspace(lo[d], hi[d], nn[d]) for d in range(3)]
        qs = [rng.uniform(-5.0, 5.0, 3) for _ in range(40)]
        for d in range(3):  # midpoints between neighbouring nodes, nodes themselves
            for i in range(nn[d] - 1):
                q = rng.uniform(-1.0, 1.0, 3)
                q[d] = 0.5 * (axes[d][i] + axes[d][i + 1])
                qs.append(q)
            q = np.array([axes[0][0], axes[1][-1], axes[2][nn[2] // 2]])
            qs.append(q)
        for s in (1e8, 1e16, 1e300):  # offsets of the other axes are absorbed by rounding: long plateaus of equal distance
            qs.append(np.array([s, 0.3, -0.2]))
            qs.append(np.array([0.1, -s, 0.2]))
            qs.append(np.array([s, s, s]))
        qs.append(np.array([np.nan, 0.0, 0.0]))
        qs.append(np.array([0.0, np.inf, 0.0]))
        for q in qs:
            with np.errstate(over="ignore", invalid="ignore"):
                assert separable(axes, q) == O.closest_point3d(axes, q), (nn, q)
