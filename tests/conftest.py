import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def rt():
    """The product package (ctypes front of librt_sssp.so)."""
    import rt_loader
    rt_loader.load_build().build()
    return rt_loader.load()


@pytest.fixture(scope="session")
def ak135(rt):
    p = rt.velocity_profile()
    return p.r, p.Vp


_MESH_CACHE = {}


@pytest.fixture(scope="session")
def annulus(O):
    def get(ntheta, nr, spacing):
        key = (ntheta, nr, spacing)
        if key not in _MESH_CACHE:
            _MESH_CACHE[key] = O.Annulus(ntheta, nr, spacing)
        return _MESH_CACHE[key]
    return get


def splitmix64(seed, count):
    """Deterministic uniform [0,1) doubles shared by oracle and GPU tests (seed 20261018)."""
    out = np.empty(count, np.float64)
    x = np.uint64(seed)
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        for i in range(count):
            x = (x + np.uint64(0x9E3779B97F4A7C15)) & M
            z = x
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            z = z ^ (z >> np.uint64(31))
            out[i] = float(z >> np.uint64(11)) * (1.0 / 9007199254740992.0)
    return out


def tie_aware_prev_mismatches(prev_new, prev_ref, dist, weight_fn, halo_nodes=()):
    """where prev differs accept iff the new predecessor is bit-exactly tight.  Returns the list
    of node ids (1-based) that are genuinely wrong."""
    bad = []
    halo_nodes = set(int(h) for h in halo_nodes)
    diff = np.nonzero(prev_new != prev_ref)[0]
    for i0 in diff:
        i, j = int(i0) + 1, int(prev_new[i0])
        if j < 1:
            bad.append(i)
            continue
        if i in halo_nodes:
            continue  # halo-set nodes inherit the predecessor of their twin (checked separately)
        if dist[j - 1] + weight_fn(i, j) != dist[i0]:
            bad.append(i)
    return bad
