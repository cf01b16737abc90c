"""The two header-only pieces every relax / push / tightness kernel inlines, run on the CPU:
* exact.h -- the exact candidate expression: equals numpy Float64 arithmetic in the reference's operation order, and in
  Float32 mode (fp64 operations each rounded to Float32) equals genuine `float` arithmetic bit for bit;
* screen.h -- the algebraic screen: whenever it answers "skip", the exact candidate is >= the incumbent, i.e. skipping
  cannot change a result; whenever a candidate is bit-exactly tight, `maybe_tight` is true.  Inputs are adversarial:
  incumbents within a few ulp of the candidate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
U8P = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def drv():
    so = os.path.join(HERE, "libscreen_test.so")
    src = os.path.join(HERE, "screen_driver.cpp")
    hdrs = [os.path.join(HERE, "..", "raytracer.jl_b200", "csrc", h) for h in ("screen.h", "exact.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so,
                               src])
    L = C.CDLL(so)
    L.screen2d_batch.argtypes = [C.c_long, C.c_int] + [F64P] * 9 + [U8P, U8P, F64P, F64P]
    L.screen3d_batch.argtypes = [C.c_long, C.c_int] + [F64P] * 11 + [U8P, U8P, F64P, F64P]
    L.group2d_batch.argtypes = [C.c_long, C.c_int, C.c_int] + [F64P] * 8 + [U8P, F64P]
    return L


def nextafter_k(a, k, dtype):
    """a moved by k units in the last place of `dtype` (k may be negative)."""
    b = a.astype(dtype)
    step = np.where(k > 0, np.inf, -np.inf).astype(dtype)
    for _ in range(int(np.max(np.abs(k)))):
        mv = np.abs(k) > 0
        b = np.where(mv, np.nextafter(b, step), b)
        k = k - np.sign(k) * mv
    return b.astype(np.float64)


def cases(rng, n, dim, f32):
    r = lambda a: a.astype(np.float32).astype(np.float64) if f32 else a
    base = rng.uniform(-6371.0, 6371.0, (dim, n))
    # neighbour offsets from 1e-3 km (secondary nodes at fine spacing) to 300 km, plus exact coincidences
    scale = 10.0 ** rng.uniform(-3.0, 2.5, n)
    off = rng.normal(size=(dim, n)) * scale
    off[:, rng.random(n) < 0.02] = 0.0
    pi = [r(base[k]) for k in range(dim)]
    pj = [r(base[k] + off[k]) for k in range(dim)]
    Ui, Uj = r(rng.uniform(1.0, 14.0, n)), r(rng.uniform(1.0, 14.0, n))
    dj = r(np.where(rng.random(n) < 0.05, 0.0, rng.uniform(0.0, 2500.0, n)))
    return pi, pj, Ui, Uj, dj


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("f32", [0, 1])
def test_screen_is_exact_safe(drv, dim, f32):
    rng = np.random.default_rng(20261018 + 10 * dim + f32)
    n = 400000
    pi, pj, Ui, Uj, dj = cases(rng, n, dim, bool(f32))
    dtype = np.float32 if f32 else np.float64
    skip, tight, delta, dker = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n), np.zeros(n)

    def run(bound, target):
        args = [np.ascontiguousarray(bound), dj] + pi + [Ui] + pj + [Uj, np.ascontiguousarray(target)]
        (drv.screen2d_batch if dim == 2 else drv.screen3d_batch)(n, f32, *args, skip, tight, delta, dker)

    run(np.full(n, np.inf), np.zeros(n))  # first pass: the exact candidate values
    d0 = delta.copy()
    assert not skip.any()  # an unreached target (incumbent Inf) is never skipped
    # independent numpy evaluation of the exact candidate (same operation order, numpy never contracts)
    P, Q = [a.astype(dtype) for a in pi], [a.astype(dtype) for a in pj]
    s = sum((a - b) * (a - b) for a, b in zip(P, Q)) if dim == 3 else (P[0] - Q[0]) * (P[0] - Q[0]) + (P[1] - Q[1]) * (P[1] - Q[1])
    if dim == 2:
        want = dj.astype(dtype) + dtype(2) * np.sqrt(s) / (Ui.astype(dtype) + Uj.astype(dtype))
    else:
        want = dj.astype(dtype) + np.sqrt(s) * (dtype(1) / np.abs(Ui.astype(dtype) + Uj.astype(dtype))) * dtype(2)
    assert np.array_equal(want.astype(np.float64), d0)
    # the expression the kernels inline (exact.h): Float64 = the reference order; Float32 = exact emulation
    assert np.array_equal(dker, d0)
    assert (d0 == dj).sum() > 1000 and np.isfinite(d0).all()  # coincident nodes (zero weight) are in the sample
    checked = skipped = 0
    for trial in range(6):
        if trial < 4:   # incumbents within +-40 ulp of the candidate: the adversarial band
            k = rng.integers(-40, 41, n)
            bound = nextafter_k(d0, k, dtype)
        elif trial == 4:  # incumbents a little further away
            bound = (d0 * (1.0 + rng.normal(size=n) * (3e-6 if f32 else 3e-9))).astype(dtype).astype(np.float64)
        else:  # anything
            bound = (dj + (d0 - dj) * rng.uniform(0.0, 3.0, n)).astype(dtype).astype(np.float64)
        ok = dj < bound  # the callers' precondition
        run(bound, d0)
        bad = ok & (skip != 0) & ~(d0 >= bound)
        assert not bad.any(), "unsafe skip at %s" % np.nonzero(bad)[0][:5]
        assert np.all(tight[np.isfinite(d0)] != 0)  # target == candidate: must never be ruled out
        checked += int(ok.sum())
        skipped += int((ok & (skip != 0)).sum())
        if trial == 5:  # sharpness: clearly losing candidates are skipped, clearly non-tight ones ruled out
            margin = 1e-4 if f32 else 1e-7
            lose = ok & (d0 > bound * (1.0 + margin)) & (d0 - dj > 1e-9)
            assert (skip[lose] != 0).mean() > 0.999
            run(bound, np.where(d0 > dj, dj + (d0 - dj) * 0.5, d0))  # a target half a weight below the candidate
            far = ok & (d0 - dj > (1e-3 if f32 else 1e-6) * np.maximum(d0, 1.0))
            assert (tight[far] == 0).mean() > 0.999
    assert checked > n and skipped > 0


@pytest.mark.parametrize("f32", [0, 1])
def test_group_screen_is_exact_safe(drv, f32):
    """group_cannot_improve_t (one test per released item and target in the push kernels): whenever it says "no source
    of this group can improve the target", the smallest EXACT candidate over the group's sources is >= the incumbent.
    Groups: up to 32 collinear-ish sources within a few km (nodes of one mesh edge); incumbents placed within a few ulp
    of the best candidate, and exactly on the disc bound."""
    rng = np.random.default_rng(20261018 + f32)
    r = (lambda a: a.astype(np.float32).astype(np.float64)) if f32 else (lambda a: a)
    ng, ns = 40000, 24
    base = rng.uniform(-6371.0, 6371.0, (2, ng))
    span = 10.0 ** rng.uniform(-2.0, 1.3, ng)                      # group extent 0.01 .. 20 km
    dirv = rng.normal(size=(2, ng))
    dirv /= np.linalg.norm(dirv, axis=0)
    tpar = rng.uniform(-0.5, 0.5, (ns, ng))
    xs = r((base[0] + dirv[0] * span * tpar + rng.normal(size=(ns, ng)) * span * 0.02).T.copy())
    zs = r((base[1] + dirv[1] * span * tpar + rng.normal(size=(ns, ng)) * span * 0.02).T.copy())
    Us = r(rng.uniform(1.0, 14.0, (ng, ns)))
    d0 = rng.uniform(0.0, 2500.0, ng)
    ds = r(d0[:, None] + rng.uniform(0.0, 3.0, (ng, ns)) * rng.random((ng, 1)))
    dist_t = 10.0 ** rng.uniform(-2.0, 2.5, ng)
    ang = rng.uniform(0, 2 * np.pi, ng)
    xt, zt = r(base[0] + dist_t * np.cos(ang)), r(base[1] + dist_t * np.sin(ang))
    Ut = r(rng.uniform(1.0, 14.0, ng))
    skip = np.zeros(ng, np.uint8)
    best = np.zeros(ng)
    # pass 1: the best exact candidate of every group
    drv.group2d_batch(ng, ns, f32, np.full(ng, 1e300), ds.reshape(-1), xs.reshape(-1), zs.reshape(-1), Us.reshape(-1), xt,
                      zt, Ut, skip, best)
    # pass 2: incumbents around that value (a few ulp either side, and generously above / below)
    eps = 2.0 ** -23 if f32 else 2.0 ** -52
    k = rng.integers(-4, 5, ng)
    dmin0 = ds.min(axis=1)
    sel = rng.random(ng)
    bound = np.where(sel < 0.4, best * (1.0 + k * eps),
                     np.where(sel < 0.7, best * rng.uniform(0.7, 1.3, ng), dmin0 + rng.random(ng) * (best - dmin0)))
    bound = r(np.maximum(bound, 0.0))
    drv.group2d_batch(ng, ns, f32, bound, ds.reshape(-1), xs.reshape(-1), zs.reshape(-1), Us.reshape(-1), xt, zt, Ut, skip,
                      best)
    wrong = (skip == 1) & (best < bound)
    assert not wrong.any(), "the group screen hid an improvement"
    # it is not vacuous: among the targets it is meant for (incumbent above the group's smallest travel time but below
    # the best candidate by a margin) a good part is pruned without looking at a single source
    dmin = ds.min(axis=1)
    meant = (bound > dmin) & (bound < best * (1.0 - 1e-3))
    assert meant.sum() > 1000 and skip[meant].mean() > 0.3
