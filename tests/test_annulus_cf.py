"""The closed-form annulus description (raytracer.jl_b200/csrc/annulus_cf.cuh, the code the CUDA builder
kernels execute) evaluated sequentially on the CPU must reproduce the oracle's literal restatement of
init_annulus exactly: integer topology identical, (theta, r) bit-identical."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def cfh():
    so = os.path.join(HERE, "libcf_host.so")
    src = os.path.join(HERE, "cf_host_driver.cpp")
    hdr = os.path.join(HERE, "..", "raytracer.jl_b200", "csrc", "annulus_cf.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so,
                               src])
    L = C.CDLL(so)
    L.cfh_build.restype = C.c_void_p
    L.cfh_build.argtypes = [C.c_int64, C.c_int64, C.c_double]
    L.cfh_sizes.argtypes = [C.c_void_p, I64P]
    L.cfh_export.argtypes = [C.c_void_p, F64P, F64P] + [I64P] * 7
    L.cfh_free.argtypes = [C.c_void_p]
    return L


def cf_mesh(L, nt, nr, sp):
    h = L.cfh_build(nt, nr, float(sp))
    s = np.zeros(6, np.int64)
    L.cfh_sizes(h, s)
    n, nel, se, nnz, hr, sn = (int(v) for v in s)
    out = dict(n=n, nel=nel, theta=np.zeros(n), r=np.zeros(n), e2n_off=np.zeros(nel + 1, np.int64),
               e2n_idx=np.zeros(se, np.int64), g_off=np.zeros(n + 1, np.int64), g_idx=np.zeros(nnz, np.int64),
               halo=np.zeros(max(2 * hr, 1), np.int64), nbr_off=np.zeros(nel + 1, np.int64),
               nbr_idx=np.zeros(max(sn, 1), np.int64), halo_rows=hr)
    L.cfh_export(h, out["theta"], out["r"], out["e2n_off"], out["e2n_idx"], out["g_off"], out["g_idx"],
                 out["halo"], out["nbr_off"], out["nbr_idx"])
    L.cfh_free(h)
    out["halo"] = out["halo"][:2 * hr]
    out["nbr_idx"] = out["nbr_idx"][:sn]
    return out


@pytest.mark.parametrize("nt,nr,sp", [(8, 2, 500.0), (24, 6, 300.0), (36, 10, 100.0), (37, 11, 77.7), (180, 50, 50),
                                      (180, 50, 20), (90, 25, 5.0), (64, 3, 0.9), (12, 40, 33.0), (720, 200, 20)])
def test_closed_form_equals_literal_restatement(cfh, annulus, nt, nr, sp):
    m = annulus(nt, nr, sp)
    c = cf_mesh(cfh, nt, nr, sp)
    assert (c["n"], c["nel"], c["halo_rows"]) == (m.n, m.nel, m.halo_rows)
    assert np.array_equal(c["theta"], m.theta) and np.array_equal(c["r"], m.r)
    assert np.array_equal(c["e2n_off"], m.e2n_off) and np.array_equal(c["e2n_idx"], m.e2n_idx)
    assert np.array_equal(c["g_off"] + 1, m.G_colptr) and np.array_equal(c["g_idx"], m.G_rowval)
    assert np.array_equal(c["halo"], m.halo)
    # neighbour lists: same sets (the closed form emits them ascending, the reference in discovery order)
    assert np.array_equal(c["nbr_off"], m.nbr_off)
    for e in range(0, m.nel, max(1, m.nel // 500)):
        assert sorted(m.nbr_idx[m.nbr_off[e]:m.nbr_off[e + 1]]) == list(c["nbr_idx"][c["nbr_off"][e]:c["nbr_off"][e + 1]])
