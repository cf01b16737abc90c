"""Host part of the magic-number division used by the 3-D push kernels (csrc/fastdiv.h): exact for every divisor /
numerator combination the kernels can meet, including the extremes."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fastdiv_exact():
    so = os.path.join(HERE, "libfastdiv_test.so")
    src = os.path.join(HERE, "fastdiv_driver.cpp")
    hdr = os.path.join(HERE, "..", "raytracer.jl_b200", "csrc", "fastdiv.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.fastdiv_mismatches.restype = C.c_longlong
    L.fastdiv_mismatches.argtypes = [C.c_uint, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]
    divisors = [1, 2, 3, 5, 7, 11, 12, 32, 33, 216, 368, 1000, 46656, 135424, 65535, 65536, 2 ** 31 - 1, 2 ** 31,
                2 ** 32 - 1]
    for d in divisors:
        assert L.fastdiv_mismatches(d, 0, 200000, 1) == 0            # small numerators, every value
        assert L.fastdiv_mismatches(d, 0, 2 ** 32 - 1, 65521) == 0   # whole range, prime stride
        assert L.fastdiv_mismatches(d, 2 ** 32 - 70000, 2 ** 32 - 1, 1) == 0  # top of the range
        for k in (1, 2, 1000, 2 ** 32 // d):                         # around multiples of d
            lo = max(0, k * d - 3)
            assert L.fastdiv_mismatches(d, lo, min(2 ** 32 - 1, k * d + 3), 1) == 0


def test_float_row_index_is_exact():
    """push3d_body (grid3d.cu) splits the flat target index t of a z-plane into (row, column) with
    r = __float2int_rz((float(t) + 0.5f) * (1.0f / float(ncol))): exact for every size a unit can have
    (ncol <= 32 + 2w columns, <= 2w + 1 rows, w <= 7)."""
    import numpy as np
    for ncol in range(1, 47):
        rcol = np.float32(1.0) / np.float32(ncol)
        t = np.arange(0, ncol * 15, dtype=np.int64)
        r = np.trunc((t.astype(np.float32) + np.float32(0.5)) * rcol).astype(np.int64)
        assert np.array_equal(r, t // ncol), ncol
